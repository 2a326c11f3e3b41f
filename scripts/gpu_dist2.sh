#!/bin/bash
# 2-GPU pass: GPU test-suite (incl. the 2-rank tests), kernel-class timings at N=1 and N=2, SpMV A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15
bash scripts/gpu_kt.sh 2
python scripts/bench_variants.py film20m 1.0 2>&1 | tail -3

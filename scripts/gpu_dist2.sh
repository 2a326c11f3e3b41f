#!/bin/bash
# 2-GPU pass: GPU test-suite (incl. the 2-rank tests), kernel-class timings at N=1 and N=2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
bash scripts/gpu_kt.sh 2

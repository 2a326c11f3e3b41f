#!/bin/bash
# single-GPU kernel-class timings at 1/8 of film20m (the per-rank size of an 8-GPU run)
mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --kernel-times --scale 0.3536 > gpurun_out/kt_s8.json 2> gpurun_out/kt_s8.err
grep -E "rank|bench:" gpurun_out/kt_s8.err

#!/bin/bash
# r01r: quaternion basis (32 B) + closed-form node-diagonal (16 B) in the Krylov kernels: parity + timing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --kernel-times > gpurun_out/kt_n1.json 2> gpurun_out/kt_n1.err
grep -E "rank|bench:" gpurun_out/kt_n1.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_film20m.json 2> gpurun_out/bench_film20m.err
cat gpurun_out/bench_film20m.json | cut -c1-300

#!/bin/bash
# GPU test-suite (no -x) + one ncu --set full capture of the matrix-free SpMV and the p kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv_node3 -s 8 -c 2 -f -o gpurun_out/prof_k_spmv_node3 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_spmv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bicg_p_node -s 8 -c 1 -f -o gpurun_out/prof_k_bicg_p_node \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_p.log 2>&1
ls -la gpurun_out/*.ncu-rep

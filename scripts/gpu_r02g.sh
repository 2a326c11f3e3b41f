#!/bin/bash
# r02g: staged element kernels (k_tet_st), full-size trajectory tests, anisotropic workload
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02g_pytest_gpu.log
tail -4 gpurun_out/r02g_pytest_gpu.log
for cfg in "film20m st FG_X=1" "film20m nost FG_TET_NOSTAGE=1" "tube5m st FG_X=1" "film20m_k st FG_X=1" "film20m_k nost FG_TET_NOSTAGE=1"; do
    set -- $cfg
    env $3 timeout 300 python bench.py --workload $1 --steps 20 --warmup 5 --no-cpu-baseline --traffic off --no-e2e --kernel-times \
        > gpurun_out/r02g_bench_$1_$2.json 2> gpurun_out/r02g_bench_$1_$2.err
    echo "$1 $2 rc=$? $(cut -c1-130 gpurun_out/r02g_bench_$1_$2.json)"
    grep -E "rank 0 (basis|tet|assemble|solve) " gpurun_out/r02g_bench_$1_$2.err
done

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_solver_paths.py tests/test_gpu_parity.py tests/test_full_test_golden.py -x -q 2>&1 | tail -4
for cfg in "film20m_k lean FG_X=1" "film20m_k nolean FG_TET_NOLEAN=1" "film20m w27 FG_X=1"; do
    set -- $cfg
    env $3 timeout 300 python bench.py --workload $1 --steps 20 --warmup 5 --no-cpu-baseline --traffic off --no-e2e --kernel-times \
        > gpurun_out/r02m_bench_$1_$2.json 2> gpurun_out/r02m_bench_$1_$2.err
    echo "$1 $2 rc=$? $(cut -c1-130 gpurun_out/r02m_bench_$1_$2.json)"
    grep -E "rank 0 (basis|tet|assemble|solve) " gpurun_out/r02m_bench_$1_$2.err
done
# adaptive warps: 1/8-size problem on one GPU, 32 warps against the picked count
for wv in 32 0; do
  FG_PK_WARPS=$wv timeout 300 python bench.py --scale 0.354 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --traffic off --kernel-times \
      2>&1 >/dev/null | grep -E "rank 0 solve\.(kernel|B_spmv|D_spmv|A_p)" | sed "s/^/warps=$wv /"
done
for wl in tube5m disk1m sp4; do
    timeout 300 python bench.py --workload $wl --steps 40 --warmup 5 --no-cpu-baseline --traffic off --no-e2e 2>/dev/null | cut -c1-110
done

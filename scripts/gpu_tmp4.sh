#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_solver_paths.py -x -q 2>&1 | tail -3
for wl in film20m disk1m sp4; do
    timeout 300 python bench.py --workload $wl --steps 40 --warmup 5 --no-cpu-baseline --traffic off --no-e2e --no-named-meshes --kernel-times \
        > gpurun_out/r02p_bench_$wl.json 2> gpurun_out/r02p_bench_$wl.err
    echo "$wl rc=$? $(cut -c1-120 gpurun_out/r02p_bench_$wl.json)"
    grep -E "rank 0 solve\." gpurun_out/r02p_bench_$wl.err | cut -c1-130
done

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_solver_paths.py tests/test_gpu_parity.py -x -q 2>&1 | tail -4
for wl in film20m sp4 disk1m ellipsoid; do
    timeout 300 python bench.py --workload $wl --steps 40 --warmup 5 --no-cpu-baseline --traffic off --no-e2e --no-named-meshes --kernel-times \
        > gpurun_out/r02n_bench_$wl.json 2> gpurun_out/r02n_bench_$wl.err
    echo "$wl rc=$? $(cut -c1-120 gpurun_out/r02n_bench_$wl.json)"
    grep -E "rank 0 solve\." gpurun_out/r02n_bench_$wl.err | cut -c1-130
done

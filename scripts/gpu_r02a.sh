#!/bin/bash
# r02a: first run of the persistent solve kernel: GPU suite, A/B against the multi-kernel path
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02a_pytest_gpu.log
tail -5 gpurun_out/r02a_pytest_gpu.log
for wl in film20m tube5m disk1m sp4 ellipsoid; do
  for sv in persistent multi; do
    timeout 300 python bench.py --workload $wl --solver $sv --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --kernel-times \
        > gpurun_out/r02a_bench_${wl}_${sv}.json 2> gpurun_out/r02a_bench_${wl}_${sv}.err
    echo "$wl $sv rc=$? $(cut -c1-160 gpurun_out/r02a_bench_${wl}_${sv}.json)"
    grep -E "solve\.|rank 0 (basis|tet|assemble|update|solve|gaps)" gpurun_out/r02a_bench_${wl}_${sv}.err | head -20
  done
done

#!/bin/bash
# r01n: deeper element prefetch pipeline; full GPU suite incl. full-size trajectories vs the oracle
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --kernel-times > gpurun_out/kt_n1.json 2> gpurun_out/kt_n1.err
grep -E "rank|bench:" gpurun_out/kt_n1.err
FG_TET_NOPIPE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --kernel-times > gpurun_out/kt_n1_nopipe.json 2> gpurun_out/kt_n1_nopipe.err
grep -E "rank|bench:" gpurun_out/kt_n1_nopipe.err | grep -E "tet|timed"
for wl in tube5m disk1m sp4; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload $wl > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
cat gpurun_out/bench_$wl.json | cut -c1-400
done

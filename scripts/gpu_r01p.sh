#!/bin/bash
# r01p: cp.async-staged element kernel: parity + A/B timing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --kernel-times > gpurun_out/kt_n1.json 2> gpurun_out/kt_n1.err
grep -E "rank|bench:" gpurun_out/kt_n1.err
FG_TET_NOASYNC=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --kernel-times > gpurun_out/kt_n1_noasync.json 2> gpurun_out/kt_n1_noasync.err
grep -E "rank|bench:" gpurun_out/kt_n1_noasync.err | grep -E "tet|timed"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --kernel-times --workload tube5m > gpurun_out/kt_tube.json 2> gpurun_out/kt_tube.err
grep -E "rank|bench:" gpurun_out/kt_tube.err | grep -E "tet|timed|bench"
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_film20m.json 2> gpurun_out/bench_film20m.err
cat gpurun_out/bench_film20m.json | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tet_iso_async -s 3 -c 1 -f -o gpurun_out/prof_k_tet_iso_async \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_k_tet_iso_async.log 2>&1

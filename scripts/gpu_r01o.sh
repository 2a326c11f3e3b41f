#!/bin/bash
# r01o: basis-free element records (projection once per node in the assembly): parity + timing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --kernel-times > gpurun_out/kt_n1.json 2> gpurun_out/kt_n1.err
grep -E "rank|bench:" gpurun_out/kt_n1.err
FG_TET_NOPIPE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --kernel-times > gpurun_out/kt_n1_nopipe.json 2> gpurun_out/kt_n1_nopipe.err
grep -E "rank|bench:" gpurun_out/kt_n1_nopipe.err | grep -E "tet|timed"
FG_NO_ISO=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --kernel-times > gpurun_out/kt_n1_noiso.json 2> gpurun_out/kt_n1_noiso.err
grep -E "rank|bench:" gpurun_out/kt_n1_noiso.err | grep -E "tet|timed"
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_film20m.json 2> gpurun_out/bench_film20m.err
cat gpurun_out/bench_film20m.json | cut -c1-300

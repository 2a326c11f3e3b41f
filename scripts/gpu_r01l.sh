#!/bin/bash
# r01l: element fast path (k_tet_iso) + 256-bit record loads: parity, full-size properties, A/B timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 1200 python -m pytest tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_fullsize.log
tail -25 gpurun_out/pytest_fullsize.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --kernel-times > gpurun_out/kt_n1.json 2> gpurun_out/kt_n1.err
grep -E "rank|bench:" gpurun_out/kt_n1.err
FG_NO_ISO=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --kernel-times > gpurun_out/kt_n1_noiso.json 2> gpurun_out/kt_n1_noiso.err
grep -E "rank|bench:" gpurun_out/kt_n1_noiso.err | grep -E "tet|assemble|timed"
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_film20m.json 2> gpurun_out/bench_film20m.err
cat gpurun_out/bench_film20m.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tet_iso -s 3 -c 1 -f -o gpurun_out/prof_k_tet_iso \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_k_tet_iso.log 2>&1

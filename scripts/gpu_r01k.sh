#!/bin/bash
# r01k: 256-bit gathers + 16-bit column offsets + no stored phat/shat: parity, full-size properties, A/B timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 1200 python -m pytest tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_fullsize.log
tail -25 gpurun_out/pytest_fullsize.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --kernel-times > gpurun_out/kt_n1.json 2> gpurun_out/kt_n1.err
grep -E "rank|bench:" gpurun_out/kt_n1.err
FG_NO_COL16=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --kernel-times > gpurun_out/kt_n1_col32.json 2> gpurun_out/kt_n1_col32.err
grep -E "rank|bench:" gpurun_out/kt_n1_col32.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --kernel-times --workload tube5m > gpurun_out/kt_tube.json 2> gpurun_out/kt_tube.err
grep -E "rank|bench:" gpurun_out/kt_tube.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_film20m.json 2> gpurun_out/bench_film20m.err
cat gpurun_out/bench_film20m.json

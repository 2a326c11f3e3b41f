#!/bin/bash
# r01z: final state of round 1 — GPU suite, bench line, launch list, SpMV full capture (traffic)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_film20m.json 2> gpurun_out/bench_film20m.err
cat gpurun_out/bench_film20m.json | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_film20m.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_spmv_node3 -s 8 -c 2 -f -o gpurun_out/prof_k_spmv_node3 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_spmv.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1

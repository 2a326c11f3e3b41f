#!/bin/bash
# tests + bench + launch list + full captures of the top kernels (SELL layout)
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_film20m.json 2> gpurun_out/bench_film20m.err
cat gpurun_out/bench_film20m.json; tail -5 gpurun_out/bench_film20m.err
python bench.py --steps 10 --warmup 3 --workload tube5m --no-cpu-baseline > gpurun_out/bench_tube5m.json 2> gpurun_out/bench_tube5m.err
cat gpurun_out/bench_tube5m.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_film20m.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spmv_sell -s 8 -c 2 -f -o gpurun_out/prof_k_spmv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_k_spmv.log 2>&1
for k in k_tet k_assemble; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_$k \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out

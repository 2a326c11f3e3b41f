#!/bin/bash
# r01j: GPU tests + bench + ncu launch list + full captures of the matrix-free SpMV, the element
# kernel, the node assembly and the fused p kernel (film20m, 1 GPU)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_film20m.json 2> gpurun_out/bench_film20m.err
cat gpurun_out/bench_film20m.json; tail -3 gpurun_out/bench_film20m.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --kernel-times > gpurun_out/kt_n1.json 2> gpurun_out/kt_n1.err
grep -E "rank|bench:" gpurun_out/kt_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_film20m.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv_node3 -s 8 -c 2 -f -o gpurun_out/prof_k_spmv_node3 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_spmv.log 2>&1
for k in k_tet k_assemble_node k_bicg_p_node k_bicg_xr; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_$k \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out

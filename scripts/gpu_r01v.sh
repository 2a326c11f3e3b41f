#!/bin/bash
# r01v: the judged profile pass of the final round-1 state (film20m, 1 GPU): tests, bench (both arms),
# ncu launch list, ncu --set full of every kernel class of the step
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_film20m.json 2> gpurun_out/bench_film20m.err
cat gpurun_out/bench_film20m.json | cut -c1-250
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err
cat gpurun_out/bench_reference_arm.json | cut -c1-250
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --kernel-times > gpurun_out/kt_n1.json 2> gpurun_out/kt_n1.err
grep -E "rank|bench:" gpurun_out/kt_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_film20m.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv_node3 -s 8 -c 2 -f -o gpurun_out/prof_k_spmv_node3 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_spmv.log 2>&1
for k in k_tet_iso k_assemble_node k_bicg_p_node k_bicg_s_node k_bicg_xr_node k_update k_basis; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_$k \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$k.log 2>&1
done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
ls gpurun_out | wc -l

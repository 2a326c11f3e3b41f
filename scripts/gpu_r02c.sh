#!/bin/bash
# r02c: balanced CTA slice ranges, result mailbox, lazy evolution: new solver-path tests, GPU suite, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_solver_paths.py -x -q 2>&1 | tail -15 > gpurun_out/r02c_pytest_solver_paths.log
tail -4 gpurun_out/r02c_pytest_solver_paths.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02c_pytest_gpu.log
tail -4 gpurun_out/r02c_pytest_gpu.log
for wl in sp4 disk1m tube5m ellipsoid; do
    timeout 300 python bench.py --workload $wl --steps 40 --warmup 5 --no-cpu-baseline --traffic off \
        > gpurun_out/r02c_bench_${wl}.json 2> gpurun_out/r02c_bench_${wl}.err
    echo "$wl rc=$? $(python - <<PY
import json
d=json.load(open("gpurun_out/r02c_bench_${wl}.json"))
r=d["roofline"]
print("%.1f steps/s  e2e %.1f  solve %.1f us  spmv %.1f us frac %.2f step_roof %.3f iters %.2f" % (d["value"], d["e2e"]["value"], r["us_per_launch"], r["spmv_phase"]["us"], r["spmv_phase"]["frac"], d["step_roofline"]["frac"], d["config"]["mean_bicgstab_iters"]))
PY
)"
done
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02c_bench_film20m.json 2> gpurun_out/r02c_bench_film20m.err
echo "film20m rc=$?"; cat gpurun_out/r02c_bench_film20m.json
FG_PK_MAILBOX=0 FG_EAGER_COMMIT=1 timeout 300 python bench.py --workload sp4 --steps 40 --warmup 5 --no-cpu-baseline --traffic off --no-e2e 2>/dev/null | cut -c1-150
timeout 300 python bench.py --workload film20m --steps 20 --warmup 5 --no-cpu-baseline --traffic off --no-e2e --kernel-times 2>&1 >/dev/null | grep -E "rank 0" 

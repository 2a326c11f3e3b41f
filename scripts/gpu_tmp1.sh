#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_solver_paths.py tests/test_gpu_parity.py -x -q 2>&1 | tail -5
for wl in film20m tube5m disk1m sp4 ellipsoid; do
    timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --traffic off \
        > gpurun_out/r02i_bench_${wl}.json 2> gpurun_out/r02i_bench_${wl}.err
    echo "$wl rc=$? $(python - <<PY
import json
d=json.load(open("gpurun_out/r02i_bench_${wl}.json"))
r=d["roofline"]
print("%.1f steps/s  e2e %.1f  solve %.1f us  frac %.3f spmv %.1f us frac %.2f step_roof %.3f iters %.2f" % (d["value"], d["e2e"]["value"], r["us_per_launch"], r["frac"], r["spmv_phase"]["us"], r["spmv_phase"]["frac"], d["step_roofline"]["frac"], d["config"]["mean_bicgstab_iters"]), {k:round(v["us"],1) for k,v in r["phases"].items()})
PY
)"
done

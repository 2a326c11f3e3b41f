#!/bin/bash
# r02e: compact-block ordering (RCB) + shared-memory staged gathers in the persistent SpMV
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02e_pytest_gpu.log
tail -4 gpurun_out/r02e_pytest_gpu.log
run() { # name env...
  local tag=$1; shift
  for wl in film20m tube5m disk1m; do
    env "$@" timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --traffic off --no-e2e --kernel-times \
        > gpurun_out/r02e_bench_${wl}_${tag}.json 2> gpurun_out/r02e_bench_${wl}_${tag}.err
    echo "$tag $wl rc=$? $(python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02e_bench_${wl}_${tag}.json"))
    r=d["roofline"]
    print("%.1f steps/s solve %.1f us frac %.3f spmv %.1f us frac %.2f step_roof %.3f iters %.2f colB %d" % (d["value"], r["us_per_launch"], r["frac"], r["spmv_phase"]["us"], r["spmv_phase"]["frac"], d["step_roofline"]["frac"], d["config"]["mean_bicgstab_iters"], r["col_index_bytes"]), {k:round(v["us"],1) for k,v in r["phases"].items()})
except Exception as e: print("ERR", e)
PY
)"
    grep -E "rank 0 (basis|tet|assemble) " gpurun_out/r02e_bench_${wl}_${tag}.err
  done
}
run staged FG_X=1
run rcb_nostage FG_NO_STAGE=1
run window FG_ORDER=window
FG_STAGE_MIN_BLOCKS=0 timeout 300 python bench.py --workload sp4 --steps 20 --warmup 5 --no-cpu-baseline --traffic off --no-e2e 2>/dev/null | cut -c1-120
timeout 300 python bench.py --workload sp4 --steps 20 --warmup 5 --no-cpu-baseline --traffic off --no-e2e 2>/dev/null | cut -c1-120
FG_ORDER=window timeout 300 python bench.py --workload sp4 --steps 20 --warmup 5 --no-cpu-baseline --traffic off --no-e2e 2>/dev/null | cut -c1-120

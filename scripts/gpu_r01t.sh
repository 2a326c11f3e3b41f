#!/bin/bash
# r01t: SpMV occupancy A/B (4 CTAs/SM x 64 registers against 5 x 48)
mkdir -p gpurun_out
for cps in 4 5; do
FG_SPMV_CPS=$cps timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --kernel-times > gpurun_out/kt_n1_cps$cps.json 2> gpurun_out/kt_n1_cps$cps.err
echo "== FG_SPMV_CPS=$cps"; grep -E "rank" gpurun_out/kt_n1_cps$cps.err | grep -E "spmv|timed"
FG_SPMV_CPS=$cps timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --kernel-times --workload tube5m > gpurun_out/kt_tube_cps$cps.json 2> gpurun_out/kt_tube_cps$cps.err
grep -E "rank" gpurun_out/kt_tube_cps$cps.err | grep -E "spmv|timed"
done
FG_SPMV_CPS=5 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3

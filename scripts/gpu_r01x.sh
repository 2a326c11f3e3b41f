#!/bin/bash
# r01x: SpMV with early L2 prefetch of the row-epilogue operands (A/B)
mkdir -p gpurun_out
for pf in 0 1 0 1; do
FG_SPMV_PF=$pf timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --kernel-times > gpurun_out/kt_n1_pf$pf.json 2> gpurun_out/kt_n1_pf$pf.err
echo "== FG_SPMV_PF=$pf"; grep -E "rank" gpurun_out/kt_n1_pf$pf.err | grep -E "spmv|timed"
done

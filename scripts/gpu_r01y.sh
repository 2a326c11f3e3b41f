#!/bin/bash
# r01y: SpMV prefetch levels (1: row-epilogue operands, 2: + first batch of the next slice), A/B + parity
mkdir -p gpurun_out
for pf in 1 2 1 2; do
FG_SPMV_PF=$pf timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --kernel-times > gpurun_out/kt_n1_pf$pf.json 2> gpurun_out/kt_n1_pf$pf.err
echo "== FG_SPMV_PF=$pf"; grep -E "rank" gpurun_out/kt_n1_pf$pf.err | grep -E "spmv|timed"
done
FG_SPMV_PF=2 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
FG_SPMV_PF=2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --kernel-times --workload tube5m 2>&1 | grep -E "rank" | grep -E "spmv|timed"
FG_SPMV_PF=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --kernel-times --workload tube5m 2>&1 | grep -E "rank" | grep -E "spmv|timed"

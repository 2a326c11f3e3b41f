#!/bin/bash
# r01u2: tiny permutation windows (rows of a slice stay mesh neighbours; more padding): does gather coalescing show?
mkdir -p gpurun_out
for w in 32 64 96; do
FG_SELL_WINDOW=$w timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --kernel-times > gpurun_out/kt_n1_win$w.json 2> gpurun_out/kt_n1_win$w.err
echo "== FG_SELL_WINDOW=$w film20m"; grep -E "rank|nnz" gpurun_out/kt_n1_win$w.err | grep -E "spmv_v|spmv_t|tet|assemble|timed|bicg_p"
done

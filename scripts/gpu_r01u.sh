#!/bin/bash
# r01u: window of the SELL row permutation (locality against padding), film20m + tube5m kernel classes
mkdir -p gpurun_out
for w in 128 256 512 1024 4096; do
FG_SELL_WINDOW=$w timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --kernel-times > gpurun_out/kt_n1_win$w.json 2> gpurun_out/kt_n1_win$w.err
echo "== FG_SELL_WINDOW=$w film20m"; grep -E "rank" gpurun_out/kt_n1_win$w.err | grep -E "spmv_v|spmv_t|tet|assemble|timed"
done
for w in 256 1024; do
FG_SELL_WINDOW=$w timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --kernel-times --workload tube5m > gpurun_out/kt_tube_win$w.json 2> gpurun_out/kt_tube_win$w.err
echo "== FG_SELL_WINDOW=$w tube5m"; grep -E "rank" gpurun_out/kt_tube_win$w.err | grep -E "spmv_v|spmv_t|tet|assemble|timed"
done

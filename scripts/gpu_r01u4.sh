#!/bin/bash
# r01u4: permutation window 96 against 1024 on the other mesh families (tube5m, disk1m)
mkdir -p gpurun_out
for wl in tube5m disk1m; do for w in 96 1024; do
FG_SELL_WINDOW=$w timeout 30 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --kernel-times --workload $wl > gpurun_out/kt_${wl}_win$w.json 2> gpurun_out/kt_${wl}_win$w.err
echo "== $wl window $w"; grep -E "rank" gpurun_out/kt_${wl}_win$w.err | grep -E "spmv_v|spmv_t|tet |timed"
done; done

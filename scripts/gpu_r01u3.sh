#!/bin/bash
# r01u3: parity smoke of the final build + ncu cache/traffic counters of the SpMV for windows 96 and 1024
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
M=gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__data_pipe_lsu_wavefronts_mem_lg.sum,lts__t_sectors_srcunit_tex_op_read.sum,dram__bytes_read.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed
for w in 96 1024; do
FG_SELL_WINDOW=$w timeout 200 ncu --metrics $M --clock-control none -k regex:k_spmv_node3 -s 8 -c 2 --csv --log-file gpurun_out/spmv_counters_win$w.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_win$w.log 2>&1
echo "== window $w"; grep -v "^==" gpurun_out/spmv_counters_win$w.csv | awk -F'","' 'NR>1{print $5, $(NF-2), $(NF-1), $NF}' | tr -d '"' | head -24
done

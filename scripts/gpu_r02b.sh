#!/bin/bash
# r02b (2 GPUs): persistent solve kernel on a slab-partitioned mesh: parity worker, bench A/B, and the
# per-rank size of an 8-GPU run (film20m at half extent on 2 ranks)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR tests/dist_gpu_worker.py > gpurun_out/r02b_dist_worker_n2_persistent.log 2>&1
echo "worker persistent rc=$?"; grep -E "DIST_GPU_OK|Error|error|assert" gpurun_out/r02b_dist_worker_n2_persistent.log | head -5
FG_SOLVER=multi timeout 300 $TR tests/dist_gpu_worker.py > gpurun_out/r02b_dist_worker_n2_multi.log 2>&1
echo "worker multi rc=$?"; grep -E "DIST_GPU_OK|Error|error|assert" gpurun_out/r02b_dist_worker_n2_multi.log | head -5
for sc in 1.0 0.5; do
  for sv in persistent multi; do
    timeout 300 $TR bench.py --gpus 2 --scale $sc --solver $sv --steps 20 --warmup 3 --no-e2e --kernel-times \
        > gpurun_out/r02b_bench_n2_x${sc}_${sv}.json 2> gpurun_out/r02b_bench_n2_x${sc}_${sv}.err
    echo "x$sc $sv rc=$? $(cut -c1-140 gpurun_out/r02b_bench_n2_x${sc}_${sv}.json)"
    grep -E "rank 0 solve\.|rank 0 (basis|tet|assemble|update|solve|halo|gaps|bicg|spmv)" gpurun_out/r02b_bench_n2_x${sc}_${sv}.err | head -24
  done
done

#!/bin/bash
# r01w (8 GPUs): multi-GPU parity worker at 8 ranks + strong-scaling bench at N=8 and N=4
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29548 tests/dist_gpu_worker.py > gpurun_out/dist_worker8.log 2>&1
echo "worker8 rc=$?"; grep -E "OK|Assert|Error" gpurun_out/dist_worker8.log | head -5
for n in 8 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  echo "bench$n rc=$?"; cat gpurun_out/bench_n$n.json | cut -c1-220; tail -2 gpurun_out/bench_n$n.err
done

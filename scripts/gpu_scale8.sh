#!/bin/bash
# 8-GPU pass: multi-GPU parity worker at 8 ranks, strong-scaling bench at N=4 and N=8
set -x
mkdir -p gpurun_out
nvidia-smi topo -m | head -14 > gpurun_out/topo.txt
for n in 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$n --master-addr 127.0.0.1 --master-port 2954$n tests/dist_gpu_worker.py > gpurun_out/dist_worker$n.log 2>&1
  echo "worker$n rc=$?"; grep -E "OK|Assert|Error" gpurun_out/dist_worker$n.log | head -5
done
for n in 4 8; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  echo "bench$n rc=$?"; cat gpurun_out/bench_n$n.json; tail -3 gpurun_out/bench_n$n.err
done

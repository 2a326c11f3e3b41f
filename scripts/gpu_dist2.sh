#!/bin/bash
# 2-GPU pass: multi-GPU parity worker, then bench at N=1 and N=2 on the same box
set -x
mkdir -p gpurun_out
nvidia-smi -L
nvidia-smi topo -m | head -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29541 tests/dist_gpu_worker.py > gpurun_out/dist_worker2.log 2>&1
echo "worker rc=$?"; tail -30 gpurun_out/dist_worker2.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench2 rc=$?"; cat gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err

#!/bin/bash
# r01s (2 GPUs): 2-rank parity of the quaternion-basis Krylov kernels + bench at N=2 and the per-rank size of N=8
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu_dist.log
tail -5 gpurun_out/pytest_gpu_dist.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_film20m_n2.json 2> gpurun_out/bench_film20m_n2.err
cat gpurun_out/bench_film20m_n2.json | cut -c1-260; tail -2 gpurun_out/bench_film20m_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --kernel-times --scale 0.5 > gpurun_out/kt_n2_quarter.json 2> gpurun_out/kt_n2_quarter.err
grep -E "rank 0|bench:" gpurun_out/kt_n2_quarter.err

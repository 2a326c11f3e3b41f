#!/bin/bash
mkdir -p gpurun_out
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR8 tests/dist_gpu_worker.py > gpurun_out/r02k_dist_worker_n8.log 2>&1
echo "worker n8 rc=$? $(grep -c DIST_GPU_OK gpurun_out/r02k_dist_worker_n8.log)"
timeout 300 $TR8 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02k_bench_n8.json 2> gpurun_out/r02k_bench_n8.err
echo "n8 rc=$? $(cut -c1-170 gpurun_out/r02k_bench_n8.json)"
timeout 300 $TR8 bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e --kernel-times > gpurun_out/r02k_bench_n8_kt.json 2> gpurun_out/r02k_bench_n8_kt.err
echo "n8 kt rc=$? $(cut -c1-170 gpurun_out/r02k_bench_n8_kt.json)"
grep -E "rank [03] " gpurun_out/r02k_bench_n8_kt.err | head -50

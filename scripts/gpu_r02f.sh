#!/bin/bash
# r02f (8 GPUs): persistent solve kernel at 8 and 4 ranks: parity worker, bench, per-phase times
mkdir -p gpurun_out
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534"
timeout 300 $TR8 tests/dist_gpu_worker.py > gpurun_out/r02f_dist_worker_n8.log 2>&1
echo "worker n8 rc=$?"; grep -E "DIST_GPU_OK|Error|error|assert" gpurun_out/r02f_dist_worker_n8.log | head -5
timeout 300 $TR8 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02f_bench_n8.json 2> gpurun_out/r02f_bench_n8.err
echo "n8 rc=$? $(cut -c1-170 gpurun_out/r02f_bench_n8.json)"
timeout 300 $TR8 bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e --kernel-times > gpurun_out/r02f_bench_n8_kt.json 2> gpurun_out/r02f_bench_n8_kt.err
echo "n8 kt rc=$? $(cut -c1-170 gpurun_out/r02f_bench_n8_kt.json)"
grep -E "rank 0 " gpurun_out/r02f_bench_n8_kt.err | head -24
timeout 300 $TR4 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02f_bench_n4.json 2> gpurun_out/r02f_bench_n4.err
echo "n4 rc=$? $(cut -c1-170 gpurun_out/r02f_bench_n4.json)"
timeout 300 $TR8 bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e --solver multi > gpurun_out/r02f_bench_n8_multi.json 2> gpurun_out/r02f_bench_n8_multi.err
echo "n8 multi rc=$? $(cut -c1-170 gpurun_out/r02f_bench_n8_multi.json)"

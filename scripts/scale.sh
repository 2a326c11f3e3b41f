#!/bin/bash
# Strong scaling on N GPUs of one box (run under gpurun, from the repo root):
#     gpurun --gpus 8 --timeout 1500 -- 'bash scripts/scale.sh 8 4 2'
# For every N given: the multi-GPU parity worker (slab-partitioned steps against the single-GPU path) and
# the bench line (its `parity` block carries <m> and the summed iteration count, identical at every N).
mkdir -p gpurun_out
for N in "$@"; do
    TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N))"
    timeout 300 $TR tests/dist_gpu_worker.py > gpurun_out/dist_worker_n$N.log 2>&1
    echo "worker n$N rc=$? $(grep -c DIST_GPU_OK gpurun_out/dist_worker_n$N.log)"
    timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
    echo "bench n$N rc=$? $(cut -c1-170 gpurun_out/bench_n$N.json)"
    timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --kernel-times > /dev/null 2> gpurun_out/kernel_classes_n$N.txt
    grep -E "rank 0 " gpurun_out/kernel_classes_n$N.txt | head -24
done

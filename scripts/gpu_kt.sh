#!/bin/bash
# per-kernel-class timing: N=1 full size, N=2 at quarter size (per-rank work = 1/8 of film20m)
mkdir -p gpurun_out
N=${1:-2}
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --kernel-times > gpurun_out/kt_n1.json 2> gpurun_out/kt_n1.err
grep -E "rank|bench:" gpurun_out/kt_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 10 --warmup 3 --scale 0.5 --no-e2e --kernel-times > gpurun_out/kt_n$N.json 2> gpurun_out/kt_n$N.err
grep -E "rank 0|bench:" gpurun_out/kt_n$N.err
cat gpurun_out/kt_n$N.json | cut -c1-400

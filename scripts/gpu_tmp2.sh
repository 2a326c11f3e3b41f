#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR tests/dist_gpu_worker.py > gpurun_out/r02i_dist_worker_n2.log 2>&1
echo "worker rc=$? $(grep -c DIST_GPU_OK gpurun_out/r02i_dist_worker_n2.log)"; grep -E "Error|error|assert" gpurun_out/r02i_dist_worker_n2.log | head -5
FG_SOLVER=multi timeout 300 $TR tests/dist_gpu_worker.py > gpurun_out/r02i_dist_worker_n2_multi.log 2>&1
echo "worker multi rc=$? $(grep -c DIST_GPU_OK gpurun_out/r02i_dist_worker_n2_multi.log)"
for sc in 0.5 1.0; do
    timeout 300 $TR bench.py --gpus 2 --scale $sc --steps 20 --warmup 5 --no-e2e --kernel-times \
        > gpurun_out/r02i_bench_n2_x${sc}.json 2> gpurun_out/r02i_bench_n2_x${sc}.err
    echo "x$sc rc=$? $(cut -c1-140 gpurun_out/r02i_bench_n2_x${sc}.json)"
    grep -E "rank 0 solve\.|rank 0 (basis|tet|assemble|solve|gaps)" gpurun_out/r02i_bench_n2_x${sc}.err | head -24
done

#!/bin/bash
# r01m (2 GPUs): element prefetch pipeline A/B at N=1, GPU suite incl. the 2-rank parity test, bench at N=2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --kernel-times > gpurun_out/kt_n1.json 2> gpurun_out/kt_n1.err
grep -E "rank|bench:" gpurun_out/kt_n1.err
FG_TET_NOPIPE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --kernel-times > gpurun_out/kt_n1_nopipe.json 2> gpurun_out/kt_n1_nopipe.err
grep -E "rank|bench:" gpurun_out/kt_n1_nopipe.err | grep -E "tet|timed"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_film20m_n2.json 2> gpurun_out/bench_film20m_n2.err
cat gpurun_out/bench_film20m_n2.json; tail -3 gpurun_out/bench_film20m_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --kernel-times --scale 0.5 > gpurun_out/kt_n2_quarter.json 2> gpurun_out/kt_n2_quarter.err
grep -E "rank 0|bench:" gpurun_out/kt_n2_quarter.err

#!/bin/bash
# round-end dry run: GPU tests, smoke, both bench arms (what the driver runs)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cat gpurun_out/bench_ref.json
timeout 900 python bench.py > gpurun_out/bench_film20m.json 2> gpurun_out/bench_film20m.err
cat gpurun_out/bench_film20m.json; tail -5 gpurun_out/bench_film20m.err

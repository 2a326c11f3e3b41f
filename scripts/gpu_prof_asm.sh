#!/bin/bash
set -x
mkdir -p gpurun_out
for k in k_tet k_assemble_rows; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_$k \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out

#!/bin/bash
# One profiling pass on ONE B200 (run under gpurun, from the repo root):
#     gpurun --timeout 1800 -- 'bash scripts/profile.sh [workload]'
# then here:  python scripts/summarize_profiles.py <tag> [workload]   (copies the summaries into profiles/)
# Writes into gpurun_out/: the GPU test log, the bench line, the ncu launch list of the same bench command
# (cold-cache, serialised: shares, not absolutes) and `ncu --set full` captures of the top kernels.
WL=${1:-film20m}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --workload $WL --steps 20 --warmup 5 > gpurun_out/bench_$WL.json 2> gpurun_out/bench_$WL.err
cut -c1-200 gpurun_out/bench_$WL.json
timeout 300 python bench.py --workload $WL --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --traffic off --no-named-meshes --kernel-times \
    2> gpurun_out/kernel_classes_$WL.txt > /dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$WL.csv \
    python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --traffic off --no-named-meshes > gpurun_out/ncu_bench.log 2>&1
for k in k_llg_solve k_tet_iso k_assemble_node k_basis; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_$k \
        python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --traffic off --no-named-meshes > gpurun_out/ncu_$k.log 2>&1
done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1

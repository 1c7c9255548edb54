#!/bin/bash
# Round-2 profiling pass (one B200, under gpurun): convert probe, ncu launch list of the bench step, ncu --set full of the
# distance, convert and re-rank kernels.  Outputs under gpurun_out/; tools/summarize_profiles_r2.py turns them into profiles/.
set -x
cd "$(dirname "$0")/.."
python tools/convert_probe.py > gpurun_out/convert_probe_r2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondaries > gpurun_out/bench_under_ncu_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dist_topc -s 2 -c 1 -f -o gpurun_out/prof_dist_r2 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondaries > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:convert_norm -c 4 -f -o gpurun_out/prof_convert_r2 \
    python tools/ncu_convert_target.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:rerank_kernel -s 2 -c 1 -f -o gpurun_out/prof_rerank_r2 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondaries > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
# Round 2, second session: the 1-GPU check list (full GPU test suite, smoke, bench lines, reference arm, ncu launch list and
# a --set full capture of the warp-per-query re-rank).  Outputs under gpurun_out/.
set -x
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time timeout 500 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_c3_r2n.err | tail -1 > gpurun_out/bench_c3_r2n.json ) 2>&1 | grep real
( time timeout 400 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_c3_r2n_ref.json ) 2>&1 | grep real
for w in c2 c4; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_${w}_r2n.json; done
timeout 300 python bench.py --workload c1 --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_c1_r2n.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches_r2n.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondaries > gpurun_out/bench_under_ncu_r2n.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rerank_warp -s 2 -c 1 -f -o gpurun_out/prof_rerank_warp_r2n python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondaries > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rerank_warp -s 2 -c 1 -f -o gpurun_out/prof_rerank_warp_c4_r2n python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline --no-secondaries > /dev/null 2>&1
python tools/bench_line.py gpurun_out/bench_c3_r2n.json gpurun_out/bench_c2_r2n.json gpurun_out/bench_c4_r2n.json gpurun_out/bench_c1_r2n.json
cut -c1-400 gpurun_out/bench_c3_r2n_ref.json
tail -2 gpurun_out/bench_c3_r2n.err

set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_c3.json; cut -c1-300 gpurun_out/bench_c3.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_c3_ref.json; cut -c1-300 gpurun_out/bench_c3_ref.json
for w in c2 c4; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_$w.json; cut -c1-200 gpurun_out/bench_$w.json; done
timeout 300 python bench.py --workload c1 --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_c1.json; cut -c1-200 gpurun_out/bench_c1.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:dist_topc -s 2 -c 1 -f -o gpurun_out/prof_dist_r1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_bench.log 2>&1
ls -la gpurun_out | tail -12

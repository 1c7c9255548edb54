#!/bin/bash
# development aid: is a per-GPU slowdown at N=2 caused by sharing the box, by NCCL, or by the half-size shard?
set -x
echo "== (a) single process, GPU0 only, full C3"
timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms/step', d['ms_per_step'], 'dist', d['kernel_ms_per_step'], d['clocks'])"
echo "== (b) single process, GPU0 only, half pool (150k) via small custom workload"
timeout 200 python tools/perf_probe_half.py 2>&1 | tail -3
echo "== (c) two independent processes, one per GPU, half pool each, no NCCL"
CUDA_VISIBLE_DEVICES=0 timeout 200 python tools/perf_probe_half.py > gpurun_out/half0.log 2>&1 &
CUDA_VISIBLE_DEVICES=1 timeout 200 python tools/perf_probe_half.py > gpurun_out/half1.log 2>&1 &
wait
tail -2 gpurun_out/half0.log gpurun_out/half1.log

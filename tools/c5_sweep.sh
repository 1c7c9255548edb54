# schedule / re-rank probes on the one-rank share of configs[4] (d = 49152); usage: bash tools/c5_sweep.sh [workload]
WL=${1:-c5s}
run() { echo "== $*"; env "$@" timeout 280 python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(round(d['ms_per_step'],1), {k:round(v,1) for k,v in d['kernel_ms_per_step'].items()}, 'unc',d['uncertified_per_step'], 'TF',round(d['roofline']['achieved']), 'e2e',round(d['e2e']['value']), [v for k,v in d.items() if k.startswith('self_check')], d['clocks']['sm_mhz'], d['clocks']['power_w_median'])"; }
run B200KNN_WIDE=1
run B200KNN_WIDE=0
run B200KNN_WIDE=1 B200KNN_CTA_GROUP=1

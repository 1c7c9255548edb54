"""Turn the ncu outputs a gpurun call brought back (gpurun_out/) into the tracked summaries under profiles/."""
import collections, csv, io, json, subprocess, sys
rows = [r for r in csv.reader(open('gpurun_out/launches_r1.csv')) if r and r[0].isdigit()]
hdr = None
for r in csv.reader(open('gpurun_out/launches_r1.csv')):
    if r and r[0] == "ID":
        hdr = r; break
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict(); tot = 0.0
for r in rows:
    name = r[ik].split('(')[0]; v = float(r[iv].replace(',', '')); u = r[iu]
    ms = v / 1e6 if u in ('ns', 'nsecond') else (v / 1e3 if u in ('us', 'usecond') else v)
    agg.setdefault(name, []).append(ms); tot += ms
out = ["# ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` (workload c3), round 1",
       "# command: ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv ...   (first 150 launches; times are cold-cache, serialised: compare SHARES)",
       "# kernel | launches | total ms | share of listed GPU time | mean ms"]
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    out.append("%s | %d | %.3f | %.1f%% | %.4f" % (k, len(v), sum(v), 100 * sum(v) / tot, sum(v) / len(v)))
open('profiles/launches_r1_c3.txt', 'w').write("\n".join(out) + "\n")
print("\n".join(out[:12]))
txt = subprocess.run("ncu -i gpurun_out/prof_dist_r1.ncu-rep --page raw --csv", shell=True, capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(txt))); h, u, v = rr[0], rr[1], rr[2]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__cluster_size', 'launch__shared_mem_per_block_dynamic', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'lts__t_sectors_srcunit_tex_lookup_hit.sum', 'lts__t_sectors_srcunit_tex_lookup_miss.sum', 'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
d = {a: (c, b) for a, b, c in zip(h, u, v) if a in want}
lines = ["# ncu --set full --clock-control none --import-source on -k regex:dist_topc -s 2 -c 1  python bench.py --steps 1 --warmup 3 (workload c3), round 1",
         "# one launch of the dominant kernel (first pass over 30000 queries x 300000 pool rows x 3072 dims, float64 features)"]
lines += ["%s = %s %s" % (k, d[k][0], d[k][1]) for k in want if k in d]
open('profiles/dist_kernel_ncu_r1_c3.txt', 'w').write("\n".join(lines) + "\n")
print("\n".join(lines))
num = lambda k: float(d[k][0].replace(',', ''))
scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1, 'Tbyte': 1e12}
traffic = num('dram__bytes_read.sum') * scale[d['dram__bytes_read.sum'][1]] + num('dram__bytes_write.sum') * scale[d['dram__bytes_write.sum'][1]]
try:
    allw = json.load(open('profiles/dist_kernel_ncu.json'))       # other workloads' entries are kept
except Exception:
    allw = {}
allw["c3"] = {"dram_bytes_per_launch": traffic, "source": "profiles/dist_kernel_ncu_r1_c3.txt (ncu --set full, one launch)",
              "algorithmic_operand_floor_bytes": 2 * (300000 + 30000) * 3072}
json.dump(allw, open('profiles/dist_kernel_ncu.json', 'w'), indent=1)
print(traffic / 1e9, "GB")

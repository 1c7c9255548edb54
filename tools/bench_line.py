"""Print the numbers of bench.py JSON lines that matter when iterating: python tools/bench_line.py file.json ..."""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:                                      # an empty or truncated file: say so, keep going
        print(f, "unreadable:", e)
        continue
    e2e = d.get("e2e") or {}
    sc = d.get("small_call") or {}
    print(f.split("/")[-1], "q/s %.0f" % d["value"], "ms %.3f" % d["ms_per_step"], "e2e_ms %.3f" % e2e.get("ms_per_step", float("nan")),
          {k: round(v, 3) for k, v in d.get("kernel_ms_per_step", {}).items()}, "unc", d.get("uncertified_per_step"),
          "frac %.3f" % d.get("roofline", {}).get("frac", float("nan")), "small_ms", sc.get("latency_ms_median"), "pageable_ms", (d.get("e2e_pageable") or {}).get("ms_per_step"), "ok", d.get("self_check"))

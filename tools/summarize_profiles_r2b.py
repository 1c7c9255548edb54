"""Round 2, second session: summaries of the warp-per-query re-rank captures under profiles/.
usage: python tools/summarize_profiles_r2b.py <rep> <dst> <header line> [<header line> ...]"""
import sys

sys.argv, args = sys.argv[:1], sys.argv[1:]
import importlib.util, os
spec = importlib.util.spec_from_file_location("r2", os.path.join(os.path.dirname(__file__), "summarize_profiles_r2.py"))
src = open(spec.origin).read().split("launch_list('gpurun_out/launches_r2.csv'")[0]      # the helpers only, not the round-2 driver code
ns = {}
exec(compile(src, spec.origin, "exec"), ns)
ns["WANT"] += ['smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warp_latency_per_inst_issued.ratio',
               'sm__maximum_warps_per_active_cycle_pct', 'l1tex__t_sector_hit_rate.pct']
ns["summary"](args[0], args[1], ["# " + a for a in args[2:]])
print(open(args[1]).read())

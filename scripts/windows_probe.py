"""One chain at N records: per-phase averages over windows of sweeps (early / late) and one traced
sweep after each window.  usage: windows_probe.py N kind "25:10,200:10" [opt=value ...]
(window "s:n": burn to sweep s, time n sweeps)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from exchanger_b200.sampler import Sampler
from exchanger_b200.synthetic import synth_model

n = int(float(sys.argv[1])); kind = sys.argv[2]
windows = [tuple(int(v) for v in w.split(":")) for w in sys.argv[3].split(",")]
model, _ = synth_model(n, bench.clust_prior(kind, n), seed=bench.DATA_SEED)
g = Sampler(model, seed=1, device=0)
for kv in sys.argv[4:]:
    k, v = kv.split("="); g.set_option(k, int(v))
keys = ("ms_clust", "ms_prep", "ms_index", "ms_cand", "ms_cand_long", "ms_wide", "ms_rounds", "ms_canon", "ms_entities", "ms_distort", "ms_total")
done = 0
for start, nt in windows:
    if start > done:
        g.sweeps(start - done); done = start
    acc = {k: 0.0 for k in keys}; launches = 0; tot = 0.0
    cnt = {k: 0 for k in ("n_rounds", "n_wide", "n_long", "n_probe", "n_tables", "n_listed", "n_scanned", "n_huge", "n_slow")}
    for _ in range(nt):
        ms, nl = g.sweeps_timed(1); st = g.stats(); launches += nl; tot += ms; done += 1
        for k in keys: acc[k] += st[k]
        for k in cnt: cnt[k] += st[k]
    print("window %d+%d opts %s: ms/sweep %.2f launches/sweep %.0f" % (start, nt, sys.argv[4:], tot / nt, launches / nt),
          {k[3:]: round(v / nt, 2) for k, v in acc.items()}, {k: round(v / nt, 1) for k, v in cnt.items()},
          {k: st[k] for k in ("n_entities", "n_links_changed")}, flush=True)
    if os.environ.get("EXB200_PROBE_TRACE", "1") == "1":
        sys.stderr.write("---- traced sweep %d\n" % (done + 1)); sys.stderr.flush()
        g.set_option("trace", 1); g.sweeps(1); g.set_option("trace", 0); done += 1

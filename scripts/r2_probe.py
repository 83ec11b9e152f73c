"""Round-2 probe: timed sweeps with per-phase stats at a given size (after a burn-in), optional option overrides.
usage: r2_probe.py N burn n_timed kind [opt=value ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from exchanger_b200.sampler import Sampler
from exchanger_b200.synthetic import synth_model

n = int(float(sys.argv[1])); burn = int(sys.argv[2]); nt = int(sys.argv[3]); kind = sys.argv[4]
model, _ = synth_model(n, bench.clust_prior(kind, n), seed=bench.DATA_SEED)
g = Sampler(model, seed=1, device=0)
for kv in sys.argv[5:]:
    k, v = kv.split("="); g.set_option(k, int(v))
g.sweeps(burn)
keys = ("ms_clust", "ms_prep", "ms_index", "ms_cand", "ms_cand_long", "ms_wide", "ms_rounds", "ms_canon", "ms_entities", "ms_distort", "ms_total")
acc = {k: 0.0 for k in keys}; launches = 0; tot = 0.0
for _ in range(nt):
    ms, nl = g.sweeps_timed(1); st = g.stats(); launches += nl; tot += ms
    for k in keys: acc[k] += st[k]
print("opts", sys.argv[5:], "ms/sweep %.2f launches/sweep %.0f" % (tot / nt, launches / nt),
      {k[3:]: round(v / nt, 2) for k, v in acc.items()},
      {k: st[k] for k in ("n_rounds", "n_wide", "n_long", "n_probe", "n_tables", "n_entities", "n_links_changed")}, flush=True)

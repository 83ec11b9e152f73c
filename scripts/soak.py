"""Soak run: many sweeps, consistency check (sizes / member lists / K recomputed from the links) at intervals."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from exchanger_b200.sampler import Sampler
from exchanger_b200.synthetic import synth_model

n = int(float(sys.argv[1])); total = int(sys.argv[2]); every = int(sys.argv[3]); kind = sys.argv[4]
model, _ = synth_model(n, bench.clust_prior(kind, n), seed=bench.DATA_SEED)
g = Sampler(model, seed=11, device=0)
t0 = time.time()
for s in range(0, total, every):
    g.sweeps(every - 1)
    g.set_option("check", 1)
    g.sweeps(1)
    g.set_option("check", 0)
    st = g.stats()
    print(f"sweep {s+every}: ok, K={st['n_entities']} wide={st['n_wide']} rounds={st['n_rounds']} last sweep {st['ms_total']:.1f} ms", flush=True)
print(f"{total} sweeps of {n} records ({kind}) in {time.time()-t0:.1f} s, every check passed")

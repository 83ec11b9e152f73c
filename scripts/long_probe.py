"""How the sweep time evolves along the chain: stats every `every` sweeps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from exchanger_b200.sampler import Sampler
from exchanger_b200.synthetic import synth_model

n = int(float(sys.argv[1])); total = int(sys.argv[2]); every = int(sys.argv[3]); kind = sys.argv[4]
model, truth = synth_model(n, bench.clust_prior(kind, n), seed=bench.DATA_SEED)
g = Sampler(model, seed=1, device=0)
for s in range(0, total, every):
    g.sweeps(every - 1)
    ms, nl = g.sweeps_timed(1)
    st = g.stats()
    th = g.lib  # noqa
    state_theta = None
    print(f"sweep {s+every}: {ms:.1f} ms launches {nl} K={st['n_entities']} wide={st['n_wide']} long={st['n_long']} rounds={st['n_rounds']} "
          f"cand={st['n_candidates']/n:.2f} changed={st['n_links_changed']} | index {st['ms_index']:.1f} cand {st['ms_cand']:.1f}+{st['ms_cand_long']:.1f} "
          f"wide {st['ms_wide']:.1f} rounds {st['ms_rounds']:.1f} ent {st['ms_entities']:.1f}", flush=True)
fin = g.state()
print("theta", np.round(fin["distort_probs"], 4), fin["clust"])

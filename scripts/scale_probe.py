"""Ad-hoc probe: synthetic RLdata-shaped model at scale, per-sweep phase times (development aid)."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from exchanger_b200 import PitmanYorRP, GammaRV, BetaRV, GeneralizedCouponRP
from exchanger_b200.synthetic import synth_model
from exchanger_b200.sampler import Sampler

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
nsw = int(sys.argv[2]) if len(sys.argv) > 2 else 30
prior = sys.argv[3] if len(sys.argv) > 3 else "py"
cp = PitmanYorRP(GammaRV(1, 1), BetaRV(1, 1)) if prior == "py" else GeneralizedCouponRP(N, GammaRV(1, 1))
t = time.time(); m, ident = synth_model(N, cp); print(f"model N={N} built in {time.time()-t:.1f}s; domains",
      [len(i.probs) for i in m.attr_indices], "nnz", [None if i.is_constant else len(i.close_val_ids) for i in m.attr_indices], flush=True)
t = time.time(); g = Sampler(m, seed=1); print(f"create {time.time()-t:.2f}s", flush=True)
for kv in sys.argv[4:]:
    k, v = kv.split('='); g.set_option(k, int(v))
for i in range(nsw):
    t0 = time.time(); g.sweeps(1); dt = time.time() - t0
    s = g.stats()
    print(f"sweep {i}: wall {dt*1e3:.1f} ms dev {s['ms_total']:.1f} | clust {s['ms_clust']:.2f} prep {s['ms_prep']:.2f} index {s['ms_index']:.2f}({s['ms_index_count']:.1f}+{s['ms_index_fill']:.1f}) "
          f"cand {s['ms_cand']:.2f}+{s['ms_cand_long']:.2f} wide {s['ms_wide']:.2f} rounds {s['ms_rounds']:.2f}({s['n_rounds']}) long={s['n_long']} canon {s['ms_canon']:.2f} ent {s['ms_entities']:.2f} dist {s['ms_distort']:.2f} | "
          f"K={s['n_entities']} cand/rec={s['n_candidates']/N:.2f} bucket/rec={s['n_bucket_items']/N:.1f} wide={s['n_wide']} tables={s['n_tables']} changed={s['n_links_changed']}", flush=True)
st = g.state()
from exchanger_b200.analysis import pairwise_measures
print("pairwise vs truth (last sample):", pairwise_measures(ident, st["links"]))
print("theta", st["distort_probs"], st["clust"])

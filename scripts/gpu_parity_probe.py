"""Ad-hoc probe: GPU sampler vs oracle, sweep by sweep (development aid, not a test)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from fixtures import readme_model, PRIORS
from checkers import OracleSampler
from exchanger_b200.sampler import Sampler

def compare(a, b, tag):
    bad = []
    for k in ("links", "ent_ids", "ent_attrs", "rec_distortions", "distort_counts"):
        if a[k].shape != b[k].shape or not np.array_equal(a[k], b[k]):
            bad.append(k)
    if not np.array_equal(a["distort_probs"], b["distort_probs"]): bad.append("distort_probs")
    for k in ("alpha", "d", "m", "kappa"):
        if a["clust"][k] != b["clust"][k]: bad.append(k)
    if bad:
        print(tag, "MISMATCH", bad)
        for k in bad:
            if k in a and hasattr(a[k], "shape") and a[k].shape == b[k].shape:
                idx = np.argwhere(a[k] != b[k])
                print("  ", k, "n diff", len(idx), "first", idx[:5].tolist())
            elif k in a["clust"]:
                print("  ", k, a["clust"][k], b["clust"][k])
    return not bad

name = sys.argv[1] if len(sys.argv) > 1 else "rldata500"
prior = sys.argv[2] if len(sys.argv) > 2 else "pitman_yor"
nsw = int(sys.argv[3]) if len(sys.argv) > 3 else 20
opts = dict(kv.split("=") for kv in sys.argv[4:])
n = int(opts.pop("n", 0)) or None
cols_n = None
from fixtures import load_rldata
N = len(load_rldata(name, n)[1])
m, ident = readme_model(name, n=n, clust_prior=PRIORS[prior](N))
g = Sampler(m, seed=11)
for k, v in opts.items(): g.set_option(k, int(v))
g.set_option("check", 1)
o = OracleSampler(m, seed=11)
ok = True
for i in range(nsw):
    t0 = time.time(); g.sweeps(1); tg = time.time() - t0
    t0 = time.time(); o.sweeps(1); to = time.time() - t0
    st = g.stats()
    same = compare(g.state(), o.state(), f"sweep {i}")
    print(f"sweep {i}: same={same} K={st['n_entities']} rounds={st['n_rounds']} wide={st['n_wide']} cand={st['n_candidates']} "
          f"tables={st['n_tables']} changed={st['n_links_changed']} gpu {tg*1e3:.2f} ms (dev {st['ms_total']:.2f}) oracle {to*1e3:.1f} ms", flush=True)
    ok &= same
    if not same: break
print("ALL SAME" if ok else "FAILED")
print(g.stats())

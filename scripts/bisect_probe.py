"""Debug aid: the failing lockstep configuration under option variants."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from checkers import OracleSampler
from fixtures import PRIORS, readme_model
from exchanger_b200.sampler import Sampler
m, _ = readme_model("rldata10000", n=3000, clust_prior=PRIORS["gen_coupon"](3000))
for extra in (dict(), dict(round0=0), dict(wide_buckets=0), dict(emit_cost_cap=0), dict(rounds_tail=0)):
    opts = dict(long_min=1, est_threshold=0); opts.update(extra)
    g, o = Sampler(m, seed=11), OracleSampler(m, seed=11)
    for k, v in opts.items(): g.set_option(k, v)
    for i in range(4):
        g.sweeps(1); o.sweeps(1)
        a, b = g.state(), o.state(); st = g.stats()
        nd = int((a["links"] != b["links"]).sum())
        print(extra, "sweep", i + 1, "links differing", nd, {k: st[k] for k in ("n_wide", "n_listed", "n_scanned", "n_rounds", "n_long", "n_probe", "n_entities")}, "first", np.nonzero(a["links"] != b["links"])[0][:5], flush=True)
        if nd: break
    g.close()

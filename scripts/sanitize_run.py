"""Small chain that takes every path of the sweep, for compute-sanitizer (memcheck / racecheck):
RLdata10000 records with missing values, a few sweeps under each scheduling (the options of
tests/test_gpu_parity.py: multi-bucket probes, warp- and block-staged lists, tiled buckets, chain and
CSR tables, the serial path with lists from buckets or from the scan over all items, rounds launched
one by one or in the persistent kernel), plus the finite-concentration and Dirichlet models, the
point estimates and the neighbourhood builder.
usage: sanitize_run.py [n_records] [sweeps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from exchanger_b200 import BetaRV, DirichletProcess, DirichletRV, GammaRV, GeneralizedCouponRP, PitmanYorRP, exchanger
from exchanger_b200.analysis import DevicePointEstimate
from exchanger_b200.capi import levenshtein_neighbours
from exchanger_b200.sampler import Sampler
from fixtures import load_rldata, readme_attr_params

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cols, _ = load_rldata("rldata10000", n)
rng = np.random.default_rng(5)
for k in ("fname_c1", "lname_c1", "by"):
    cols[k] = [None if rng.random() < 0.15 else v for v in cols[k]]
models = {
    "py": exchanger(cols, readme_attr_params(), PitmanYorRP(GammaRV(1, 1), BetaRV(1, 1))),
    "gc": exchanger(cols, readme_attr_params(), GeneralizedCouponRP(n, GammaRV(1, 1))),
}
ap = readme_attr_params(prior=BetaRV(1, 4))
for k in ap:
    ap[k].distort_dist_prior = DirichletProcess(GammaRV(2, 1e-2))
ap["bm"].entity_dist_prior = DirichletRV(1.0)
models["dp"] = exchanger(cols, ap, PitmanYorRP(GammaRV(1, 1), BetaRV(1, 1)))
runs = [("py", {}), ("gc", {}), ("dp", {}),
        ("py", dict(table_cost_permille=20000, table_min_usage=0, bucket_cap=8)),
        ("py", dict(bucket_cap=0, huge_min=4)),
        ("gc", dict(fast_path=0)), ("gc", dict(hop_cap=1)), ("gc", dict(hop_cap=3, cand_cap=4, long_min=2)),
        ("py", dict(wide_threshold=2)), ("py", dict(wide_threshold=2, wide_buckets=0)),
        ("gc", dict(rounds_tail=0, long_min=2)), ("gc", dict(long_min=0))]
for name, opts in runs:
    with Sampler(models[name], seed=3) as g:
        for k, v in opts.items():
            g.set_option(k, v)
        g.set_option("check", 1)
        g.sweeps(sweeps)
        st = g.stats()
        print(name, opts, "K", st["n_entities"], "rounds", st["n_rounds"], "wide", st["n_wide"], "probe", st["n_probe"],
              "huge", st["n_huge"], "slow", st["n_slow"], flush=True)
with Sampler(models["py"], seed=4) as g, DevicePointEstimate(n, slots=2) as pe:
    hist = g.run(6, 1, 0, point_estimate=pe)
    res = pe.finish(hist["links"])
    print("point estimate: overflow", res["n_overflow"], "clusters", len(set(res["membership"].tolist())), flush=True)
dom = sorted({v for v in cols["lname_c1"] if v is not None})
rp, ci, pr, mef = levenshtein_neighbours(dom, 3)
print("neighbours", len(dom), len(ci), flush=True)
print("SANITIZE_RUN_DONE")

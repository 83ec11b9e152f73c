"""Parity at a size the small fixtures do not reach: synthetic records, CUDA sampler against the
oracle after every sweep (integer state identical), consistency check on."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench
from checkers import OracleSampler
from exchanger_b200.sampler import Sampler
from exchanger_b200.synthetic import synth_model

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200_000
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
for kind in sys.argv[3:] or ["gen_coupon", "pitman_yor"]:
    model, _ = synth_model(n, bench.clust_prior(kind, n), seed=7)
    gpu, oracle = Sampler(model, seed=5), OracleSampler(model, seed=5)
    gpu.set_option("check", 1)
    for s in range(sweeps):
        t0 = time.time(); gpu.sweeps(1); t1 = time.time(); oracle.sweeps(1); t2 = time.time()
        a, b = gpu.state(), oracle.state()
        st = gpu.stats()
        same = all(np.array_equal(a[k], b[k]) for k in ("links", "ent_ids", "ent_attrs", "rec_distortions", "distort_counts")) \
            and np.array_equal(a["distort_probs"], b["distort_probs"]) and a["clust"] == b["clust"]
        print(f"{kind} N={n} sweep {s+1}: identical={same} gpu {1e3*(t1-t0):.1f} ms oracle {t2-t1:.1f} s | K={st['n_entities']} "
              f"wide={st['n_wide']} long={st['n_long']} probe={st['n_probe']} huge={st['n_huge']} tables={st['n_tables']} rounds={st['n_rounds']}", flush=True)
        if not same:
            d = np.nonzero(a["links"] != b["links"])[0]
            print("  links differ at", d[:10])
            sys.exit(1)
    gpu.close()
print("parity at scale: ok")

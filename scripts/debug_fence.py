import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from checkers import OracleSampler
from exchanger_b200 import BetaRV, PitmanYorRP, GammaRV, exchanger
from exchanger_b200.sampler import Sampler
from fixtures import load_rldata, readme_attr_params

cols, _ = load_rldata("rldata500", 200)
rng = np.random.default_rng(1)
for k in ("fname_c1", "lname_c1", "by", "bm", "bd"):
    cols[k] = [None if rng.random() < 0.45 else v for v in cols[k]]
m = exchanger(cols, readme_attr_params(), PitmanYorRP(GammaRV(1, 1), BetaRV(1, 1)))
gpu, oracle = Sampler(m, seed=11), OracleSampler(m, seed=11)
gpu.set_option("check", 1); gpu.set_option("wide_threshold", 150); oracle.set_wide_threshold(150)
gpu.set_option("trace", 1)
for it in range(6):
    gpu.sweeps(1); oracle.sweeps(1)
    a, b = gpu.state(), oracle.state()
    st = gpu.stats()
    print("sweep", it + 1, "wide", st["n_wide"], "long", st["n_long"], "rounds", st["n_rounds"], "oracle n_wide", oracle.n_wide() if hasattr(oracle, "n_wide") else "?")
    d = np.nonzero(a["links"] != b["links"])[0]
    if len(d):
        print("links differ at records", d[:20], "gpu", a["links"][d[:20]], "oracle", b["links"][d[:20]])
        for r in d[:5]:
            print(" rec", r, "x", m.rec_attrs[:, r], "z gpu", a["rec_distortions"][:, r], "z oracle", b["rec_distortions"][:, r])
        break

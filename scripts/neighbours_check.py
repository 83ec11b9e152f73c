import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from exchanger_b200.capi import levenshtein_neighbours
from exchanger_b200.model import compute_close_values, transform_dist_fn, Levenshtein
from exchanger_b200.synthetic import synth_rldata
codes, doms, names, ident = synth_rldata(40000)
for a in (3, 4):
    dom = doms[a]
    for inc in (False, True):
        t = time.time(); rp, ci, pr, mef = levenshtein_neighbours(dom, 3, include_self=inc); tg = time.time() - t
        t = time.time(); rp2, ci2, pr2, mef2 = compute_close_values(dom, transform_dist_fn(Levenshtein(), 3), inc); tc = time.time() - t
        ok = np.array_equal(rp, rp2) and np.array_equal(ci, ci2) and np.allclose(pr, pr2, rtol=1e-12) and np.allclose(mef, mef2, rtol=1e-12, equal_nan=True)
        print(names[a], "D", len(dom), "include_self", inc, "nnz", len(ci), "gpu %.3fs py %.1fs" % (tg, tc), "MATCH" if ok else "MISMATCH")

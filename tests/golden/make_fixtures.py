"""Regenerates tests/golden/*.npz from the reference's datasets.

Run in the build container (needs /root/reference):  python tests/golden/make_fixtures.py
The .npz files hold the RLdata500 / RLdata10000 tables of the reference
(data/RLdata500.rda, data/RLdata10000.rda; documented in R/data.R:1-60) as plain string
columns plus the ground-truth identity vectors, so that tests and the benchmark can run
where /root/reference does not exist (the GPU box).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from exchanger_b200.rdata import data_frame_columns, read_rda  # noqa: E402

REF = os.environ.get("EXCHANGER_REFERENCE", "/root/reference")

for name in ("RLdata500", "RLdata10000"):
    d = read_rda(os.path.join(REF, "data", name + ".rda"))
    cols = data_frame_columns(d[name])
    out = {}
    for k, v in cols.items():
        out[k] = np.array(["" if x is None else (str(int(x)) if isinstance(x, float) else str(x)) for x in v])
        out[k + "__na"] = np.array([x is None for x in v])
    out["identity"] = np.array(d["identity." + name].value, dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, name.lower() + ".npz"), **out)
    print(name, {k: len(set(v)) for k, v in cols.items()})

import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from exchanger_b200.synthetic import synth_model
n = int(float(sys.argv[1]))
model, _ = synth_model(n, bench.clust_prior("gc", n), seed=bench.DATA_SEED)
for a, idx in enumerate(model.attr_indices):
    if idx.is_constant: continue
    rp = np.asarray(idx.close_row_ptr); L = np.diff(rp)
    x = model.rec_attrs[a]; x = x[x >= 0]
    Lr = L[x]  # row length seen by records (entity-value orientation, close to the inverted one)
    print(model.attr_names[a], "D", len(L), "nnz", rp[-1], "mean row", L.mean(), "max", L.max(), "pcts", np.percentile(L, [50, 90, 99]),
          "| weighted by records: mean", Lr.mean(), "pcts", np.percentile(Lr, [50, 90, 99]))

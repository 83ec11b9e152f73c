"""Print the dependency rounds of one sweep (EXB200_TRACE) after a burn-in."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from exchanger_b200.sampler import Sampler
from exchanger_b200.synthetic import synth_model

n = int(float(sys.argv[1])); burn = int(sys.argv[2]); kind = sys.argv[3]
model, _ = synth_model(n, bench.clust_prior(kind, n), seed=bench.DATA_SEED)
g = Sampler(model, seed=1, device=0)
g.sweeps(burn)
sys.stderr.write("=== traced sweep\n"); sys.stderr.flush()
import ctypes
g.lib.exb200_set_option(g.h, b"trace", 1)
g.sweeps(1)
st = g.stats()
print({k: round(v, 2) if isinstance(v, float) else v for k, v in st.items()})

"""Break the end-to-end call sequence (create -> run -> state -> destroy) into its stages."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from exchanger_b200.model import model_with_state
from exchanger_b200.sampler import Sampler
from exchanger_b200.synthetic import synth_model

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
kind = sys.argv[3] if len(sys.argv) > 3 else "gc"
model, _ = synth_model(n, bench.clust_prior(kind, n), seed=bench.DATA_SEED)
g = Sampler(model, seed=1, device=0)
g.sweeps(25)
hm = model_with_state(model, g.state())
g.close()
torch.cuda.synchronize()
g3 = Sampler(hm, seed=2, device=0)
t0 = time.perf_counter()
ms, nl = g3.sweeps_timed(steps)
print(f"same {steps} sweeps without history: device {ms/steps:.1f} ms/sweep, wall {(time.perf_counter()-t0)/steps*1e3:.1f} ms/sweep, {nl/steps:.0f} launches/sweep")
st = g3.stats()
print("last sweep phases:", {k: round(v, 2) for k, v in st.items() if k.startswith("ms_")}, "rounds", st["n_rounds"], "wide", st["n_wide"])
g3.close()
for rep in range(2):
    t0 = time.perf_counter()
    g2 = Sampler(hm, seed=2, device=0)
    t1 = time.perf_counter()
    hist = g2.run(steps, 1, 0)
    t2 = time.perf_counter()
    fin = g2.state()
    t3 = time.perf_counter()
    g2.close()
    t4 = time.perf_counter()
    print(f"rep {rep}: create {t1-t0:.3f}s (flat {getattr(g2,'t_flat',float('nan')):.3f}) run {t2-t1:.3f}s "
          f"({(t2-t1)/steps*1e3:.1f} ms/sweep) state {t3-t2:.3f}s close {t4-t3:.3f}s total {t4-t0:.3f}s")

# the same call sequence with the point estimate accumulated on the device instead of a link history
from exchanger_b200.analysis import DevicePointEstimate
t0 = time.perf_counter()
g4 = Sampler(hm, seed=2, device=0)
pe = DevicePointEstimate(n)
t1 = time.perf_counter()
hist = g4.run(steps, 1, 0, keep_links=False, point_estimate=pe)
t2 = time.perf_counter()
res = pe.finish()
t3 = time.perf_counter()
pe.close(); g4.close()
print(f"device point estimate instead of the link history: create {t1-t0:.3f}s run {t2-t1:.3f}s ({(t2-t1)/steps*1e3:.1f} ms/sweep) "
      f"finish {t3-t2:.3f}s; {len(set(res['membership'].tolist()))} shared most probable clusters, overflow {res['n_overflow']}")

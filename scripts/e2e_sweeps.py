"""Wall time of every saved sweep of a fresh handle (where does exb200_run spend more than the device time?)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from exchanger_b200.model import model_with_state
from exchanger_b200.sampler import Sampler
from exchanger_b200.synthetic import synth_model
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
model, _ = synth_model(n, bench.clust_prior("gc", n), seed=bench.DATA_SEED)
g = Sampler(model, seed=1, device=0); g.sweeps(25); hm = model_with_state(model, g.state()); g.close()
for rep in range(2):
    t0 = time.perf_counter(); g2 = Sampler(hm, seed=2, device=0); t1 = time.perf_counter()
    walls, devs = [], []
    for i in range(8):
        t = time.perf_counter(); h = g2.run(1, 1, 0); walls.append((time.perf_counter() - t) * 1e3); devs.append(g2.stats()["ms_total"])
    t = time.perf_counter(); h = g2.run(12, 1, 0); w12 = (time.perf_counter() - t) * 1e3
    g2.close()
    print(f"rep {rep}: create {t1-t0:.3f}s; run(1) wall ms {[round(w,1) for w in walls]} device ms {[round(d,1) for d in devs]}; run(12) {w12/12:.1f} ms/sweep", flush=True)

"""Per-record cost of the reference's own State::update() (oracle/_ref) as the record count grows,
on one host core: the synthetic generator of bench.py at several sizes, each chain started from the
state after 20 sweeps of the oracle, one warm-up and `sweeps` timed sweeps.
usage: ref_scaling.py size [size ...]   (prints one line per size)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from checkers import OracleSampler, RefSampler, cpu_neighbours
from exchanger_b200.model import model_with_state
from exchanger_b200.synthetic import synth_model

for arg in sys.argv[1:]:
    n = int(float(arg))
    t0 = time.perf_counter()
    model, _ = synth_model(n, bench.clust_prior("gen_coupon", n), seed=bench.DATA_SEED,
                           neighbour_fn=lambda dom, thr: cpu_neighbours(dom, thr, False))
    t1 = time.perf_counter()
    o = OracleSampler(model, seed=1); o.sweeps(20)
    t2 = time.perf_counter()
    ref = RefSampler(model_with_state(model, o.state()), seed=1); del o
    ref.update(1)
    sweeps = 2 if n >= 500_000 else 4
    t3 = time.perf_counter()
    secs = ref.update_timed(sweeps)
    t4 = time.perf_counter()
    per = (t4 - t3) / sweeps / n * 1e6
    print(f"records {n}: model {t1-t0:.1f}s, oracle 20 sweeps {t2-t1:.1f}s, reference {(t4-t3)/sweeps:.2f} s/sweep = {per:.1f} us/record "
          f"(links {secs[1]/(t4-t3)*100:.0f}% entities {secs[2]/(t4-t3)*100:.0f}%), K {ref.n_entities()}", flush=True)

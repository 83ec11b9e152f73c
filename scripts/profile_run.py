"""Short steady-state run for ncu: builds the synthetic model, burns in, runs a few sweeps."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from exchanger_b200 import PitmanYorRP, GammaRV, BetaRV, GeneralizedCouponRP
from exchanger_b200.synthetic import synth_model
from exchanger_b200.sampler import Sampler
N = int(float(sys.argv[1])); burn = int(sys.argv[2]); n = int(sys.argv[3]); prior = sys.argv[4] if len(sys.argv) > 4 else "gc"
cp = PitmanYorRP(GammaRV(1, 1), BetaRV(1, 1)) if prior == "py" else GeneralizedCouponRP(N, GammaRV(1, 1))
m, _ = synth_model(N, cp)
g = Sampler(m, seed=1)
g.sweeps(burn)
import torch; torch.cuda.synchronize()
print("PROFILE_START launches so far unknown; sweeps", n, flush=True)
ms, nl = g.sweeps_timed(n)
print("ms", ms, "launches", nl, g.stats(), flush=True)

"""Where exb200_create's time goes when several ranks create handles at once (run under torchrun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
from exchanger_b200.capi import FlatModel
from exchanger_b200.sampler import Sampler
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
def barrier():
    torch.cuda.synchronize()
    if world > 1: dist.barrier(); torch.cuda.synchronize()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
model, _ = bench.shared_model(n, "gen_coupon", "beta150", rank, world, local, barrier)
for rep in range(3):
    barrier(); t0 = time.perf_counter()
    f = FlatModel(model, 1)
    t1 = time.perf_counter()
    s = Sampler(model, seed=1, device=local)
    t2 = time.perf_counter()
    s.close(); barrier()
    t3 = time.perf_counter()
    print(f"rank {rank} rep {rep}: FlatModel {t1-t0:.3f}s  Sampler (FlatModel + exb200_create) {t2-t1:.3f}s (flat {getattr(s, 't_flat', float('nan')):.3f})  close {t3-t2:.3f}s  cores {os.cpu_count()}", flush=True)

"""Within-chain sharding check (run under torchrun, one rank per GPU): the sharded chain must equal
the single-GPU chain state for state; prints per-sweep times of both."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist
import bench
from exchanger_b200.sampler import Sampler
from exchanger_b200.sharded import ShardedSampler
from exchanger_b200.synthetic import synth_model

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29511")
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
kind = sys.argv[3] if len(sys.argv) > 3 else "gen_coupon"
model, _ = synth_model(n, bench.clust_prior(kind, n), seed=bench.DATA_SEED)
sh = ShardedSampler(model, seed=3, device=local)
ref = Sampler(model, seed=3, device=local) if rank == 0 else None
ok = True
for s in range(sweeps):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter(); sh.sweeps(1); torch.cuda.synchronize(); dist.barrier(); t1 = time.perf_counter()
    if rank == 0:
        t2 = time.perf_counter(); ref.sweeps(1); t3 = time.perf_counter()
        a, b = sh.state(), ref.state()
        same = all(np.array_equal(a[k], b[k]) for k in ("links", "ent_ids", "ent_attrs", "rec_distortions", "distort_counts")) \
            and np.array_equal(a["distort_probs"], b["distort_probs"]) and a["clust"] == b["clust"]
        ok &= same
        print(f"sweep {s+1}: sharded over {world} GPU(s) {1e3*(t1-t0):.1f} ms, single GPU {1e3*(t3-t2):.1f} ms, identical={same}", flush=True)
# every rank holds the same chain
st = sh.state()
sig = torch.tensor([int(st["links"].astype(np.int64).sum() % (1 << 40)), int(st["rec_distortions"].sum()), int(st["ent_attrs"].astype(np.int64).sum() % (1 << 40))],
                   device="cuda", dtype=torch.int64)
lo, hi = sig.clone(), sig.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
if rank == 0:
    print("replicas agree:", bool((lo == hi).all().item()), "| sharded == single GPU:", ok)
dist.destroy_process_group()
sys.exit(0 if ok and bool((lo == hi).all().item()) else 1)

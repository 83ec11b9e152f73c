#!/usr/bin/env python
"""Benchmark of the Gibbs sweep (BASELINE.json metric: Gibbs sweeps/s and record-link updates/s on
synthetic RLdata-shaped records, 1-8 B200).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA sampler
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own C++ on host cores

A step is one application of State::update() (src/state.h:60-75) to the whole record set.
Prints ONE JSON line (rank 0).  With N > 1 GPUs every rank runs an independent chain on the same
records (BASELINE config 5; no data-path collective) and `value` is the aggregate over chains.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (records, cluster prior, description)
    "c3": (1_000_000, "pitman_yor", "synthetic 1M RLdata-shaped records, 5 attributes, Pitman-Yor prior"),
    "c4": (10_000_000, "gen_coupon", "synthetic 10M RLdata-shaped records, 5 attributes, GeneralizedCoupon RP"),
}
A = 5
DATA_SEED = 20260101
# The reference's sweep costs 40-60 us per record on one host core, growing with the record count
# (its containers are hash maps; measured on the B200 box's host: profiles/README.md, "reference
# scaling"), so a whole 10 M sweep is about ten minutes per core.  Its arm therefore times a SAMPLE of
# the same generator, as large as the wall budget allows, started from a burnt-in state.
REF_US_PER_RECORD = 50.0          # planning figure for the sample size only
REF_BUDGET_S = 150.0              # timed + warm-up sweeps of the reference arm
CPU_BASELINE_BUDGET_S = 20.0      # the one-core baseline inside the normal bench line
REF_BURNIN = 20                   # sweeps (by the oracle) before the reference is timed: SURVEY.md 8d window


def distort_prior(name: str):
    from exchanger_b200 import BetaRV
    return BetaRV(1, 1) if name == "beta11" else BetaRV(1, 50)


def ref_sample_records(n_workload: int, sweeps: int, budget_s: float) -> int:
    n = int(budget_s / (max(sweeps, 1) * REF_US_PER_RECORD * 1e-6))
    return int(max(20_000, min(n_workload, 1_000_000, (n // 10_000) * 10_000)))


def clust_prior(kind: str, n: int):
    from exchanger_b200 import BetaRV, GammaRV, GeneralizedCouponRP, PitmanYorRP
    if kind == "pitman_yor":
        return PitmanYorRP(GammaRV(1, 1), BetaRV(1, 1))
    return GeneralizedCouponRP(n, GammaRV(1, 1))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, smax, reasons = [], None, set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                smax = float(r[2])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- reference arm
def _ref_worker(args):
    """One chain of the reference's own State::update() (oracle/_ref) on the bounded sample: the model
    is prepared on the host (neighbourhoods by oracle/exnbr.c), brought to the window SURVEY.md 8d
    prescribes by REF_BURNIN sweeps of the oracle (whose chain has the reference's law and costs a
    fraction of its time), handed to the reference as its initial state, then timed."""
    n_rec, kind, seed, warm, steps, prior, threads = args
    from checkers import OracleSampler, RefSampler, cpu_neighbours
    from exchanger_b200.model import model_with_state
    from exchanger_b200.synthetic import synth_model

    def nbr(dom, thr):
        return cpu_neighbours(dom, thr, False, threads)

    model, _ = synth_model(n_rec, clust_prior(kind, n_rec), seed=DATA_SEED, neighbour_fn=nbr,
                           distort_prior=distort_prior(prior))
    o = OracleSampler(model, seed=seed)
    o.sweeps(REF_BURNIN)
    ref = RefSampler(model_with_state(model, o.state()), seed=seed)
    del o
    ref.update(warm)
    t0 = time.perf_counter()
    ref.update(steps)
    return time.perf_counter() - t0


def reference_throughput(kind: str, n_workload: int, steps: int, warmup: int, procs: int, budget_s: float,
                         prior: str = "beta150"):
    """sweeps/s of the reference CPU sampler, expressed at the workload size: the sample's
    link-updates/s divided by the workload's record count (optimistic for the reference, whose
    per-record cost grows with N).  Returns (sweeps/s, link-updates/s, seconds, sample records)."""
    import multiprocessing as mp
    n_sample = ref_sample_records(n_workload, steps + warmup, budget_s)
    threads = max(1, (os.cpu_count() or 1) // procs)
    jobs = [(n_sample, kind, 100 + i, warmup, steps, prior, threads) for i in range(procs)]
    if procs == 1:
        secs = [_ref_worker(jobs[0])]
    else:
        with mp.get_context("spawn").Pool(procs) as pool:
            secs = pool.map(_ref_worker, jobs)
    t = max(secs)
    updates_per_s = procs * steps * n_sample / t
    return updates_per_s / n_workload, updates_per_s, t, n_sample


def sample_text(procs: int, steps: int, n_sample: int) -> str:
    return (f"{procs} independent chain(s) of the reference's own State::update() (oracle/_ref), one per host core, each on a "
            f"{n_sample}-record sample of the same synthetic generator, started from the state after {REF_BURNIN} sweeps, "
            f"{steps} timed sweeps each; sweeps/s = link-updates/s / workload records (the reference is single-threaded "
            "and its per-record cost grows with the record count, so this flatters it)")


def run_reference(opt, n_records, kind, desc):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import __graft_entry__ as entry
    entry.build()   # builds the checkers (oracle/) only where they are stale; the CUDA library is not used here
    procs = max(1, min(os.cpu_count() or 1, 32))
    v, ups, t, n_sample = reference_throughput(kind, n_records, opt.steps, opt.warmup, procs, REF_BUDGET_S, opt.distort_prior)
    line = {
        "impl": "reference", "metric": "gibbs_sweeps_per_s", "value": v, "unit": "sweeps/s",
        "link_updates_per_s": ups, "one_core_link_updates_per_s": ups / procs,
        "n_gpus": opt.gpus, "steps": opt.steps, "warmup": opt.warmup,
        "ms_per_step": 1e3 * t / opt.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "records": n_records, "attributes": A, "clust_prior": kind,
                   "distort_prior": "Beta(1,1)" if opt.distort_prior == "beta11" else "Beta(1,50)",
                   "sample_records_per_chain": n_sample},
        "cpu_baseline": {"value": v, "unit": "sweeps/s", "cores": procs, "kind": "reference",
                         "sample": sample_text(procs, opt.steps, n_sample)},
        "e2e": {"value": v, "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit_line(line)


# ----------------------------------------------------------------------------- this repo's arm
def shared_model(n_records, kind, prior, rank, world, local, barrier):
    """The synthetic model, built ONCE per node: rank 0 generates the records and computes the value
    neighbourhoods on ITS OWN GPU, the other ranks load the pickled model from /dev/shm."""
    import pickle
    from exchanger_b200.capi import levenshtein_neighbours
    from exchanger_b200.synthetic import synth_model
    path = f"/dev/shm/exb200_model_{n_records}_{kind}_{prior}_{DATA_SEED}_{os.environ.get('MASTER_PORT', '0')}.pkl"
    t0 = time.perf_counter()
    model = None
    if rank == 0:
        model, _ = synth_model(n_records, clust_prior(kind, n_records), seed=DATA_SEED, distort_prior=distort_prior(prior),
                               neighbour_fn=lambda dom, thr: levenshtein_neighbours(dom, thr, device=local))
        if world > 1:
            with open(path + ".tmp", "wb") as f:
                pickle.dump(model, f, protocol=pickle.HIGHEST_PROTOCOL)
            os.replace(path + ".tmp", path)
    if world > 1:
        barrier()
        if rank != 0:
            with open(path, "rb") as f:
                model = pickle.load(f)
        barrier()
        if rank == 0:
            os.unlink(path)
    return model, time.perf_counter() - t0


def run_ours(opt, n_records, kind, desc):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import __graft_entry__ as entry
    if rank == 0:
        entry.build()
    if world > 1:
        dist.barrier()
    from exchanger_b200.model import model_with_state
    from exchanger_b200.sampler import Sampler

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    model, t_build = shared_model(n_records, kind, opt.distort_prior, rank, world, local, barrier)
    gpu = Sampler(model, seed=1000 + rank, device=local)
    phase_keys = ("ms_clust", "ms_prep", "ms_index", "ms_index_count", "ms_index_fill", "ms_cand", "ms_cand_fast", "ms_cand_long",
                  "ms_wide", "ms_rounds", "ms_tail", "ms_canon", "ms_entities", "ms_entities_single", "ms_distort", "ms_total")

    def timed(n, acc=None):
        """n sweeps one by one: (device ms, launches), per-phase sums into acc"""
        ms_sum, nl_sum = 0.0, 0
        for _ in range(n):
            ms, nl = gpu.sweeps_timed(1)
            ms_sum += ms
            nl_sum += nl
            if acc is not None:
                st = gpu.stats()
                for k in phase_keys:
                    acc[k] += st[k]
                for k in ("n_candidates", "n_rounds", "n_wide", "n_tables", "n_slow", "n_probe", "n_long", "n_listed", "n_scanned"):
                    acc[k] = acc.get(k, 0) + st[k]
        return ms_sum, nl_sum

    gpu.sweeps(opt.burnin)          # untimed: reach the window SURVEY.md 8d prescribes (sweeps 21-120)
    ms_warm, _ = timed(opt.warmup)  # W warm-up steps (sweeps 21 .. 20+W; they count for the long window below)
    barrier()
    clocks = ClockSampler(local)
    acc = {k: 0.0 for k in phase_keys}
    t0 = time.perf_counter()
    dev_ms, launches = timed(opt.steps, acc)      # EXACTLY K steps
    barrier()
    wall = time.perf_counter() - t0
    clk = clocks.stop()
    k_ent = gpu.stats()["n_entities"]
    secs = torch.tensor([max(wall, dev_ms / 1e3)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(secs, op=dist.ReduceOp.MAX)
    t_max = float(secs.item())
    value = world * opt.steps / t_max
    # the end-to-end runs below start from the chain as it is NOW (the state after the timed steps), so
    # that they time the same sweeps of the chain as `value` does and not the later, slower ones
    host_model = model_with_state(model, gpu.state())

    # ---- end to end through the public API with host buffers (create -> run -> state -> destroy);
    # five repetitions, the median is reported (single shots are noisy: exb200_create alone was seen to take
    # 0.16 s or 0.47 s on the same box within one run - host-side page faults, frees, neighbours on the host).
    # They run while the first chain is still open and before it goes on through the long window: after
    # 200 more sweeps that handle holds a dozen more tables, and creating / destroying handles next to its
    # release measured three to five times slower (0.86 s against 0.17 s for exb200_create)
    import gc
    e2e_runs, e2e_all_stages = [], []
    for _ in range(5):
        gc.collect()
        barrier()
        te = time.perf_counter()
        with Sampler(host_model, seed=2000 + rank, device=local) as g2:
            t_create = time.perf_counter() - te
            h2d = g2.flat.nbytes()
            hist = g2.run(opt.steps, 1, 0)     # every sweep saved: links row copied back each step
            t_run = time.perf_counter() - te - t_create
            fin = g2.state()
            t_state = time.perf_counter() - te - t_create - t_run
        barrier()
        te = time.perf_counter() - te
        e2e_stages = {"create": round(t_create, 3), "flatten": round(getattr(g2, "t_flat", 0.0), 3), "run": round(t_run, 3),
                      "get_state": round(t_state, 3), "destroy": round(te - t_create - t_run - t_state, 3)}
        e2e_all_stages.append(e2e_stages)
        d2h = sum(v.nbytes for v in hist.values() if hasattr(v, "nbytes")) + \
            sum(v.nbytes for v in fin.values() if hasattr(v, "nbytes"))
        del hist, fin
        es = torch.tensor([te], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(es, op=dist.ReduceOp.MAX)
        e2e_runs.append(world * opt.steps / float(es.item()))
    e2e_value = sorted(e2e_runs)[len(e2e_runs) // 2]

    # ---- the rest of the chain, device-timed on rank 0's clock: the whole window of SURVEY.md 8d
    # (sweeps 21-120) and a late figure (sweeps 201-220: the sweep gets slower as more records carry
    # distorted names)
    windows = None
    if opt.long_window and opt.burnin == 20 and opt.warmup + opt.steps <= 100:
        ms_rest, _ = timed(100 - opt.warmup - opt.steps)
        ms_21_120 = ms_warm + dev_ms + ms_rest
        gpu.sweeps(80)
        ms_late, _ = timed(20)
        windows = {"sweeps_21_120": {"value": 100.0 / (ms_21_120 * 1e-3), "unit": "sweeps/s", "ms_per_sweep": ms_21_120 / 100.0},
                   "sweeps_201_220": {"value": 20.0 / (ms_late * 1e-3), "unit": "sweeps/s", "ms_per_sweep": ms_late / 20.0},
                   "what": "one chain per GPU, device time of this rank's chain (CUDA events on the library's stream)"}
    gpu.close()

    if rank == 0:
        # ---- roofline of the dominant single kernel (algorithmic bytes: DESIGN.md section 5)
        n, cbar, rho = n_records, acc["n_candidates"] / opt.steps / n_records, k_ent / n_records
        tbar = acc["n_tables"] / opt.steps  # blocking tables built per sweep
        fast = acc["ms_cand_fast"] > 0
        # Single kernels with a CUDA-event time of their own (one launch per sweep each; the index build,
        # the serial path and the rest of the entity update are multi-kernel phases).  Algorithmic bytes:
        # DESIGN.md section 5 - the links figure of SURVEY 8d, 5A + 8 + c(4A + 8), is split between the
        # candidate kernel (x, z, candidate tuples) and the rounds (lambda, candidate sizes).
        kernels = {
            "k_rounds_tail (links: dependency rounds, persistent)": (acc["ms_tail"], n * (4 + 4 * cbar)),
            ("k_cand_fast (links: candidate generation, first pass)" if fast else "k_cand (links: candidate generation)"):
                (acc["ms_cand_fast"] if fast else acc["ms_cand"], n * (5 * A + 4 + cbar * (4 * A + 4))),
            "k_distort (distortion indicators)": (acc["ms_distort"], n * (9 * A + 4)),
            "k_prep (links: new-entity attributes)": (acc["ms_prep"], n * (5 * A + 4 + 4 * A + 8)),
        }
        kernels = {k: v for k, v in kernels.items() if v[0] > 0}
        name, (ms_k, bytes_k) = max(kernels.items(), key=lambda kv: kv[1][0])  # the dominant one: largest share of the step
        peak, peak_src = peaks()
        achieved = bytes_k / (ms_k / opt.steps * 1e-3) / 1e9
        per_kernel = {k.split(" ")[0]: {"ms": v[0] / opt.steps, "algorithmic_bytes": v[1],
                                        "achieved": v[1] / (v[0] / opt.steps * 1e-3) / 1e9,
                                        "frac": v[1] / (v[0] / opt.steps * 1e-3) / 1e9 / peak} for k, v in kernels.items()}
        sweep_bytes = n * ((5 * A + 8 + cbar * (4 * A + 8)) + (4 + 4 * A + 4 * A * rho) + (9 * A + 4))
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(name.split(" ")[0])
        # ---- CPU baseline: the reference's own sampler, one core, bounded sample
        cpu = None
        if world == 1 and not opt.no_cpu_baseline:
            try:
                v, ups, _, n_sample = reference_throughput(kind, n_records, 2, 1, 1, CPU_BASELINE_BUDGET_S, opt.distort_prior)
                cpu = {"value": v, "unit": "sweeps/s", "cores": 1, "kind": "reference",
                       "link_updates_per_s": ups, "sample": sample_text(1, 2, n_sample)}
            except Exception as e:  # the checker is missing: say so, do not fake a number
                cpu = {"value": None, "unit": "sweeps/s", "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}
        line = {
            "metric": "gibbs_sweeps_per_s", "value": value, "unit": "sweeps/s",
            "link_updates_per_s": value * n_records,
            "n_gpus": world, "steps": opt.steps, "warmup": opt.warmup, "ms_per_step": 1e3 * t_max / opt.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "records": n_records, "attributes": A, "clust_prior": kind,
                       "distort_prior": "Beta(1,1) (README model)" if opt.distort_prior == "beta11" else "Beta(1,50)",
                       "name_frequencies": "Zipf exponent 0.5 (synthetic.py)", "parallelism": f"{world} independent chain(s)",
                       "burnin_sweeps": opt.burnin, "timed_sweeps": f"{opt.burnin + opt.warmup + 1}-{opt.burnin + opt.warmup + opt.steps}",
                       "model_build_s": round(t_build, 1),
                       "l2": f"inputs larger than L2: resident state {n_records * 150 / 1e6:.0f} MB streamed every sweep"},
            "clocks": clk, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "sweeps/s", "h2d_bytes_per_step": int(h2d / opt.steps),
                    "d2h_bytes_per_step": int(d2h / opt.steps), "runs": e2e_runs, "stages_s_last_run": e2e_stages, "stages_s": e2e_all_stages,
                    "what": "exb200_create + exb200_run(steps saved sweeps) + exb200_get_state + exb200_destroy, host buffers; median of 5"},
            "roofline": {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bytes_k, "kernel_ms": ms_k / opt.steps,
                         "kernels": per_kernel,
                         "sweep": {"algorithmic_bytes": sweep_bytes, "ms": acc["ms_total"] / opt.steps,
                                   "achieved": sweep_bytes / (acc["ms_total"] / opt.steps * 1e-3) / 1e9,
                                   "frac": sweep_bytes / (acc["ms_total"] / opt.steps * 1e-3) / 1e9 / peak}},
            "cpu_baseline": cpu,
            "windows": windows,
            "phases_ms": {k[3:]: acc[k] / opt.steps for k in phase_keys},
            "sweep_stats": {"candidates_per_record": cbar, "entities_per_record": rho,
                            "rounds_per_sweep": acc["n_rounds"] / opt.steps, "wide_records_per_sweep": acc["n_wide"] / opt.steps,
                            "tables_per_sweep": tbar, "records_left_to_bucket_kernels": acc["n_slow"] / opt.steps,
                            "records_listed_through_tables_per_sweep": acc["n_listed"] / opt.steps,
                            "records_scanned_against_all_items_per_sweep": acc["n_scanned"] / opt.steps,
                            "launches_per_sweep": launches / opt.steps},
        }
        emit_line(line)
    if world > 1:
        dist.destroy_process_group()


def run_sharded(opt, n_records, kind, desc):
    """--mode shard: ONE chain over all the GPUs (within-chain sharding) -
    strong scaling.  The default mode (independent chains, weak scaling) is what the driver runs."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29512")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    import __graft_entry__ as entry
    if rank == 0:
        entry.build()
    dist.barrier()
    from exchanger_b200.capi import levenshtein_neighbours
    from exchanger_b200.sharded import ShardedSampler
    from exchanger_b200.synthetic import synth_model
    model, _ = synth_model(n_records, clust_prior(kind, n_records), seed=DATA_SEED, distort_prior=distort_prior(opt.distort_prior),
                           neighbour_fn=lambda dom, thr: levenshtein_neighbours(dom, thr, device=local))
    g = ShardedSampler(model, seed=1000, device=local)
    g.sweeps(opt.burnin + opt.warmup)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    clocks = ClockSampler(local)
    t0 = time.perf_counter()
    g.sweeps(opt.steps)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    secs = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    dist.all_reduce(secs, op=dist.ReduceOp.MAX)
    clk = clocks.stop()
    if rank == 0:
        t = float(secs.item())
        emit_line({
            "metric": "gibbs_sweeps_per_s", "value": opt.steps / t, "unit": "sweeps/s",
            "link_updates_per_s": opt.steps * n_records / t, "n_gpus": world, "steps": opt.steps, "warmup": opt.warmup,
            "ms_per_step": 1e3 * t / opt.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "records": n_records, "attributes": A, "clust_prior": kind,
                       "parallelism": f"one chain sharded over {world} GPU(s): links replicated, entity attributes and "
                                      "distortion indicators on slices, all-gather + all-reduce per sweep",
                       "burnin_sweeps": opt.burnin},
            "clocks": clk, "timing": "wall clock around the timed sweeps, max over ranks (collectives included)",
        })
    dist.destroy_process_group()


_JSON_FD = None  # the process's real stdout, kept for the one JSON line


def emit_line(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    # stdout carries exactly one JSON line.  Libraries write there too (NCCL prints its version banner and
    # its NCCL_DEBUG output to stdout from C), so file descriptor 1 is pointed at stderr for the run and
    # the JSON line goes to the descriptor that was stdout.
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("EXB200_WORKLOAD", "c4"), choices=list(WORKLOADS))
    ap.add_argument("--records", type=int, default=0, help="override the record count (debugging)")
    ap.add_argument("--burnin", type=int, default=20, help="untimed sweeps before the warm-up")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--distort-prior", default="beta150", choices=["beta150", "beta11"],
                    help="prior of the distortion probabilities: Beta(1,50) (default) or the README model's Beta(1,1)")
    ap.add_argument("--no-long-window", dest="long_window", action="store_false",
                    help="skip the extra sweeps 21-120 / 201-220 figures")
    ap.add_argument("--mode", default="chains", choices=["chains", "shard"],
                    help="N > 1: independent chains (default, weak scaling) or one chain sharded over the GPUs")
    opt = ap.parse_args()
    opt.warmup = max(opt.warmup, 3)
    n_records, kind, desc = WORKLOADS[opt.workload]
    if opt.records:
        n_records = opt.records
        desc += f" (records overridden to {n_records})"
    if opt.impl == "reference":
        run_reference(opt, n_records, kind, desc)
    elif opt.mode == "shard":
        run_sharded(opt, n_records, kind, desc)
    else:
        run_ours(opt, n_records, kind, desc)


if __name__ == "__main__":
    main()

"""GPU parity tests: the CUDA sampler, called through the C ABI, against the oracle on the same
seeded inputs.  Integer outputs (links, entity attributes, distortion indicators and counts)
must be identical; theta and the cluster hyper-parameters are host draws from those integers
and must be identical too."""
import numpy as np
import pytest

from checkers import OracleSampler
from exchanger_b200 import BetaRV, PitmanYorRP, GammaRV
from exchanger_b200.analysis import clusters_to_membership, pairwise_measures, smp_clusters
from exchanger_b200.capi import Exb200Error
from exchanger_b200.sampler import Sampler, run_inference
from fixtures import PRIORS, load_rldata, model_with_files, readme_attr_params, readme_model
from exchanger_b200 import exchanger

pytestmark = pytest.mark.gpu

INT_KEYS = ("links", "ent_ids", "ent_attrs", "rec_distortions", "distort_counts")


def assert_same_state(a, b, where=""):
    for k in INT_KEYS:
        assert a[k].shape == b[k].shape and np.array_equal(a[k], b[k]), f"{k} differs {where}"
    assert np.array_equal(a["distort_probs"], b["distort_probs"]), f"theta differs {where}"
    assert np.array_equal(a["distort_dist_concs"], b["distort_dist_concs"]), f"DP concentrations differ {where}"
    for k in ("alpha", "d", "m", "kappa"):
        assert a["clust"][k] == b["clust"][k], f"{k} differs {where}"
    assert a["iteration"] == b["iteration"]


def lockstep(model, n_sweeps, seed=11, every=1, wide_threshold=None, **options):
    gpu, oracle = Sampler(model, seed=seed), OracleSampler(model, seed=seed)
    gpu.set_option("check", 1)
    if wide_threshold is not None:  # part of the sampling specification: set on both sides
        gpu.set_option("wide_threshold", wide_threshold)
        oracle.set_wide_threshold(wide_threshold)
    for k, v in options.items():
        gpu.set_option(k, v)
    stats = []
    for i in range(0, n_sweeps, every):
        for _ in range(every):
            gpu.sweeps(1)
            stats.append(gpu.stats())
        oracle.sweeps(every)
        assert_same_state(gpu.state(), oracle.state(), f"after sweep {i + every}")
    gpu.close()
    return stats


@pytest.mark.parametrize("prior", list(PRIORS))
def test_rldata500_lockstep_every_prior(prior):
    """C1 (RLdata500, README attributes) under every cluster-prior family."""
    m, _ = readme_model("rldata500", clust_prior=PRIORS[prior](500))
    stats = lockstep(m, 30)
    assert max(s["n_rounds"] for s in stats) >= 2  # the dependency rounds were exercised


@pytest.mark.parametrize("wide_buckets", [1, 0])
@pytest.mark.parametrize("threshold", [0, 2, 5])
def test_wide_records_take_their_turn_first(threshold, wide_buckets):
    """Scan order of the specification: records with more than `wide_threshold` possible links go
    first (serial path against every item), the rest follows in ascending id.  Small thresholds
    push many records of RLdata500 through that path - their lists made from the bucket the candidate
    scan found too long and sorted (wide_buckets=1), or by one more scan over all items."""
    m, _ = readme_model("rldata500", n=250)
    stats = lockstep(m, 8, wide_threshold=threshold, wide_buckets=wide_buckets, emit_chunk=65536 if threshold else 16)
    assert sum(s["n_wide"] for s in stats) > (50 if threshold == 0 else 0)


def test_records_without_a_usable_blocking_attribute():
    """Records without any usable blocking attribute (here: every attribute missing or distorted
    in a data set with many NAs): their possible links are listed by the batched scan over all the
    items; those with few of them keep that list and wait for the rounds like everybody else, the
    others are wide by the specification and go first."""
    cols, _ = load_rldata("rldata500", 200)
    rng = np.random.default_rng(1)
    for k in ("fname_c1", "lname_c1", "by", "bm", "bd"):
        cols[k] = [None if rng.random() < 0.45 else v for v in cols[k]]
    m = exchanger(cols, readme_attr_params(), PitmanYorRP(GammaRV(1, 1), BetaRV(1, 1)))
    stats = lockstep(m, 6, wide_threshold=150)
    assert sum(s["n_wide"] for s in stats) > 3


def test_block_staged_and_warp_committed_lists_give_the_same_chain():
    """cand_cap=1 sends every record with two or more candidates through the block-wide staging
    kernel; long_min=0 lets a warp (instead of a thread) take every record's turn."""
    m, _ = readme_model("rldata500")
    stats = lockstep(m, 12, cand_cap=1, long_min=0)
    assert sum(s["n_long"] for s in stats) > 500
    m, _ = readme_model("rldata10000", n=3000, clust_prior=PRIORS["gen_coupon"](3000))
    lockstep(m, 6, long_min=1, est_threshold=0)


def test_missing_values_and_many_attributes():
    """NA paths: the mostly-missing second name components as two extra attributes (7 in all)."""
    cols, _ = load_rldata("rldata500", 300)
    m = exchanger(cols, readme_attr_params(with_c2=True), PitmanYorRP(GammaRV(1, 1), BetaRV(1, 1)))
    assert (m.rec_attrs < 0).sum() > 0
    lockstep(m, 20)


def test_several_files():
    m, _ = model_with_files("rldata500", n_files=3)
    lockstep(m, 20)


def test_entity_value_not_excluded():
    """exclude_entity_value = FALSE: random indicator for agreeing values (records.cpp:109-112),
    empirical new-entity draw (links.cpp:449-452), whole-domain entity weights."""
    m, _ = readme_model("rldata500", n=250, exclude=False)
    lockstep(m, 15)


def test_informative_distortion_prior_and_fixed_theta():
    cols, _ = load_rldata("rldata500")
    ap = readme_attr_params(prior=BetaRV(10, 1000))
    ap["bm"].distort_prob_prior = 0.02  # fixed probability: DistortProbs::update skips it
    m = exchanger(cols, ap, PRIORS["pitman_yor"](500))
    lockstep(m, 25)


def test_rldata10000_lockstep():
    """C2-shaped: RLdata10000 with the README attributes, Ewens prior; the early sweeps have the
    long candidate tails."""
    m, _ = readme_model("rldata10000", clust_prior=PRIORS["ewens"](10000))
    stats = lockstep(m, 6)
    assert stats[-1]["n_candidates"] > 10000


def test_rldata10000_pitman_yor_steady_state():
    m, _ = readme_model("rldata10000")
    lockstep(m, 24, every=8)


def test_link_conditional_matches_oracle():
    m, _ = readme_model("rldata500")
    gpu, oracle = Sampler(m, seed=4), OracleSampler(m, seed=4)
    gpu.sweeps(12)
    oracle.sweeps(12)
    n_cmp = 0
    for r in range(500):
        gi, gw, gn = gpu.link_conditional(r)
        oi, ow, on = oracle.link_conditional(r)
        assert np.array_equal(gi, oi)
        assert np.allclose(gw, ow, rtol=1e-12, atol=0) or np.array_equal(gw, ow)
        assert abs(gn - on) <= 1e-12 * abs(on)
        n_cmp += len(gi)
    assert n_cmp > 50
    gpu.close()


def test_results_do_not_depend_on_batching_thresholds():
    """Same seed, different wide-path thresholds, rounds launched one by one (rounds_tail=0) or in
    the persistent kernel, candidates from the thread-per-record kernel over chain tables or (fast_path=0,
    or chains cut short by hop_cap) from the bucket-scanning kernels -> identical chains (property test that holds at any size: the sweep is
    sequentially consistent, batching is only scheduling)."""
    m, _ = readme_model("rldata10000", n=4000)
    states = []
    for opts in (dict(), dict(cand_cap=4, long_min=2), dict(cand_cap=32, bucket_cap=0, est_threshold=1), dict(bucket_cap=0, huge_min=0),
                 dict(table_cost_permille=1000000), dict(table_cost_permille=0, est_threshold=0),
                 dict(rounds_tail=0), dict(rounds_tail=0, long_min=2), dict(rounds_tail=1, long_min=0),
                 dict(fast_path=0), dict(hop_cap=1), dict(wide_buckets=0), dict(wide_buckets=0, fast_path=0, rounds_tail=0), dict(hop_cap=0, long_min=0), dict(hop_cap=3, cand_cap=4, bucket_cap=0),
                 dict(round0=0), dict(round0=0, rounds_tail=0), dict(table_min_usage=150), dict(table_min_usage=150, emit_cost_cap=0), dict(table_min_usage=150, emit_cost_cap=64),
                 dict(table_min_usage=150, emit_chunk=5), dict(emit_chunk=1)):
        with Sampler(m, seed=3) as g:
            for k, v in opts.items():
                g.set_option(k, v)
            g.sweeps(6)
            states.append(g.state())
    for s in states[1:]:
        assert_same_state(states[0], s)


def test_multi_bucket_probes_and_warp_scanned_buckets():
    """Expensive tables (table_cost_permille high) leave rare key masks to tables with one extra
    small-domain attribute whose every value is probed; bucket_cap=8 sends mid-sized buckets to the
    warp-per-record launch.  Same chain as the oracle either way."""
    cols, _ = load_rldata("rldata10000", 3000)
    rng = np.random.default_rng(5)
    for k in ("fname_c1", "lname_c1", "by"):  # missing values make rare key masks
        cols[k] = [None if rng.random() < 0.15 else v for v in cols[k]]
    m = exchanger(cols, readme_attr_params(), PitmanYorRP(GammaRV(1, 1), BetaRV(1, 1)))
    stats = lockstep(m, 6, table_cost_permille=20000, table_min_usage=0, bucket_cap=8)
    assert sum(s["n_probe"] for s in stats) > 100
    stats = lockstep(m, 4, bucket_cap=0, huge_min=4)  # block-staged lists and tiled buckets
    assert sum(s["n_huge"] for s in stats) > 10


def test_records_no_table_serves_are_listed_through_the_cheapest_table():
    """Rare key masks get no table of their own (table_min_usage high): their records' possible
    links are listed through the cheapest table that IS built - a chain table probed with up to four
    unpinned key attributes, or the bucket of a CSR table - instead of a scan over all the items
    (emit_cost_cap=0 forces that scan).  Same chain as the oracle either way."""
    cols, _ = load_rldata("rldata10000", 4000)
    rng = np.random.default_rng(7)
    for k in ("fname_c1", "lname_c1", "by", "bd"):  # missing values make rare key masks
        cols[k] = [None if rng.random() < 0.15 else v for v in cols[k]]
    m = exchanger(cols, readme_attr_params(), PitmanYorRP(GammaRV(1, 1), BetaRV(1, 1)))
    stats = lockstep(m, 6, table_min_usage=150)
    assert sum(s["n_listed"] for s in stats) > 100, [(s["n_listed"], s["n_scanned"], s["n_tables"]) for s in stats]
    stats = lockstep(m, 4, table_min_usage=150, wide_threshold=20)  # ... some of them wide after all
    assert sum(s["n_listed"] for s in stats) > 100 and sum(s["n_wide"] for s in stats) > 10
    stats = lockstep(m, 4, table_min_usage=150, wide_threshold=20, emit_chunk=7)  # ... chunk by chunk
    assert sum(s["n_listed"] for s in stats) > 100 and sum(s["n_wide"] for s in stats) > 10
    # a scratch too small for the lists of all the records together: chunks are halved until they fit, and a
    # record whose own list does not fit is left to the scan over all items
    stats = lockstep(m, 4, table_min_usage=150, wide_threshold=20, emit_scratch=300)
    assert sum(s["n_listed"] for s in stats) > 50 and sum(s["n_scanned"] for s in stats) > 0
    stats = lockstep(m, 4, table_min_usage=150, emit_cost_cap=0)
    assert sum(s["n_scanned"] for s in stats) > 100 and sum(s["n_listed"] for s in stats) == 0


def test_synthetic_records_at_scale_lockstep():
    """The generator of the 1 M / 10 M workloads at a size the oracle still finishes in seconds
    (120 k records, GeneralizedCoupon as in config C4): identical states sweep by sweep."""
    from bench import DATA_SEED, clust_prior
    from exchanger_b200.synthetic import synth_model
    m, _ = synth_model(120_000, clust_prior("gen_coupon", 120_000), seed=DATA_SEED)
    stats = lockstep(m, 3)
    assert stats[-1]["n_entities"] < 120_000 and stats[-1]["n_tables"] >= 2


@pytest.mark.parametrize("kind", ["gen_coupon", "pitman_yor"])
def test_one_million_records_lockstep_from_the_first_sweep(kind):
    """The benchmark's own generator at 1 M records (config C3, and C4's prior family), from the
    all-singletons start through sweep 30 - the regime bench.py times: multi-bucket probes,
    warp- and block-staged lists, the serial path and a dozen dependency rounds are all live.
    States are compared with the oracle every five sweeps (the oracle searches the shortest bucket
    of the record's pinned attributes, so a sweep costs it about a second)."""
    from bench import DATA_SEED, clust_prior
    from exchanger_b200.synthetic import synth_model
    n = 1_000_000
    m, _ = synth_model(n, clust_prior(kind, n), seed=DATA_SEED)
    stats = lockstep(m, 30, every=5)
    # (tiled buckets need more than 8192 items in one bucket: test_multi_bucket_probes_and_warp_scanned_buckets)
    # (n_listed: records no table served directly, listed through the cheapest built table)
    tot = {k: sum(s[k] for s in stats) for k in ("n_probe", "n_long", "n_wide", "n_listed")}
    assert all(v > 0 for v in tot.values()), tot
    assert max(s["n_rounds"] for s in stats) >= 8 and stats[-1]["n_entities"] < 0.95 * n


def test_sweep_in_slices_equals_the_sweep():
    """The slice calls of within-chain sharding (exb200_shard_*): two slices per phase on one GPU,
    and the ShardedSampler with a single rank, give the chain of exb200_sweeps."""
    import ctypes as C
    import socket
    import torch.distributed as dist
    from exchanger_b200.capi import c_i32p
    from exchanger_b200.sharded import ShardedSampler
    m, _ = readme_model("rldata10000", n=3000, clust_prior=PRIORS["gen_coupon"](3000))
    whole, sliced = Sampler(m, seed=21), Sampler(m, seed=21)
    FA = whole.F * whole.A
    for _ in range(5):
        whole.sweeps(1)
        lib, h = sliced.lib, sliced.h
        sliced._check(lib.exb200_shard_links(h))
        for lo, hi in ((0, 1700), (1700, 3000)):
            sliced._check(lib.exb200_shard_entities(h, lo, hi))
        total = np.zeros((2, FA), np.int32)
        for lo, hi in ((0, 1200), (1200, 3000)):
            nd, co = np.zeros(FA, np.int32), np.zeros(FA, np.int32)
            sliced._check(lib.exb200_shard_distort(h, lo, hi, nd.ctypes.data_as(c_i32p), co.ctypes.data_as(c_i32p)))
            total += np.stack([nd, co])
        sliced._check(lib.exb200_shard_finish(h, total[0].ctypes.data_as(c_i32p), total[1].ctypes.data_as(c_i32p)))
        assert_same_state(whole.state(), sliced.state())
    sliced.close()
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1)
    try:
        sh = ShardedSampler(m, seed=21)
        sh.sweeps(5)
        whole2 = Sampler(m, seed=21)
        whole2.sweeps(5)
        assert_same_state(whole2.state(), sh.state())
        sh.close(); whole2.close()
    finally:
        dist.destroy_process_group()
    whole.close()


def test_run_inference_schedule_history_and_resume():
    """Save schedule of sample() (src/sample.cpp:171-197) and resuming from fit.state
    (R/run_inference.R:104-115): 3 + 3 resumed samples equal one run of 6."""
    m, _ = readme_model("rldata500", n=300)
    fit = run_inference(m, n_samples=6, thin_interval=2, burnin_interval=3, seed=9)
    assert fit.history["links"].shape == (6, 300)
    assert fit.state.iteration == 3 + 2 * 5  # first sample at the last burn-in sweep
    assert fit.mcpar == (3, 13, 2)
    assert fit.history["clust_params"].shape == (6, 2) and fit.history["clust_params_names"] == ["alpha", "d"]
    # the oracle on the same schedule
    o = OracleSampler(m, seed=9)
    o.sweeps(3)
    for s in range(6):
        if s:
            o.sweeps(2)
        st = o.state()
        assert np.array_equal(fit.history["links"][s], st["links"])
        assert fit.history["n_linked_ents"][s, 0] == len(st["ent_ids"])
        assert np.array_equal(fit.history["distort_probs"][s], st["distort_probs"].T.ravel())
        assert np.array_equal(fit.history["distort_counts"][s], st["distort_counts"].T.ravel())
    a = run_inference(m, n_samples=3, thin_interval=2, burnin_interval=3, seed=9)
    b = run_inference(a, n_samples=3, thin_interval=2, seed=9)
    assert np.array_equal(b.history["links"], fit.history["links"])
    assert b.state.iteration == fit.state.iteration


@pytest.mark.parametrize("gamma_prior", [True, False])
def test_finite_dp_concentration_lockstep(gamma_prior):
    """vignette-style model (vignettes/RLdata500.Rmd:45): finite Dirichlet-process concentration on
    every attribute - the urn terms of links.cpp:251-279, the Chinese-restaurant terms of
    entities.cpp:275-288 and DistortDistConcs::update (distort_dist_concs.cpp:32-94)."""
    from exchanger_b200 import DirichletProcess
    cols, _ = load_rldata("rldata500")
    ap = readme_attr_params(prior=BetaRV(1, 4))
    for k in ap:
        ap[k].distort_dist_prior = DirichletProcess(GammaRV(2, 1e-2) if gamma_prior else 25.0)
    m = exchanger(cols, ap, PitmanYorRP(GammaRV(1, 1), BetaRV(1, 1)))
    lockstep(m, 40, every=2)
    gpu = Sampler(m, seed=2)
    gpu.sweeps(30)
    concs = gpu.state()["distort_dist_concs"]
    assert np.all(np.isfinite(concs)) and (not gamma_prior or not np.allclose(concs, 200.0))
    gpu.close()


def test_dirichlet_entity_prior_lockstep():
    """entity_dist_prior = DirichletRV: Entities::update_distributions (entities.cpp:133-156) redraws
    phi every sweep and every phi-dependent table is rebuilt."""
    from exchanger_b200 import DirichletRV
    cols, _ = load_rldata("rldata500", 400)
    ap = readme_attr_params()
    for k in ("bm", "by", "fname_c1"):
        ap[k].entity_dist_prior = DirichletRV(1.0)
    m = exchanger(cols, ap, PitmanYorRP(GammaRV(1, 1), BetaRV(1, 1)))
    lockstep(m, 25)


def test_invalid_models_fail_loudly():
    """Errors the reference raises through Rcpp::stop come back as a status and a message."""
    m, _ = readme_model("rldata500", n=100)
    m.links = m.links.copy()
    m.links[3] = 99999  # an entity that does not exist
    with pytest.raises(Exb200Error) as e:
        Sampler(m)
    assert e.value.status == -1


@pytest.mark.slow
def test_rldata500_posterior_f1():
    """README run (100 samples, thin 10, burn-in 1000): smp_clusters pairwise F1 0.98 +- one pair
    (README.md:137-138; reference over 7 seeds: 0.97-0.99), K around 450."""
    m, ident = readme_model("rldata500")
    fit = run_inference(m, n_samples=100, thin_interval=10, burnin_interval=1000, seed=5)
    pm = pairwise_measures(ident, clusters_to_membership(smp_clusters(fit.history["links"]), 500))
    assert 0.96 <= pm["f1score"] <= 1.0 and pm["tp"] >= 48
    assert abs(fit.history["n_linked_ents"].mean() - 449.65) < 1.5


def test_device_point_estimates_match_the_host_restatement():
    """mp_clusters / smp_clusters on the device (no history on the GPU side) against the host
    restatement of src/point_estimate.cpp on a real chain, from history rows and straight from the
    running chain."""
    from exchanger_b200.analysis import DevicePointEstimate, mp_clusters
    m, _ = readme_model("rldata500")
    g = Sampler(m, seed=5)
    g.sweeps(200)
    rows = []
    pairs = np.array([[0, 1], [10, 11], [3, 250], [498, 499], [42, 42]])
    with DevicePointEstimate(500) as from_rows, DevicePointEstimate(500) as from_chain:
        from_chain.set_pairs(pairs)
        for _ in range(40):
            g.sweeps(5)
            from_chain.add_chain(g)
            rows.append(g.state()["links"].copy())
            from_rows.add(rows[-1])
        hist = np.stack(rows)
        a, b = from_rows.finish(history=hist), from_chain.finish(history=hist)
        from exchanger_b200.chains import coreference_probs
        assert np.array_equal(from_chain.coreference_probs(), coreference_probs(hist, pairs))
    g.close()
    best, prob = mp_clusters(hist)
    want = clusters_to_membership(smp_clusters((best, prob)), 500)
    for got in (a, b):
        assert got["n_overflow"] == 0
        assert np.allclose(got["probabilities"], prob)
        # same partition: a bijection between the two labelings
        pairs = set(zip(got["membership"].tolist(), want.tolist()))
        assert len(pairs) == len(set(got["membership"].tolist())) == len(set(want.tolist()))
        # the first sample in which the most probable cluster appeared holds exactly that cluster
        for r in (0, 17, 250, 499):
            if not got["tied"][r]:
                s = got["first_sample"][r]
                assert tuple(np.flatnonzero(hist[s] == hist[s, r]).tolist()) == best[r]
    with pytest.raises(Exb200Error):
        with DevicePointEstimate(4) as pe:
            pe.add(np.array([0, 1, 9, 3], np.int32))  # label outside [0, 2N)


def test_run_inference_with_device_point_estimate_and_no_link_history():
    """run_inference(point_estimate=True): the estimates accumulated on the device while the chain
    runs equal those computed from the history; with keep_links=False no link history is kept."""
    m, _ = readme_model("rldata500", n=300)
    fit = run_inference(m, n_samples=30, thin_interval=3, burnin_interval=60, seed=2, point_estimate=True)
    want = clusters_to_membership(smp_clusters(fit.history["links"]), 300)
    got = fit.point_estimate["membership"]
    assert len(set(zip(got.tolist(), want.tolist()))) == len(set(got.tolist())) == len(set(want.tolist()))
    lean = run_inference(m, n_samples=30, thin_interval=3, burnin_interval=60, seed=2, point_estimate=True, keep_links=False)
    assert lean.history["links"].shape == (0, 300)
    assert np.array_equal(lean.point_estimate["probabilities"], fit.point_estimate["probabilities"])
    assert np.array_equal(lean.history["n_linked_ents"], fit.history["n_linked_ents"])
    if not fit.point_estimate["tied"].any():  # ties are settled from the history, which `lean` does not have
        assert np.array_equal(lean.point_estimate["membership"], fit.point_estimate["membership"])


def test_abort_and_progress_callbacks_keep_the_partial_history():
    """abort_cb / progress_cb of exb200_run play Progress::check_abort / p.increment
    (src/sample.cpp:200-201): an abort after the 7th sweep stops the loop, the rows saved so far stay
    (the reference breaks out of its loop and returns the partly filled matrices), the chain stands
    at the aborted sweep and can go on."""
    m, _ = readme_model("rldata500", n=200)
    seen = []
    with Sampler(m, seed=4) as g:
        hist = g.run(10, thin_interval=2, burnin_interval=0, abort=lambda: len(seen) >= 6, progress=seen.append)
        assert hist.get("aborted") and hist["n_saved"] == 3 and g.last_n_saved == 3   # sweeps 2, 4, 6 were saved
        assert seen == list(range(1, 7))          # the abort is polled before the progress call, as in sample.cpp:200-201
        assert g.state()["iteration"] == 7
        o = OracleSampler(m, seed=4)
        for s in range(3):
            o.sweeps(2)
            assert np.array_equal(hist["links"][s], o.state()["links"])
        assert not hist["links"][3:].any()        # rows never reached stay as allocated
        o.sweeps(1)
        assert_same_state(g.state(), o.state(), "at the aborted sweep")
        again = g.run(2, thin_interval=1)         # the handle is still usable
        o.sweeps(2)
        assert "aborted" not in again and np.array_equal(again["links"][1], o.state()["links"])


def test_invalid_weights_are_reported_like_the_reference():
    """A record without any valid link - entity-value pmf zero at its value, own singleton entity
    removed - is the reference's "no valid links for record" (src/links.cpp:415-419, CHECK_WEIGHTS /
    AliasSampler::check_weights src/alias_sampler.h:76-97): EXB200_ERR_WEIGHTS here, an error in the
    oracle, and the handle refuses further work."""
    cols, _ = load_rldata("rldata500", 120)
    ap = readme_attr_params()
    months = sorted({v for v in cols["bm"] if v is not None})
    pmf = {v: 1.0 / (len(months) - 1) for v in months}
    pmf[cols["bm"][0]] = 0.0                       # record 0 (and its like) can neither stay nor found an entity
    ap["bm"].entity_dist_prior = pmf
    m = exchanger(cols, ap, PitmanYorRP(GammaRV(1, 1), BetaRV(1, 1)))
    with Sampler(m, seed=1) as g:
        with pytest.raises(Exb200Error) as e:
            g.sweeps(1)
        assert e.value.status == -5 and "no valid outcome" in str(e.value)
        with pytest.raises(Exb200Error):
            g.sweeps(1)
    o = OracleSampler(m, seed=1)
    with pytest.raises(RuntimeError):
        o.sweeps(1)


def test_random_coupon_m_lockstep():
    """GeneralizedCouponRP with a ShiftedNegBinomRV prior on m: the rnbinom draw of
    GenCouponClustParams::update (src/clust_params.cpp:66-71) moves m every sweep."""
    from exchanger_b200 import GeneralizedCouponRP, ShiftedNegBinomRV
    cols, _ = load_rldata("rldata500", 300)
    m = exchanger(cols, readme_attr_params(), GeneralizedCouponRP(ShiftedNegBinomRV(600, 0.5), GammaRV(1, 1)))
    lockstep(m, 25)
    with Sampler(m, seed=3) as g:
        hist = g.run(12, 1, 0)
    assert hist["clust_params_names"] == ["kappa", "m"]
    ms = hist["clust_params"][:, 1]
    assert len(set(ms.tolist())) > 3 and np.all(ms == np.round(ms)) and np.all(ms >= hist["n_linked_ents"][:, 0])


def test_device_point_estimate_with_more_clusters_than_slots():
    """A record seen in more distinct clusters than the device table has slots: the dropped
    occurrences are counted, the record is flagged, and with the history at hand its estimate is
    redone on the host, so the result still equals the restatement of src/point_estimate.cpp;
    without the history a RuntimeWarning names the problem."""
    from exchanger_b200.analysis import DevicePointEstimate, mp_clusters
    rng = np.random.default_rng(3)
    n, n_samples = 60, 24
    hist = np.stack([rng.integers(0, 12, n) for _ in range(n_samples)]).astype(np.int32)  # every sample a fresh partition
    best, prob = mp_clusters(hist)
    want = clusters_to_membership(smp_clusters((best, prob)), n)
    with DevicePointEstimate(n, slots=2) as pe:
        for row in hist:
            pe.add(row)
        got = pe.finish(hist)
        assert got["n_overflow"] > 0 and not got["overflowed"].any()
        assert np.allclose(got["probabilities"], prob)
        _, inv_w = np.unique(want, return_inverse=True)
        firsts = {}
        canon = np.array([firsts.setdefault(k, i) for i, k in enumerate(inv_w)])
        assert np.array_equal(got["membership"], canon)
        with pytest.warns(RuntimeWarning, match="distinct clusters"):
            lean = pe.finish(None)
        assert lean["overflowed"].any()
    with DevicePointEstimate(n, slots=n_samples) as pe:   # slots >= samples can never overflow
        for row in hist:
            pe.add(row)
        full = pe.finish(hist)
        assert full["n_overflow"] == 0 and np.array_equal(full["membership"], canon)


@pytest.mark.parametrize("include_self", [False, True])
def test_neighbourhood_builder_matches_compute_close_values(include_self):
    """k_neighbours (exb200_neighbours_*) against the restatement of compute_close_values
    (R/AttributeIndex.R:117-137): the name domains of RLdata10000 and ~5 k synthetic names in two domains (with
    typo variants, duplicates of length, isolated values and the empty string), threshold 3 and 1:
    exact row_ptr / val_ids, probabilities and max_exp_factor to 1e-12."""
    from exchanger_b200.capi import levenshtein_neighbours
    from exchanger_b200.model import Levenshtein, compute_close_values, transform_dist_fn
    from exchanger_b200.synthetic import synth_rldata
    cols, _ = load_rldata("rldata10000")
    domains = []
    for name in ("fname_c1", "lname_c1"):
        seen = []
        for v in cols[name]:
            if v is not None and v not in seen:
                seen.append(v)
        domains.append(seen)
    _, doms, _, _ = synth_rldata(50_000)                        # D(fname) ~ 2 k, D(lname) ~ 5 k with typo variants
    domains.append(doms[3] + ["", "X", "QQQQQQQQQQQQQQ", "ZZZZZZZZZZZZZZZZZZZZZZZZZZZZZZZ"])  # isolated values, 31 bytes
    domains.append(doms[4][:3000])
    for dom in domains:
        for thr in (3, 1):
            rp, ci, pr, mef = levenshtein_neighbours(dom, thr, include_self=include_self)
            rp2, ci2, pr2, mef2 = compute_close_values(dom, transform_dist_fn(Levenshtein(), thr), include_self)
            assert np.array_equal(rp, rp2) and np.array_equal(ci, ci2), (len(dom), thr)
            assert np.allclose(pr, pr2, rtol=1e-12, atol=0) and np.allclose(mef, mef2, rtol=1e-12, atol=0)
            if not include_self:
                iso = np.flatnonzero(np.diff(rp2) == 0)
                assert np.all(mef[iso] == 0.0)                   # a value without a finite neighbour (AttributeIndex.R:135)
    with pytest.raises(ValueError):        # 31 bytes is the limit of the packed layout
        levenshtein_neighbours(["A" * 32], 3)
    with pytest.raises(Exb200Error):       # ... and 8 of the threshold
        levenshtein_neighbours(["AB", "AC"], 9)


def test_r_glue_end_to_end_on_the_gpu():
    """exchanger_b200/rglue/sample_b200.cpp (the replacement of src/sample.cpp, compiled against the
    stand-in Rcpp headers) linked with the REAL libexb200.so: `.sample()` on an ExchangERModel S4 object
    returns the chain that the Python binding returns for the same Philox key - history, final state,
    1-based labels, R's column-major matrices."""
    import os
    import test_rglue as tg
    gpu_so = os.path.join(tg.ROOT, "oracle", "_build", "librglue_gpu.so")
    if not os.path.exists(gpu_so):
        pytest.skip("glue harness not built")
    lib = tg.harness(gpu_so)
    m, _ = readme_model("rldata500", n=200)
    out = tg.run(lib, m, 5, 2, 4, which=1, seed=77)
    fit = run_inference(m, n_samples=5, thin_interval=2, burnin_interval=4, seed=int(lib.rglue_seed_for(77)))
    N, A = m.n_records, m.n_attrs
    slots = out["fit"]["slots"]
    hist = dict(zip(slots["history"]["names"], slots["history"]["items"]))
    assert list(hist) == ["links", "distort_probs", "distort_counts", "distort_dist_concs", "n_linked_ents", "clust_params"]
    links = np.array(hist["links"]["values"]).reshape(N, 5).T
    assert np.array_equal(links, fit.history["links"] + 1)
    assert hist["n_linked_ents"]["values"] == fit.history["n_linked_ents"][:, 0].tolist()
    assert np.array_equal(np.array(hist["distort_probs"]["values"]).reshape(A, 5).T, fit.history["distort_probs"])
    assert np.array_equal(np.array(hist["distort_counts"]["values"]).reshape(A, 5).T, fit.history["distort_counts"])
    assert np.array_equal(np.array(hist["clust_params"]["values"]).reshape(2, 5).T, fit.history["clust_params"])
    assert hist["clust_params"]["attrs"]["colnames"]["values"] == ["alpha", "d"]
    st = slots["state"]["slots"]
    assert st["iteration"]["values"] == [fit.state.iteration] == [12]
    assert np.array_equal(np.array(st["links"]["values"]), fit.state.links + 1)
    K = len(fit.state.ent_ids)
    assert np.array_equal(np.array(st["ent_attrs"]["values"]).reshape(K, A).T, fit.state.ent_attrs + 1)
    assert np.array_equal(np.array(st["rec_distortions"]["values"]).reshape(N, A).T, fit.state.rec_distortions)
    assert out["increments"] == 12


def test_handles_reuse_the_memory_of_destroyed_ones():
    """Device slabs and pinned history rows of a destroyed handle are kept for the next one
    (``exb200_trim`` gives them back): chains run one after the other - and after a trim - are the chain
    a fresh process would run, history rows included."""
    from exchanger_b200 import capi
    m, _ = readme_model("rldata10000", n=3000)
    runs = []
    for k in range(4):
        if k == 3:
            capi.trim()
        with Sampler(m, seed=5) as g:
            hist = g.run(4, 2, 3)
            runs.append((g.state(), hist["links"].copy()))
    for st, links in runs[1:]:
        assert_same_state(runs[0][0], st)
        assert np.array_equal(runs[0][1], links)
    big, _ = readme_model("rldata10000")          # a larger model after smaller ones: slabs of another size
    with Sampler(big, seed=5) as g:
        g.set_option("check", 1)
        g.sweeps(3)
    capi.trim()

"""Multi-chain host logic on CPU: two ranks over gloo."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from exchanger_b200.chains import chain_seed, coreference_counts, coreference_probs, merged_coreference_probs


def test_chain_seeds_are_distinct():
    seeds = {chain_seed(7, r) for r in range(64)} | {chain_seed(8, r) for r in range(64)}
    assert len(seeds) == 128


def test_coreference_probs_single_chain():
    links = np.array([[0, 0, 2, 3], [0, 1, 1, 3], [0, 0, 2, 2]])
    pairs = np.array([[0, 1], [2, 3], [1, 2]])
    assert coreference_counts(links, pairs).tolist() == [2, 1, 1]
    assert np.allclose(coreference_probs(links, pairs), [2 / 3, 1 / 3, 1 / 3])
    with pytest.raises(ValueError):
        coreference_counts(links, np.array([[0, 9]]))


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    links = rng.integers(0, 4, size=(5 + rank, 12))  # chains of different lengths
    pairs = np.array([[0, 1], [2, 3], [4, 11]])
    merged = merged_coreference_probs(links, pairs)
    out[rank] = (links, merged)
    dist.destroy_process_group()


def test_merged_coreference_probs_two_ranks_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    pairs = np.array([[0, 1], [2, 3], [4, 11]])
    both = np.concatenate([out[0][0], out[1][0]], axis=0)
    expect = coreference_probs(both, pairs)
    assert np.allclose(out[0][1], expect) and np.allclose(out[1][1], expect)


# ---- within-chain sharding: slice arithmetic and the exchange pattern (two ranks over gloo)
def test_slice_bounds_cover_the_range():
    from exchanger_b200.sharded import slice_bounds
    for n in (2, 7, 500, 10_000_019):
        for world in (1, 2, 3, 8):
            if world > n:
                continue
            chunks = [slice_bounds(n, world, r) for r in range(world)]
            assert all(c[0] == chunks[0][0] for c in chunks) and chunks[0][0] * world <= 2 * n
            covered = np.concatenate([np.arange(lo, hi) for _, lo, hi in chunks])
            assert np.array_equal(covered, np.arange(n))


def _shard_worker(rank, world, port, out):
    import torch
    from exchanger_b200.sharded import exchange_ragged, exchange_slices, slice_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 11
    chunk, lo, hi = slice_bounds(n, world, rank)
    truth_rows = torch.arange(2 * n * 8, dtype=torch.int32).reshape(2 * n, 8)
    rows = torch.zeros_like(truth_rows)
    rows[lo:hi] = truth_rows[lo:hi]                     # what this rank computed
    rows[n:] = truth_rows[n:]                           # slots beyond N are replicated anyway
    exchange_slices(dist, rows, chunk, rank, world)
    truth_z = torch.arange(100, 100 + n, dtype=torch.int32)
    z = torch.zeros(n, dtype=torch.int32)
    z[lo:hi] = truth_z[lo:hi]
    exchange_ragged(dist, z, torch.zeros(world * chunk, dtype=torch.int32), chunk, lo, hi, rank, world)
    out[rank] = (bool(torch.equal(rows, truth_rows)), bool(torch.equal(z, truth_z)))
    dist.destroy_process_group()


def test_slice_exchange_two_ranks_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = mp.Manager().dict()
    mp.spawn(_shard_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] == (True, True) and out[1] == (True, True)


class _ToyChain:
    """Stand-in for the library behind `sharded_sweep`: deterministic toy phases whose results depend
    on the other ranks' slices, so a missing or misplaced exchange changes the outcome."""

    def __init__(self, n, world, rank):
        import torch
        from exchanger_b200.sharded import slice_bounds
        self.N, self.FA, self.it = n, 5, 0
        self.chunk, self.lo, self.hi = slice_bounds(n, world, rank)
        self.ent_y = torch.zeros(2 * n, 8, dtype=torch.int32)
        self.rec_z = torch.zeros(n, dtype=torch.int32)
        self.z_stage = torch.zeros(world * self.chunk, dtype=torch.int32)
        self.counts = torch.zeros(2 * self.FA, dtype=torch.int64)
        self.theta = []

    def links(self):
        import torch
        self.ent_y[self.N:] = (torch.arange(self.N, dtype=torch.int32)[:, None] + self.it) % 97   # replicated part

    def entities(self, lo, hi):
        import torch
        rows = torch.arange(lo, hi, dtype=torch.int32)
        self.ent_y[lo:hi] = (rows[:, None] * 7 + self.it * 3 + torch.arange(8, dtype=torch.int32) + self.rec_z[lo:hi, None]) % 1000

    def distort(self, lo, hi):
        import torch
        r = torch.arange(lo, hi)
        self.rec_z[lo:hi] = (self.ent_y[(r * 13 + 5) % self.N, 0] + self.ent_y[r, 1] + self.it) & 31   # reads other slices
        z = self.rec_z[lo:hi]
        nd = np.array([int(((z >> a) & 1).sum()) for a in range(self.FA)], np.int32)
        return nd, (nd[::-1] * 2).astype(np.int32)   # both count matrices are sums over records

    def finish(self, nd, co):
        self.theta.append((int(nd.sum()), int(co.sum())))
        self.it += 1

    def fence(self):
        pass


def _toy_worker(rank, world, port, out):
    from exchanger_b200.sharded import sharded_sweep
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    chain = _ToyChain(23, world, rank)
    for _ in range(4):
        sharded_sweep(chain, dist, rank, world)
    out[rank] = (chain.ent_y[:23].numpy().copy(), chain.rec_z.numpy().copy(), chain.theta)
    dist.destroy_process_group()


def test_sharded_sweep_equals_the_unsharded_one_two_ranks_gloo():
    from exchanger_b200.sharded import sharded_sweep
    whole = _ToyChain(23, 1, 0)
    for _ in range(4):
        sharded_sweep(whole, None, 0, 1)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = mp.Manager().dict()
    mp.spawn(_toy_worker, args=(2, port, out), nprocs=2, join=True)
    for rank in (0, 1):
        y, z, theta = out[rank]
        assert np.array_equal(y, whole.ent_y[:23].numpy()) and np.array_equal(z, whole.rec_z.numpy())
        assert theta == whole.theta


def test_benchmark_model_survives_pickling():
    """bench.py builds the synthetic model once per node (rank 0) and hands it to the other ranks as a
    pickle in /dev/shm: everything a model holds must be picklable (a closure in an attribute's
    distance function once was not, and every multi-GPU run of the benchmark died on rank 0)."""
    import pickle
    import bench
    from checkers import cpu_neighbours
    from exchanger_b200.capi import FlatModel
    from exchanger_b200.synthetic import synth_model
    m, _ = synth_model(3000, bench.clust_prior("gen_coupon", 3000), seed=bench.DATA_SEED, neighbour_fn=cpu_neighbours)
    m2 = pickle.loads(pickle.dumps(m, protocol=pickle.HIGHEST_PROTOCOL))
    a, b = FlatModel(m, 1), FlatModel(m2, 1)
    assert a.nbytes() == b.nbytes()
    assert np.array_equal(m.rec_attrs, m2.rec_attrs) and np.array_equal(m.links, m2.links)
    d = m2.attr_params[3].dist_fn  # the transformed Levenshtein distance still works
    assert callable(d)

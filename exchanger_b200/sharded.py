"""Within-chain sharding over the GPUs of one node (BASELINE config 4).

One process per GPU (``torch.distributed``, NCCL).  Every rank holds the whole chain and runs the
links update identically (keyed randomness: the replicas agree bit for bit without talking);
entity attributes and distortion indicators are computed on the rank's slice of the entity slots /
records and exchanged over NVLink: one all-gather of entity tuples, one of distortion masks, one
all-reduce of the two ``F x A`` count matrices per sweep.  The result is the
single-GPU chain, whatever the number of ranks.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import c_i32p
from .sampler import Sampler


def slice_bounds(n: int, world: int, rank: int) -> tuple[int, int, int]:
    """Equal slices of ``chunk = ceil(n / world)`` entity slots / records: ``(chunk, lo, hi)``."""
    chunk = (n + world - 1) // world
    lo = min(n, rank * chunk)
    return chunk, lo, min(n, lo + chunk)


def exchange_slices(dist, full, chunk: int, rank: int, world: int, group=None):
    """All-gather of equal slices in place: on return rows ``[r * chunk, (r + 1) * chunk)`` of
    ``full`` hold what rank ``r`` had there.  ``full`` has at least ``world * chunk`` rows."""
    out = full[: world * chunk]
    dist.all_gather_into_tensor(out, out[rank * chunk:(rank + 1) * chunk].clone(), group=group)
    return out


def exchange_ragged(dist, values, stage, chunk: int, lo: int, hi: int, rank: int, world: int, group=None):
    """All-gather of the slices ``values[lo:hi]`` of an array whose length is not a multiple of the
    slice size, through the padded staging array ``stage`` (``world * chunk`` entries)."""
    stage[rank * chunk: rank * chunk + (hi - lo)] = values[lo:hi]
    exchange_slices(dist, stage, chunk, rank, world, group)
    values.copy_(stage[: values.shape[0]])
    return values


def sharded_sweep(chain, dist, rank: int, world: int, group=None):
    """One sweep of a chain held by every rank.  ``chain`` provides the slice
    calls (``links()``, ``entities(lo, hi)``, ``distort(lo, hi) -> (ndist, corr)``,
    ``finish(ndist, corr)``, ``fence()``), the exchanged arrays (``ent_y`` [rows, ints], ``rec_z`` [N],
    staging ``z_stage`` [world * chunk], ``counts`` [2 * FA] int64) and ``N``, ``FA``, ``chunk``,
    ``lo``, ``hi``.  Collectives: one all-gather of the entity tuples, one of the distortion masks,
    one all-reduce of the two count matrices."""
    chain.links()                                   # replicated: keyed randomness, identical on every rank
    chain.entities(chain.lo, chain.hi)
    chain.fence()
    if world > 1:
        exchange_slices(dist, chain.ent_y, chain.chunk, rank, world, group)
        chain.fence()
    nd, co = chain.distort(chain.lo, chain.hi)
    if world > 1:
        exchange_ragged(dist, chain.rec_z, chain.z_stage, chain.chunk, chain.lo, chain.hi, rank, world, group)
        chain.counts.copy_(chain.counts.new_tensor(np.concatenate([nd, co]).astype(np.int64)))
        dist.all_reduce(chain.counts, group=group)
        chain.fence()
        both = chain.counts.cpu().numpy()
        nd, co = both[: chain.FA].astype(np.int32), both[chain.FA:].astype(np.int32)
    chain.finish(nd, co)


class _DeviceArray:
    """Zero-copy view of a library buffer for ``torch.as_tensor`` (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, shape: tuple, typestr: str):
        self.__cuda_array_interface__ = dict(shape=shape, typestr=typestr, data=(ptr, False), version=2)


class ShardedSampler:
    """``Sampler`` whose sweeps are sharded over the ranks of a ``torch.distributed`` group."""

    def __init__(self, model, seed: int = 0, device: int = 0, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.sampler = Sampler(model, seed=seed, device=device)   # same seed on every rank: same chain
        self.lib, self.h = self.sampler.lib, self.sampler.h
        self.N, self.FA = self.sampler.N, self.sampler.F * self.sampler.A
        self.chunk, self.lo, self.hi = slice_bounds(self.N, self.world, self.rank)
        self.device = torch.device("cuda", device)
        ptr, nbytes, row = C.c_void_p(), C.c_int64(), C.c_int32()
        self.sampler._check(self.lib.exb200_device_buffer(self.h, b"ent_y", C.byref(ptr), C.byref(nbytes), C.byref(row)))
        self.row_ints = row.value // 4
        # entity tuples: slots [0, world * chunk) are exchanged; the slots beyond N hold potential
        # entities that the next links update redraws anyway
        self.ent_y = torch.as_tensor(_DeviceArray(ptr.value, (2 * self.N, self.row_ints), "<i4"), device=self.device)
        self.sampler._check(self.lib.exb200_device_buffer(self.h, b"rec_z", C.byref(ptr), C.byref(nbytes), C.byref(row)))
        self.rec_z = torch.as_tensor(_DeviceArray(ptr.value, (self.N,), "<i4"), device=self.device)
        self.z_stage = torch.zeros(self.world * self.chunk, dtype=torch.int32, device=self.device)
        self.counts = torch.zeros(2 * self.FA, dtype=torch.int64, device=self.device)

    # ---- the slice calls of the C ABI, as the backend of `sharded_sweep`
    def links(self):
        self.sampler._check(self.lib.exb200_shard_links(self.h))

    def entities(self, lo: int, hi: int):
        self.sampler._check(self.lib.exb200_shard_entities(self.h, lo, hi))

    def distort(self, lo: int, hi: int):
        nd, co = np.zeros(self.FA, np.int32), np.zeros(self.FA, np.int32)
        self.sampler._check(self.lib.exb200_shard_distort(self.h, lo, hi, nd.ctypes.data_as(c_i32p), co.ctypes.data_as(c_i32p)))
        return nd, co

    def finish(self, nd: np.ndarray, co: np.ndarray):
        nd, co = np.ascontiguousarray(nd, np.int32), np.ascontiguousarray(co, np.int32)
        self.sampler._check(self.lib.exb200_shard_finish(self.h, nd.ctypes.data_as(c_i32p), co.ctypes.data_as(c_i32p)))

    def fence(self):
        """The library's stream and torch's: nothing of either is in flight afterwards."""
        self.sampler._check(self.lib.exb200_stream_sync(self.h))
        self.torch.cuda.synchronize(self.device)

    def sweeps(self, n: int = 1):
        for _ in range(n):
            sharded_sweep(self, self.dist, self.rank, self.world, self.group)

    def state(self):
        return self.sampler.state()

    def stats(self):
        return self.sampler.stats()

    def close(self):
        self.sampler.close()

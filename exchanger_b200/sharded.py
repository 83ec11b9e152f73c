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

    def sweeps(self, n: int = 1):
        torch, dist, lib, h, chk = self.torch, self.dist, self.lib, self.h, self.sampler._check
        nd, co = np.zeros(self.FA, np.int32), np.zeros(self.FA, np.int32)
        for _ in range(n):
            chk(lib.exb200_shard_links(h))
            chk(lib.exb200_shard_entities(h, self.lo, self.hi))
            chk(lib.exb200_stream_sync(h))
            if self.world > 1:   # entity tuples of every slice, in place
                exchange_slices(dist, self.ent_y, self.chunk, self.rank, self.world, self.group)
                torch.cuda.synchronize(self.device)
            chk(lib.exb200_shard_distort(h, self.lo, self.hi, nd.ctypes.data_as(c_i32p), co.ctypes.data_as(c_i32p)))
            if self.world > 1:   # distortion masks of every slice, counts summed
                exchange_ragged(dist, self.rec_z, self.z_stage, self.chunk, self.lo, self.hi, self.rank, self.world, self.group)
                self.counts.copy_(torch.from_numpy(np.concatenate([nd, co]).astype(np.int64)))
                dist.all_reduce(self.counts, group=self.group)
                torch.cuda.synchronize(self.device)
                both = self.counts.cpu().numpy().astype(np.int32)
                nd, co = np.ascontiguousarray(both[: self.FA]), np.ascontiguousarray(both[self.FA:])
            chk(lib.exb200_shard_finish(h, nd.ctypes.data_as(c_i32p), co.ctypes.data_as(c_i32p)))

    def state(self):
        return self.sampler.state()

    def stats(self):
        return self.sampler.stats()

    def close(self):
        self.sampler.close()

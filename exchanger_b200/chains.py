"""Independent chains, one per GPU (BASELINE.json configs[4]): seeds per rank and the merge of the
chains' link histories into posterior coreference probabilities.

The reference runs one chain per `run_inference` call and has no multi-chain helper
(SURVEY.md 2.3); `coreference_probs` (R/coreference_probs.R:54-82) is the per-chain statistic that
is merged here.  No collective touches the sampler's data path: only these small count vectors are
all-reduced (gloo on CPU, NCCL on GPUs) after sampling.
"""
from __future__ import annotations

import numpy as np


def chain_seed(base_seed: int, rank: int) -> int:
    """Philox key of chain `rank`: distinct keys give independent streams; the key is the only
    thing that differs between chains."""
    return (int(base_seed) * 0x9E3779B97F4A7C15 + int(rank) + 1) & 0xFFFFFFFFFFFFFFFF


def coreference_counts(links: np.ndarray, pairs: np.ndarray) -> np.ndarray:
    """For each record pair (0-based indices), the number of samples in which both records are
    linked to the same entity - the numerator of coreference_probs (R/coreference_probs.R:77-79)."""
    links = np.asarray(links)
    pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    if links.ndim != 2:
        raise ValueError("`links` must be an n_samples x n_records matrix")
    if pairs.size and (pairs.min() < 0 or pairs.max() >= links.shape[1]):
        raise ValueError("`pairs` refers to records that are not in `links`")
    return (links[:, pairs[:, 0]] == links[:, pairs[:, 1]]).sum(axis=0).astype(np.int64)


def coreference_probs(links: np.ndarray, pairs: np.ndarray) -> np.ndarray:
    """Single-chain posterior coreference probabilities (R/coreference_probs.R:54-82)."""
    links = np.asarray(links)
    return coreference_counts(links, pairs) / float(links.shape[0])


def merged_coreference_probs(links: np.ndarray, pairs: np.ndarray, group=None) -> np.ndarray:
    """Coreference probabilities over ALL chains of a `torch.distributed` job: each rank passes
    the link history of its own chain; counts and sample numbers are summed over ranks."""
    import torch
    import torch.distributed as dist

    counts = coreference_counts(links, pairs)
    buf = torch.from_numpy(np.concatenate([counts, [np.asarray(links).shape[0]]]).astype(np.int64))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        if dist.get_backend(group) == "nccl":
            buf = buf.cuda()
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        buf = buf.cpu()
    out = buf.numpy()
    return out[:-1] / float(out[-1])

"""Point estimates from the link history: mirror of ``mp_clusters`` / ``smp_clusters``
(R/analysis.R:53-186, src/point_estimate.cpp:58-136) and the pairwise measures the README
reports (README.md:117-144, computed there with the ``clevr`` package).

Host post-processing that runs once per fit; it is outside the Gibbs sweep.
"""
from __future__ import annotations

from collections import Counter, defaultdict

import numpy as np


def _clusters_of(membership: np.ndarray):
    """membership vector -> list of clusters as sorted record-index tuples
    (point_estimate.cpp:26-40)."""
    order = np.argsort(membership, kind="stable")
    sorted_m = membership[order]
    cuts = np.flatnonzero(np.diff(sorted_m)) + 1
    return [tuple(g.tolist()) for g in np.split(order, cuts)]


def mp_clusters(samples: np.ndarray):
    """Most probable cluster of every record (point_estimate.cpp:58-103).

    ``samples`` is the ``n_samples x N`` link history.  Returns ``(clusters, probabilities)``:
    for each record the cluster (tuple of 0-based record indices) that contains it most often
    along the chain and its relative frequency.  Ties go to the cluster that sorts first (the
    reference's tie-break follows the iteration order of an unordered_map and is unspecified).
    """
    samples = np.asarray(samples)
    n_samples, n = samples.shape
    freq: Counter = Counter()
    for s in range(n_samples):
        freq.update(_clusters_of(samples[s]))
    best: list = [None] * n
    best_f = np.zeros(n, np.int64)
    for cl in sorted(freq):
        f = freq[cl]
        for r in cl:
            if f > best_f[r]:
                best_f[r] = f
                best[r] = cl
    return best, best_f / float(n_samples)


def smp_clusters(samples_or_mp):
    """Shared most probable clusters (point_estimate.cpp:105-136): records whose most probable
    clusters are identical are grouped together.  Accepts a link history or the result of
    :func:`mp_clusters`; returns a list of clusters (lists of 0-based record indices)."""
    if isinstance(samples_or_mp, tuple):
        best = samples_or_mp[0]
    else:
        best = mp_clusters(samples_or_mp)[0]
    groups = defaultdict(list)
    for r, cl in enumerate(best):
        groups[cl].append(r)
    return [groups[k] for k in sorted(groups)]


def clusters_to_membership(clusters, n: int) -> np.ndarray:
    out = np.full(n, -1, np.int64)
    for i, cl in enumerate(clusters):
        out[np.asarray(cl, dtype=np.int64)] = i
    return out


def pairwise_measures(true_membership: np.ndarray, pred_membership: np.ndarray) -> dict:
    """Pairwise precision / recall / F1 of a predicted clustering (the measures of
    README.md:125-144)."""
    t = np.asarray(true_membership)
    p = np.asarray(pred_membership)

    def n_pairs(m):
        _, c = np.unique(m, return_counts=True)
        return int((c * (c - 1) // 2).sum())

    joint = np.stack([t, p], axis=1)
    _, c = np.unique(joint, axis=0, return_counts=True)
    tp = int((c * (c - 1) // 2).sum())
    true_pairs, pred_pairs = n_pairs(t), n_pairs(p)
    prec = tp / pred_pairs if pred_pairs else 1.0
    rec = tp / true_pairs if true_pairs else 1.0
    f1 = 2 * prec * rec / (prec + rec) if prec + rec else 0.0
    return dict(precision=prec, recall=rec, f1score=f1, true_pairs=true_pairs, pred_pairs=pred_pairs, tp=tp)

"""Point estimates from the link history: mirror of ``mp_clusters`` / ``smp_clusters``
(R/analysis.R:53-186, src/point_estimate.cpp:58-136) and the pairwise measures the README
reports (README.md:117-144, computed there with the ``clevr`` package).

Host post-processing that runs once per fit; it is outside the Gibbs sweep.
:class:`DevicePointEstimate` computes the same estimates on the GPU from the stream of saved link
vectors, without the ``n_samples x N`` history.
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


class DevicePointEstimate:
    """``mp_clusters`` / ``smp_clusters`` on the device (``exb200_pe_*`` of the C ABI): add the saved
    samples one at a time - rows of a history (``add``) or the current links of a running chain
    (``add_chain``, nothing leaves the GPU) - then ``finish``.

    A cluster is identified by a 128-bit fingerprint of its member set.  When two clusters of a
    record are equally frequent the reference keeps the lexicographically smaller member set
    (``std::map<set<int>, int>`` order, point_estimate.cpp:65-87); the device keeps the smaller
    fingerprint and flags the record in ``tied``.  With the history at hand (``finish(history=...)``)
    tied records are settled on the host exactly as the reference does."""

    def __init__(self, n_records: int, slots: int = 8, device: int = 0):
        import ctypes as C
        from . import capi
        self._C, self._capi = C, capi
        self.lib = capi.load_library()
        self.n = int(n_records)
        self.h = C.c_void_p()
        rc = self.lib.exb200_pe_create(self.n, int(slots), int(device), C.byref(self.h))
        if rc != capi.EXB200_OK:
            self.h = None
            raise capi.Exb200Error(rc, self.lib.exb200_last_error(None).decode())

    def _check(self, rc):
        if rc != self._capi.EXB200_OK:
            raise self._capi.Exb200Error(rc, self.lib.exb200_last_error(None).decode())

    def add(self, links: np.ndarray):
        row = np.ascontiguousarray(links, dtype=np.int32)
        if row.shape != (self.n,):
            raise ValueError("one link per record expected")
        self._check(self.lib.exb200_pe_add(self.h, row.ctypes.data_as(self._capi.c_i32p)))

    def add_chain(self, sampler):
        self._check(self.lib.exb200_pe_add_chain(self.h, sampler.h))

    def set_pairs(self, pairs: np.ndarray):
        """Record pairs (0-based, ``[n_pairs, 2]``) whose coreference is counted from now on
        (``coreference_probs``, R/coreference_probs.R:54-82)."""
        self._pairs = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32).reshape(-1, 2))
        self._check(self.lib.exb200_pe_set_pairs(self.h, self._pairs.ctypes.data_as(self._capi.c_i32p), len(self._pairs)))

    def coreference_counts(self):
        """``(counts, n_samples)`` for the pairs of :meth:`set_pairs`."""
        counts = np.zeros(len(self._pairs), np.int32)
        ns = self._C.c_int32()
        self._check(self.lib.exb200_pe_pair_counts(self.h, counts.ctypes.data_as(self._capi.c_i32p), self._C.byref(ns)))
        return counts.astype(np.int64), int(ns.value)

    def coreference_probs(self) -> np.ndarray:
        counts, ns = self.coreference_counts()
        return counts / float(ns)

    def finish(self, history: np.ndarray | None = None) -> dict:
        """``membership`` (records with equal values form one shared most probable cluster),
        ``probabilities`` (of each record's most probable cluster), ``first_sample``, ``tied``,
        ``n_overflow``."""
        C, capi = self._C, self._capi
        memb, prob = np.empty(self.n, np.int32), np.empty(self.n)
        first, tied = np.empty(self.n, np.int32), np.empty(self.n, np.int32)
        over = C.c_int64()
        self._check(self.lib.exb200_pe_finish(self.h, memb.ctypes.data_as(capi.c_i32p), prob.ctypes.data_as(capi.c_f64p),
                                              first.ctypes.data_as(capi.c_i32p), tied.ctypes.data_as(capi.c_i32p), C.byref(over)))
        overflowed = np.zeros(self.n, np.uint8)
        if over.value:
            # records seen in more than `slots` distinct clusters: some occurrences were dropped, so
            # their most probable cluster may be wrong - redone below when the history is at hand
            self._check(self.lib.exb200_pe_overflowed(self.h, overflowed.ctypes.data_as(C.POINTER(C.c_uint8))))
        redo = (tied != 0) | (overflowed != 0)
        if history is not None and redo.any():
            # the reference's tie-break needs the member sets: only these records' clusters are rebuilt
            # (point_estimate.cpp:65-99 restated for single records)
            history = np.asarray(history)

            def members(s, r):
                return tuple(np.flatnonzero(history[s] == history[s, r]).tolist())

            group = {int(r): ("device", int(memb[r])) for r in range(self.n) if not redo[r]}
            for r in np.flatnonzero(redo):
                seen = [members(s, r) for s in range(history.shape[0])]
                freq = Counter(seen)
                top = max(freq.values())
                cl = min(c for c, f in freq.items() if f == top)
                prob[r] = top / float(history.shape[0])
                first[r] = seen.index(cl)
                key = ("set", cl)
                for q in cl:   # a settled record whose most probable cluster is this very set
                    if not redo[q] and members(int(first[q]), q) == cl:
                        key = group[q]
                        break
                group[int(r)] = key
            heads = {}
            for r in range(self.n):
                heads.setdefault(group[r], r)
            memb = np.array([heads[group[r]] for r in range(self.n)], np.int32)
            overflowed[:] = 0
        elif over.value:
            import warnings
            warnings.warn(f"device point estimate: {int(overflowed.sum())} records were seen in more than the table's "
                          f"slots of distinct clusters ({int(over.value)} occurrences dropped); their estimates are "
                          "unreliable - pass the link history to finish() or create the estimate with more slots "
                          "(slots >= n_samples cannot overflow)", RuntimeWarning, stacklevel=2)
        return dict(membership=memb, probabilities=prob, first_sample=first, tied=tied.astype(bool), n_overflow=int(over.value),
                    overflowed=overflowed.astype(bool))

    def close(self):
        if self.h is not None:
            self.lib.exb200_pe_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

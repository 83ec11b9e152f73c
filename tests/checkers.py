"""ctypes wrappers of the two CPU checkers (test infrastructure, never imported by the
product): oracle/_ref/libexref.so (the reference's own C++, built by oracle/Makefile) and
oracle/_build/libexoracle.so (the C restatement)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from exchanger_b200.capi import CModel, CState, FlatModel, c_f64p, c_i32p

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libexref.so")
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "libexoracle.so")
NBR_SO = os.path.join(ROOT, "oracle", "_build", "libexnbr.so")


def cpu_neighbours(domain, threshold: int, include_self: bool = False, n_threads: int | None = None):
    """Thresholded-Levenshtein neighbourhoods on the host cores (oracle/exnbr.c): same result as
    ``model.compute_close_values`` / the GPU builder, for the CPU legs of bench.py.
    Returns ``(row_ptr, val_ids, probs, max_exp_factor)``."""
    if not os.path.exists(NBR_SO):
        build_checkers()
    lib = C.CDLL(NBR_SO)
    lib.exnbr_create.restype = C.c_void_p
    lib.exnbr_create.argtypes = [C.c_void_p, c_i32p, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.exnbr_nnz.restype = C.c_int64
    lib.exnbr_nnz.argtypes = [C.c_void_p]
    lib.exnbr_get.argtypes = [C.c_void_p, C.POINTER(C.c_int64), c_i32p, c_f64p, c_f64p]
    lib.exnbr_destroy.argtypes = [C.c_void_p]
    n = len(domain)
    chars = np.zeros((n, 32), np.uint8)
    lens = np.zeros(n, np.int32)
    for i, s in enumerate(domain):
        b = s.encode("latin-1", errors="replace")
        if len(b) > 31:
            raise ValueError("strings longer than 31 bytes are not supported")
        lens[i] = len(b)
        chars[i, :len(b)] = np.frombuffer(b, np.uint8)
    h = lib.exnbr_create(chars.ctypes.data_as(C.c_void_p), lens.ctypes.data_as(c_i32p), n, int(threshold),
                         int(include_self), int(n_threads or os.cpu_count() or 1))
    if not h:
        raise ValueError("exnbr_create: bad arguments")
    try:
        nnz = lib.exnbr_nnz(h)
        row_ptr, cols, probs, mef = np.zeros(n + 1, np.int64), np.zeros(nnz, np.int32), np.zeros(nnz), np.zeros(n)
        lib.exnbr_get(h, row_ptr.ctypes.data_as(C.POINTER(C.c_int64)), cols.ctypes.data_as(c_i32p),
                      probs.ctypes.data_as(c_f64p), mef.ctypes.data_as(c_f64p))
    finally:
        lib.exnbr_destroy(h)
    return row_ptr, cols, probs, mef


def build_checkers():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"], check=True)


class RefSampler:
    """The reference's own State::update() (src/state.h:60-75) behind the shim."""

    def __init__(self, model, seed=1):
        if not os.path.exists(REF_SO):
            build_checkers()
        lib = C.CDLL(REF_SO)
        lib.exref_create.restype = C.c_void_p
        lib.exref_create.argtypes = [C.POINTER(CModel)]
        lib.exref_update.argtypes = [C.c_void_p, C.c_int]
        lib.exref_update_timed.argtypes = [C.c_void_p, c_f64p]
        lib.exref_destroy.argtypes = [C.c_void_p]
        lib.exref_seed.argtypes = [C.c_uint64]
        lib.exref_n_entities.argtypes = [C.c_void_p]
        lib.exref_get_state.argtypes = [C.c_void_p, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, c_f64p,
                                        c_i32p, c_f64p]
        lib.exref_possible_links.argtypes = [C.c_void_p, C.c_int, c_i32p, C.c_int]
        for f in ("exref_loglik_existing", "exref_loglik_new", "exref_prior_existing",
                  "exref_prior_new", "exref_log_weight_entity_value", "exref_distortion_prob",
                  "exref_exp_distortion_prob"):
            getattr(lib, f).restype = C.c_double
        lib.exref_loglik_existing.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.exref_loglik_new.argtypes = [C.c_void_p, C.c_int]
        lib.exref_prior_existing.argtypes = [C.c_void_p, C.c_int]
        lib.exref_prior_new.argtypes = [C.c_void_p, C.c_int]
        lib.exref_cluster_size.argtypes = [C.c_void_p, C.c_int]
        lib.exref_update_links.argtypes = [C.c_void_p]
        lib.exref_update_distributions.argtypes = [C.c_void_p]
        lib.exref_get_phi.argtypes = [C.c_void_p, C.c_int, C.c_int, c_f64p]
        lib.exref_log_weight_entity_value.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        lib.exref_distortion_prob.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        lib.exref_exp_distortion_prob.argtypes = [C.c_void_p, C.c_int, C.c_int]
        self.lib = lib
        self.model = model
        self.flat = FlatModel(model, seed)
        lib.exref_seed(seed)
        self.h = lib.exref_create(C.byref(self.flat.c))
        if not self.h:
            raise RuntimeError("exref_create failed")
        self.N, self.A, self.F = model.n_records, model.n_attrs, model.n_files

    def update(self, n=1):
        if self.lib.exref_update(self.h, n) != 0:
            raise RuntimeError("reference threw")

    def update_timed(self, n=1):
        secs = np.zeros(6)
        for _ in range(n):
            self.lib.exref_update_timed(self.h, secs.ctypes.data_as(c_f64p))
        return secs

    def n_entities(self):
        return self.lib.exref_n_entities(self.h)

    def state(self):
        N, A, F = self.N, self.A, self.F
        links = np.zeros(N, np.int32)
        n_ent = C.c_int32()
        ent_ids = np.zeros(N, np.int32)
        ent_attrs = np.zeros((N, A), np.int32)
        rec_dist = np.zeros((N, A), np.int32)
        theta = np.zeros((A, F))
        counts = np.zeros((A, F), np.int32)
        clust = np.zeros(2)
        self.lib.exref_get_state(self.h, links.ctypes.data_as(c_i32p), C.byref(n_ent),
                                 ent_ids.ctypes.data_as(c_i32p), ent_attrs.ctypes.data_as(c_i32p),
                                 rec_dist.ctypes.data_as(c_i32p), theta.ctypes.data_as(c_f64p),
                                 counts.ctypes.data_as(c_i32p), clust.ctypes.data_as(c_f64p))
        K = n_ent.value
        return dict(links=links, ent_ids=ent_ids[:K].copy(), ent_attrs=ent_attrs[:K].T.copy(),
                    rec_distortions=rec_dist.T.copy(), distort_probs=theta.T.copy(),
                    distort_counts=counts.T.copy(), clust=clust)

    def possible_links(self, rid, cap=4096):
        out = np.zeros(cap, np.int32)
        n = self.lib.exref_possible_links(self.h, rid, out.ctypes.data_as(c_i32p), cap)
        return out[:min(n, cap)].copy()

    def __del__(self):
        try:
            self.lib.exref_destroy(self.h)
        except Exception:
            pass


class OracleSampler:
    """The C restatement (oracle/exoracle.c)."""

    def __init__(self, model, seed=1):
        if not os.path.exists(ORACLE_SO):
            build_checkers()
        lib = C.CDLL(ORACLE_SO)
        lib.exo_create.restype = C.c_void_p
        lib.exo_create.argtypes = [C.POINTER(CModel)]
        lib.exo_destroy.argtypes = [C.c_void_p]
        lib.exo_sweeps.argtypes = [C.c_void_p, C.c_int]
        lib.exo_sweep_phases.argtypes = [C.c_void_p, C.c_int]
        lib.exo_set_brute.argtypes = [C.c_void_p, C.c_int]
        lib.exo_set_wide_threshold.argtypes = [C.c_void_p, C.c_int]
        lib.exo_n_wide.argtypes = [C.c_void_p]
        lib.exo_get_state.argtypes = [C.c_void_p, C.POINTER(CState)]
        lib.exo_get_stats.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), c_i32p, c_i32p]
        lib.exo_last_error.restype = C.c_char_p
        lib.exo_last_error.argtypes = [C.c_void_p]
        lib.exo_link_conditional.argtypes = [C.c_void_p, C.c_int, C.c_int, c_i32p, c_f64p, c_f64p]
        lib.exo_log_weight_entity_value.restype = C.c_double
        lib.exo_log_weight_entity_value.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        lib.exo_get_phi.argtypes = [C.c_void_p, C.c_int, c_f64p]
        lib.exo_update_distributions.argtypes = [C.c_void_p]
        self.lib = lib
        self.model = model
        self.flat = FlatModel(model, seed)
        self.h = lib.exo_create(C.byref(self.flat.c))
        self.N, self.A, self.F = model.n_records, model.n_attrs, model.n_files
        err = lib.exo_last_error(self.h)
        if err:
            raise RuntimeError(err.decode())

    def sweeps(self, n=1):
        rc = self.lib.exo_sweeps(self.h, n)
        if rc:
            raise RuntimeError(self.lib.exo_last_error(self.h).decode())

    def phases(self, mask):
        rc = self.lib.exo_sweep_phases(self.h, mask)
        if rc:
            raise RuntimeError(self.lib.exo_last_error(self.h).decode())

    def set_brute(self, on=True):
        self.lib.exo_set_brute(self.h, int(on))

    def set_wide_threshold(self, w):
        self.lib.exo_set_wide_threshold(self.h, int(w))

    def n_wide(self):
        return self.lib.exo_n_wide(self.h)

    def state(self):
        return read_state(lambda st: self.lib.exo_get_state(self.h, C.byref(st)), self.N, self.A, self.F)

    def stats(self):
        nc, ch, k = C.c_longlong(), C.c_int32(), C.c_int32()
        self.lib.exo_get_stats(self.h, C.byref(nc), C.byref(ch), C.byref(k))
        return dict(n_candidates=nc.value, n_links_changed=ch.value, n_entities=k.value)

    def link_conditional(self, rid, cap=65536):
        ids = np.zeros(cap, np.int32)
        lw = np.zeros(cap)
        lwn = C.c_double()
        n = self.lib.exo_link_conditional(self.h, rid, cap, ids.ctypes.data_as(c_i32p),
                                          lw.ctypes.data_as(c_f64p), C.byref(lwn))
        return ids[:n].copy(), lw[:n].copy(), lwn.value

    def log_weight_entity_value(self, eid, a, v):
        return self.lib.exo_log_weight_entity_value(self.h, eid, a, v)

    def __del__(self):
        try:
            self.lib.exo_destroy(self.h)
        except Exception:
            pass


def read_state(fill, N, A, F):
    """Allocate an exb200_state, let `fill(state)` populate it, return numpy views in the
    model's orientation ([A, K], [A, N], [F, A])."""
    st = CState()
    links = np.zeros(N, np.int32)
    ent_ids = np.zeros(N, np.int32)
    ent_attrs = np.zeros((N, A), np.int32)
    rec_dist = np.zeros((N, A), np.int32)
    theta = np.zeros((A, F))
    counts = np.zeros((A, F), np.int32)
    concs = np.zeros(A)
    st.links = links.ctypes.data_as(c_i32p)
    st.ent_ids = ent_ids.ctypes.data_as(c_i32p)
    st.ent_attrs = ent_attrs.ctypes.data_as(c_i32p)
    st.rec_distortions = rec_dist.ctypes.data_as(c_i32p)
    st.distort_probs = theta.ctypes.data_as(c_f64p)
    st.distort_counts = counts.ctypes.data_as(c_i32p)
    st.distort_dist_concs = concs.ctypes.data_as(c_f64p)
    fill(st)
    K = st.n_entities
    cl = st.clust
    return dict(iteration=st.iteration, links=links, ent_ids=ent_ids[:K].copy(),
                ent_attrs=ent_attrs[:K].T.copy(), rec_distortions=rec_dist.T.copy(),
                distort_probs=theta.T.copy(), distort_counts=counts.T.copy(),
                distort_dist_concs=concs,
                clust=dict(alpha=cl.alpha, d=cl.d, m=cl.m, kappa=cl.kappa, kind=cl.kind))

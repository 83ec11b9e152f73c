"""ctypes mirror of ``include/exb200.h`` and the flattening of an
:class:`~exchanger_b200.model.ExchangERModel` into the POD structs the C ABI takes.

This is the Python counterpart of the R glue shown in INTEGRATION.md: it does what
``readS4State`` (src/sample.cpp:71-105) does on the way in and what
``buildFinalS4State`` (src/sample.cpp:14-63) does on the way out, nothing more.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

from .model import BetaRV, DirichletRV, ExchangERModel, GammaRV, GeneralizedCouponRP, \
    PitmanYorRP, ShiftedNegBinomRV

EXB200_OK = 0
EXB200_ERR_WEIGHTS, EXB200_ERR_ABORTED = -5, -6
ERR_NAMES = {0: "OK", -1: "INVALID", -2: "UNSUPPORTED", -3: "NO_DEVICE", -4: "CUDA", -5: "WEIGHTS",
             -6: "ABORTED", -7: "OOM"}
CLUST_PITMAN_YOR, CLUST_GEN_COUPON, CLUST_COUPON = 0, 1, 2
ENTDIST_FIXED, ENTDIST_DIRICHLET = 0, 1

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_f64p = C.POINTER(C.c_double)
# callbacks of exb200_run: int abort_cb(void* ctx), void progress_cb(void* ctx, int32_t completed)
ABORT_CB = C.CFUNCTYPE(C.c_int, C.c_void_p)
PROGRESS_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int32)


class CAttr(C.Structure):
    _fields_ = [
        ("domain_size", C.c_int32), ("is_constant", C.c_int32),
        ("exclude_entity_value", C.c_int32), ("entity_dist_kind", C.c_int32),
        ("probs", c_f64p), ("entity_dist", c_f64p),
        ("close_row_ptr", c_i64p), ("close_val_ids", c_i32p), ("close_probs", c_f64p),
        ("max_exp_factor", c_f64p),
        ("distort_prob_shape1", C.c_double), ("distort_prob_shape2", C.c_double),
        ("distort_dist_conc", C.c_double),
        ("conc_prior_shape", C.c_double), ("conc_prior_rate", C.c_double),
    ]


class CClust(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("alpha", C.c_double), ("d", C.c_double), ("m", C.c_double), ("kappa", C.c_double),
        ("alpha_prior_shape", C.c_double), ("alpha_prior_rate", C.c_double),
        ("d_prior_shape1", C.c_double), ("d_prior_shape2", C.c_double),
        ("kappa_prior_shape", C.c_double), ("kappa_prior_rate", C.c_double),
        ("m_prior_size", C.c_double), ("m_prior_prob", C.c_double),
    ]


class CModel(C.Structure):
    _fields_ = [
        ("n_records", C.c_int32), ("n_attrs", C.c_int32), ("n_files", C.c_int32),
        ("n_entities", C.c_int32), ("iteration", C.c_int32),
        ("rec_attrs", c_i32p), ("rec_distortions", c_i32p), ("file_ids", c_i32p),
        ("links", c_i32p), ("ent_ids", c_i32p), ("ent_attrs", c_i32p),
        ("distort_probs", c_f64p), ("attrs", C.POINTER(CAttr)),
        ("clust", CClust), ("seed", C.c_uint64),
    ]


class CHistory(C.Structure):
    _fields_ = [
        ("links", c_i32p), ("distort_probs", c_f64p), ("distort_counts", c_i32p),
        ("distort_dist_concs", c_f64p), ("n_linked_ents", c_i32p), ("clust_params", c_f64p),
        ("n_saved", C.c_int32),
    ]


class CState(C.Structure):
    _fields_ = [
        ("iteration", C.c_int32), ("n_entities", C.c_int32),
        ("links", c_i32p), ("ent_ids", c_i32p), ("ent_attrs", c_i32p),
        ("rec_distortions", c_i32p), ("distort_probs", c_f64p), ("distort_counts", c_i32p),
        ("distort_dist_concs", c_f64p), ("clust", CClust),
    ]


class CStats(C.Structure):
    _fields_ = [
        ("n_candidates", C.c_int64), ("n_bucket_items", C.c_int64),
        ("n_rounds", C.c_int32), ("n_wide", C.c_int32), ("n_entities", C.c_int32),
        ("n_links_changed", C.c_int32), ("n_tables", C.c_int32), ("n_long", C.c_int32),
        ("ms_clust", C.c_float), ("ms_prep", C.c_float), ("ms_index", C.c_float),
        ("ms_cand", C.c_float), ("ms_rounds", C.c_float), ("ms_canon", C.c_float),
        ("ms_entities", C.c_float), ("ms_distort", C.c_float), ("ms_total", C.c_float),
        ("ms_cand_long", C.c_float), ("ms_wide", C.c_float), ("ms_index_count", C.c_float),
        ("ms_index_fill", C.c_float), ("n_probe", C.c_int32), ("n_huge", C.c_int32),
        ("ms_cand_fast", C.c_float), ("n_slow", C.c_int32), ("n_listed", C.c_int32), ("n_scanned", C.c_int32),
        ("ms_tail", C.c_float), ("ms_entities_single", C.c_float),
    ]


def _ptr(arr, typ):
    return None if arr is None else arr.ctypes.data_as(typ)


def clust_to_c(params, prior) -> CClust:
    """clust_params / clust_prior slots -> exb200_clust (what read_clust_params reads,
    src/clust_params.cpp:259-281)."""
    c = CClust()
    if isinstance(prior, PitmanYorRP):
        c.kind = CLUST_PITMAN_YOR
        c.alpha, c.d = float(params.alpha), float(params.d)
        if isinstance(prior.alpha, GammaRV):
            c.alpha_prior_shape, c.alpha_prior_rate = prior.alpha.shape, prior.alpha.rate
        if isinstance(prior.d, BetaRV):
            c.d_prior_shape1, c.d_prior_shape2 = prior.d.shape1, prior.d.shape2
    elif isinstance(prior, GeneralizedCouponRP):
        inf_kappa = not isinstance(prior.kappa, GammaRV) and math.isinf(prior.kappa)
        c.kind = CLUST_COUPON if inf_kappa else CLUST_GEN_COUPON
        c.m, c.kappa = float(params.m), float(params.kappa)
        if isinstance(prior.kappa, GammaRV):
            c.kappa_prior_shape, c.kappa_prior_rate = prior.kappa.shape, prior.kappa.rate
        if isinstance(prior.m, ShiftedNegBinomRV):
            c.m_prior_size, c.m_prior_prob = prior.m.size, prior.m.prob
    else:
        raise ValueError("Unrecognized clust_prior")
    return c


class FlatModel:
    """Owns the contiguous numpy buffers behind a :class:`CModel` (the library copies
    them during ``exb200_create``; they only have to outlive that call)."""

    def __init__(self, model: ExchangERModel, seed: int):
        A, N = model.rec_attrs.shape
        self.keep = []

        def own(a, dt):
            a = np.ascontiguousarray(a, dtype=dt)
            self.keep.append(a)
            return a

        # R matrices are column-major: [A x N] with a fastest == C-order [N, A]
        self.rec_attrs = own(model.rec_attrs.T, np.int32)
        self.rec_distortions = own(model.rec_distortions.T, np.int32)
        self.file_ids = own(model.file_ids, np.int32)
        self.links = own(model.links, np.int32)
        self.ent_ids = own(model.ent_ids, np.int32)
        self.ent_attrs = own(model.ent_attrs.T, np.int32)
        self.distort_probs = own(model.distort_probs.T, np.float64)  # [F x A] col-major
        self.attrs = (CAttr * A)()
        for a, (spec, idx) in enumerate(zip(model.attr_params, model.attr_indices)):
            ca = self.attrs[a]
            D = len(idx.probs)
            ca.domain_size = D
            ca.is_constant = int(idx.is_constant)
            ca.exclude_entity_value = int(spec.exclude_entity_value)
            ca.probs = _ptr(own(idx.probs, np.float64), c_f64p)
            prior = spec.entity_dist_prior
            if isinstance(prior, DirichletRV):
                ca.entity_dist_kind = ENTDIST_DIRICHLET
                ca.entity_dist = _ptr(own(np.full(D, prior.alpha), np.float64), c_f64p)
            elif prior is not None:
                ca.entity_dist_kind = ENTDIST_FIXED
                pmf = np.array([prior[k] for k in idx.domain], dtype=np.float64)
                ca.entity_dist = _ptr(own(pmf, np.float64), c_f64p)
            if not idx.is_constant:
                ca.close_row_ptr = _ptr(own(idx.close_row_ptr, np.int64), c_i64p)
                ca.close_val_ids = _ptr(own(idx.close_val_ids, np.int32), c_i32p)
                ca.close_probs = _ptr(own(idx.close_probs, np.float64), c_f64p)
                ca.max_exp_factor = _ptr(own(idx.max_exp_factor, np.float64), c_f64p)
            p = spec.distort_prob_prior
            if isinstance(p, BetaRV):
                ca.distort_prob_shape1, ca.distort_prob_shape2 = p.shape1, p.shape2
            ca.distort_dist_conc = float(model.distort_dist_concs[a])
            al = spec.distort_dist_prior.alpha
            if isinstance(al, GammaRV):
                ca.conc_prior_shape, ca.conc_prior_rate = al.shape, al.rate
        m = CModel()
        m.n_records, m.n_attrs, m.n_files = N, A, model.n_files
        m.n_entities = len(model.ent_ids)
        m.iteration = int(model.iteration)
        m.rec_attrs = _ptr(self.rec_attrs, c_i32p)
        m.rec_distortions = _ptr(self.rec_distortions, c_i32p)
        m.file_ids = _ptr(self.file_ids, c_i32p)
        m.links = _ptr(self.links, c_i32p)
        m.ent_ids = _ptr(self.ent_ids, c_i32p)
        m.ent_attrs = _ptr(self.ent_attrs, c_i32p)
        m.distort_probs = _ptr(self.distort_probs, c_f64p)
        m.attrs = self.attrs
        m.clust = clust_to_c(model.clust_params, model.clust_prior)
        m.seed = C.c_uint64(seed & 0xFFFFFFFFFFFFFFFF)
        self.c = m

    def nbytes(self) -> int:
        return int(sum(a.nbytes for a in self.keep))


_LIB = None
_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libexb200.so")


class Exb200Error(RuntimeError):
    """Raised for every non-OK status of the C ABI (the reference raises R errors via
    Rcpp::stop / BEGIN_RCPP, src/RcppExports.cpp:38-47)."""

    def __init__(self, status: int, message: str):
        super().__init__(f"exb200 {ERR_NAMES.get(status, status)}: {message}")
        self.status = status


def load_library(path: str | None = None):
    """dlopen the CUDA library.  There is no fallback: a missing library is an error."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    p = path or _LIB_PATH
    if not os.path.exists(p):
        raise Exb200Error(-3, f"{p} not found; build it with `python -c 'import __graft_entry__ as g; "
                              "g.build()'` (no CPU fallback exists)")
    lib = C.CDLL(p)
    H = C.c_void_p
    lib.exb200_abi_version.restype = C.c_int
    lib.exb200_trim.restype = None
    lib.exb200_trim.argtypes = []
    lib.exb200_device_count.restype = C.c_int
    lib.exb200_create.argtypes = [C.POINTER(CModel), C.c_int, C.POINTER(H)]
    lib.exb200_run.argtypes = [H, C.c_int32, C.c_int32, C.c_int32, C.POINTER(CHistory),
                               C.c_void_p, C.c_void_p, C.c_void_p]
    lib.exb200_sweeps.argtypes = [H, C.c_int32]
    lib.exb200_sweeps_timed.argtypes = [H, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_int64)]
    lib.exb200_get_state.argtypes = [H, C.POINTER(CState)]
    lib.exb200_get_stats.argtypes = [H, C.POINTER(CStats)]
    lib.exb200_link_conditional.argtypes = [H, C.c_int32, C.c_int32, c_i32p, c_f64p, c_f64p]
    lib.exb200_set_option.argtypes = [H, C.c_char_p, C.c_int64]
    lib.exb200_destroy.argtypes = [H]
    lib.exb200_destroy.restype = None
    lib.exb200_last_error.argtypes = [H]
    lib.exb200_last_error.restype = C.c_char_p
    lib.exb200_shard_links.argtypes = [H]
    lib.exb200_shard_entities.argtypes = [H, C.c_int32, C.c_int32]
    lib.exb200_shard_distort.argtypes = [H, C.c_int32, C.c_int32, c_i32p, c_i32p]
    lib.exb200_shard_finish.argtypes = [H, c_i32p, c_i32p]
    lib.exb200_device_buffer.argtypes = [H, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
    lib.exb200_stream_sync.argtypes = [H]
    lib.exb200_pe_create.argtypes = [C.c_int32, C.c_int32, C.c_int, C.POINTER(H)]
    lib.exb200_pe_add.argtypes = [H, c_i32p]
    lib.exb200_pe_add_chain.argtypes = [H, H]
    lib.exb200_pe_finish.argtypes = [H, c_i32p, c_f64p, c_i32p, c_i32p, c_i64p]
    lib.exb200_pe_overflowed.argtypes = [H, C.POINTER(C.c_uint8)]
    lib.exb200_pe_destroy.argtypes = [H]
    lib.exb200_pe_destroy.restype = None
    lib.exb200_attach_point_estimate.argtypes = [H, H]
    lib.exb200_pe_set_pairs.argtypes = [H, c_i32p, C.c_int64]
    lib.exb200_pe_pair_counts.argtypes = [H, c_i32p, c_i32p]
    lib.exb200_neighbours_create.argtypes = [C.c_void_p, c_i32p, C.c_int32, C.c_int32, C.c_int32, C.c_int,
                                             C.POINTER(H)]
    lib.exb200_neighbours_nnz.argtypes = [H]
    lib.exb200_neighbours_nnz.restype = C.c_int64
    lib.exb200_neighbours_get.argtypes = [H, c_i64p, c_i32p, c_f64p, c_f64p]
    lib.exb200_neighbours_destroy.argtypes = [H]
    lib.exb200_neighbours_destroy.restype = None
    if path is None:
        _LIB = lib
    return lib


def trim() -> None:
    """Give the device slabs and pinned rows that destroyed samplers left in the library's cache back to
    the driver (``exb200_trim``)."""
    load_library().exb200_trim()


def levenshtein_neighbours(domain, threshold: int, include_self: bool = False, device: int = 0):
    """Thresholded-Levenshtein neighbourhoods of a domain of (ASCII) strings on the GPU: the
    ``close_values`` / ``max_exp_factor`` slots that ``compute_close_values`` builds
    (R/AttributeIndex.R:117-137).  Returns ``(row_ptr, val_ids, probs, max_exp_factor)``."""
    lib = load_library()
    n = len(domain)
    chars = np.zeros((n, 32), dtype=np.uint8)
    lens = np.zeros(n, dtype=np.int32)
    for i, s in enumerate(domain):
        b = s.encode("latin-1", errors="replace")
        if len(b) > 31:
            raise ValueError("strings longer than 31 bytes are not supported by the GPU neighbourhood builder")
        lens[i] = len(b)
        chars[i, :len(b)] = np.frombuffer(b, dtype=np.uint8)
    h = C.c_void_p()
    rc = lib.exb200_neighbours_create(chars.ctypes.data_as(C.c_void_p), lens.ctypes.data_as(c_i32p), n,
                                      int(threshold), int(include_self), device, C.byref(h))
    if rc != EXB200_OK:
        raise Exb200Error(rc, lib.exb200_last_error(None).decode())
    try:
        nnz = lib.exb200_neighbours_nnz(h)
        row_ptr = np.zeros(n + 1, np.int64)
        cols = np.zeros(nnz, np.int32)
        probs = np.zeros(nnz)
        mef = np.zeros(n)
        lib.exb200_neighbours_get(h, row_ptr.ctypes.data_as(c_i64p), cols.ctypes.data_as(c_i32p),
                                  probs.ctypes.data_as(c_f64p), mef.ctypes.data_as(c_f64p))
    finally:
        lib.exb200_neighbours_destroy(h)
    return row_ptr, cols, probs, mef


EXPORTED_SYMBOLS = [
    "exb200_abi_version", "exb200_trim", "exb200_device_count", "exb200_create", "exb200_run", "exb200_sweeps", "exb200_sweeps_timed",
    "exb200_get_state", "exb200_get_stats", "exb200_link_conditional", "exb200_set_option",
    "exb200_destroy", "exb200_last_error", "exb200_neighbours_create", "exb200_neighbours_nnz",
    "exb200_neighbours_get", "exb200_neighbours_destroy",
    "exb200_shard_links", "exb200_shard_entities", "exb200_shard_distort", "exb200_shard_finish",
    "exb200_device_buffer", "exb200_stream_sync",
    "exb200_pe_create", "exb200_pe_add", "exb200_pe_add_chain", "exb200_pe_finish", "exb200_pe_destroy",
    "exb200_attach_point_estimate", "exb200_pe_set_pairs", "exb200_pe_pair_counts", "exb200_pe_overflowed",
]

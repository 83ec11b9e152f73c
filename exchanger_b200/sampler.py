"""``run_inference`` / ``ExchangERFit`` over the C ABI - the Python counterpart of the R entry
points (R/run_inference.R:60-115, R/ExchangERFit.R, src/sample.cpp:118-219).

All sampling happens in ``libexb200.so`` (hand-written CUDA for sm_100a).  There is no CPU
path: without the library or without a GPU every call raises :class:`Exb200Error`.
"""
from __future__ import annotations

import time
import ctypes as C
import warnings
from dataclasses import dataclass, field

import numpy as np

from . import capi
from .capi import CHistory, CState, CStats, Exb200Error, FlatModel, c_f64p, c_i32p
from .model import ExchangERModel, GeneralizedCouponRP, PitmanYorRP, _is_rv, model_with_state


class Sampler:
    """A device-resident chain: thin wrapper of ``exb200_handle``."""

    def __init__(self, model: ExchangERModel, seed: int = 0, device: int = 0):
        self.lib = capi.load_library()
        self.model = model
        self.N, self.A, self.F = model.n_records, model.n_attrs, model.n_files
        t0 = time.perf_counter()
        self.flat = FlatModel(model, seed)
        self.t_flat = time.perf_counter() - t0   # host-side flattening, for the end-to-end breakdown
        self.h = C.c_void_p()
        rc = self.lib.exb200_create(C.byref(self.flat.c), device, C.byref(self.h))
        if rc != capi.EXB200_OK:
            self.h = None
            raise Exb200Error(rc, self.lib.exb200_last_error(None).decode())

    def _check(self, rc):
        if rc != capi.EXB200_OK:
            raise Exb200Error(rc, self.lib.exb200_last_error(self.h).decode())

    def set_option(self, name: str, value: int):
        self._check(self.lib.exb200_set_option(self.h, name.encode(), int(value)))

    def sweeps(self, n: int = 1):
        """``n`` applications of State::update() (src/state.h:60-75), nothing copied back."""
        self._check(self.lib.exb200_sweeps(self.h, int(n)))

    def sweeps_timed(self, n: int):
        """``n`` sweeps; returns (device milliseconds from CUDA events on the library's stream,
        kernels launched)."""
        ms, nl = C.c_float(), C.c_int64()
        self._check(self.lib.exb200_sweeps_timed(self.h, int(n), C.byref(ms), C.byref(nl)))
        return ms.value, nl.value

    def run(self, n_samples: int, thin_interval: int = 1, burnin_interval: int = 0, *,
            keep_links: bool = True, point_estimate=None, abort=None, progress=None) -> dict:
        """The sampling loop of ``sample()`` (src/sample.cpp:171-202) into fresh host arrays.
        ``point_estimate`` (an ``analysis.DevicePointEstimate``) receives every saved sample on the
        device; with ``keep_links=False`` the link vectors then never leave the GPU (``links`` of the
        history is empty).

        ``abort()`` is polled after every sweep and plays ``Progress::check_abort`` (sample.cpp:200):
        when it returns true the loop stops and - like the reference, which breaks out of its loop and
        returns what it has - the history comes back with ``n_saved`` rows filled (the others zero)
        and ``aborted=True``.  ``progress(completed_sweeps)`` plays ``p.increment()`` (sample.cpp:201)."""
        N, FA, A = self.N, self.F * self.A, self.A
        self._check(self.lib.exb200_attach_point_estimate(self.h, point_estimate.h if point_estimate is not None else None))
        hist = dict(
            links=np.zeros((n_samples if keep_links else 0, N), np.int32),
            distort_probs=np.zeros((n_samples, FA)),
            distort_counts=np.zeros((n_samples, FA), np.int32),
            distort_dist_concs=np.zeros((n_samples, A)),
            n_linked_ents=np.zeros((n_samples, 1), np.int32),
            clust_params=np.zeros((n_samples, 2)),
        )
        ch = CHistory()
        ch.links = hist["links"].ctypes.data_as(c_i32p) if keep_links else None
        ch.distort_probs = hist["distort_probs"].ctypes.data_as(c_f64p)
        ch.distort_counts = hist["distort_counts"].ctypes.data_as(c_i32p)
        ch.distort_dist_concs = hist["distort_dist_concs"].ctypes.data_as(c_f64p)
        ch.n_linked_ents = hist["n_linked_ents"].ctypes.data_as(c_i32p)
        ch.clust_params = hist["clust_params"].ctypes.data_as(c_f64p)
        abort_c = capi.ABORT_CB(lambda _ctx: 1 if abort() else 0) if abort is not None else None
        progress_c = capi.PROGRESS_CB(lambda _ctx, done: progress(int(done))) if progress is not None else None
        rc = self.lib.exb200_run(self.h, n_samples, thin_interval, burnin_interval, C.byref(ch),
                                 C.cast(abort_c, C.c_void_p) if abort_c else None,
                                 C.cast(progress_c, C.c_void_p) if progress_c else None, None)
        self.last_n_saved = int(ch.n_saved)
        if rc == capi.EXB200_ERR_ABORTED:   # partial history, as the reference returns after an interrupt
            hist["aborted"], hist["n_saved"] = True, int(ch.n_saved)
        else:
            self._check(rc)
        # keep only the random cluster parameters, named as ClustParams::to_R_vec does
        # (src/clust_params.cpp:161-216)
        prior = self.model.clust_prior
        if isinstance(prior, PitmanYorRP):
            keep = [("alpha", 0, _is_rv(prior.alpha)), ("d", 1, _is_rv(prior.d))]
        else:
            keep = [("kappa", 0, _is_rv(prior.kappa)), ("m", 1, _is_rv(prior.m))]
        cols = [(n, i) for n, i, random in keep if random]
        if cols:
            hist["clust_params"] = hist["clust_params"][:, [i for _, i in cols]]
            hist["clust_params_names"] = [n for n, _ in cols]
        else:
            del hist["clust_params"]
        return hist

    def state(self) -> dict:
        from_c = CState()
        N, A, F = self.N, self.A, self.F
        links = np.empty(N, np.int32)
        ent_ids = np.empty(N, np.int32)
        ent_attrs = np.empty((N, A), np.int32)
        rec_dist = np.empty((N, A), np.int32)
        theta = np.zeros((A, F))
        counts = np.zeros((A, F), np.int32)
        concs = np.zeros(A)
        from_c.links = links.ctypes.data_as(c_i32p)
        from_c.ent_ids = ent_ids.ctypes.data_as(c_i32p)
        from_c.ent_attrs = ent_attrs.ctypes.data_as(c_i32p)
        from_c.rec_distortions = rec_dist.ctypes.data_as(c_i32p)
        from_c.distort_probs = theta.ctypes.data_as(c_f64p)
        from_c.distort_counts = counts.ctypes.data_as(c_i32p)
        from_c.distort_dist_concs = concs.ctypes.data_as(c_f64p)
        self._check(self.lib.exb200_get_state(self.h, C.byref(from_c)))
        K = from_c.n_entities
        cl = from_c.clust
        # [A, K] / [A, N] matrices are column-major views of what the library wrote (R's layout)
        return dict(iteration=from_c.iteration, links=links, ent_ids=ent_ids[:K],
                    ent_attrs=ent_attrs[:K].T, rec_distortions=rec_dist.T,
                    distort_probs=theta.T.copy(), distort_counts=counts.T.copy(),
                    distort_dist_concs=concs,
                    clust=dict(alpha=cl.alpha, d=cl.d, m=cl.m, kappa=cl.kappa, kind=cl.kind))

    def stats(self) -> dict:
        s = CStats()
        self._check(self.lib.exb200_get_stats(self.h, C.byref(s)))
        return {name: getattr(s, name) for name, _ in CStats._fields_}

    def link_conditional(self, rid: int, cap: int = 65536):
        ids = np.zeros(cap, np.int32)
        lw = np.zeros(cap)
        lwn = C.c_double()
        n = self.lib.exb200_link_conditional(self.h, rid, cap, ids.ctypes.data_as(c_i32p),
                                             lw.ctypes.data_as(c_f64p), C.byref(lwn))
        if n < 0:
            self._check(n)
        return ids[:n].copy(), lw[:n].copy(), lwn.value

    def close(self):
        if getattr(self, "h", None):
            self.lib.exb200_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


@dataclass
class ExchangERFit:
    """R/ExchangERFit.R: ``history`` (matrices with one row per saved sample) and the final
    ``state`` (an :class:`ExchangERModel` the chain can be resumed from).  ``mcpar`` is
    ``(start_iter, end_iter, thin_interval)`` as in ``attr(history, 'mcpar')``."""

    history: dict
    state: ExchangERModel
    mcpar: tuple = field(default=(0, 0, 1))
    point_estimate: dict | None = None   # run_inference(point_estimate=True): DevicePointEstimate.finish()


def _is_integer(x) -> bool:
    return float(x) == int(x)


def run_inference(x, n_samples: int, thin_interval: int = 1, burnin_interval: int = 0, *,
                  seed: int = 0, device: int = 0, point_estimate: bool = False, keep_links: bool = True) -> ExchangERFit:
    """Mirror of ``run_inference`` (R/run_inference.R:60-115).

    ``x`` is an :class:`ExchangERModel` (new chain, or a chain resumed from ``fit.state``) or an
    :class:`ExchangERFit` (resume and concatenate the histories, :104-115).  ``seed`` plays the
    role of ``set.seed()``: it is the Philox key of the chain; the absolute iteration is part of
    every counter, so a resumed chain continues the same stream.

    ``point_estimate=True`` also accumulates ``mp_clusters`` / ``smp_clusters`` on the device while the
    chain runs (``fit.point_estimate``: ``analysis.DevicePointEstimate.finish()``); with
    ``keep_links=False`` the ``n_samples x N`` link history is then not kept at all.
    """
    if isinstance(x, ExchangERFit):
        if burnin_interval != 0:
            warnings.warn(f"ignoring burnin_interval={burnin_interval} when resuming")
        old_thin = x.mcpar[2]
        if thin_interval != old_thin:
            warnings.warn(f"using thin_interval={old_thin} from previous run")
        new = run_inference(x.state, n_samples, thin_interval, burnin_interval, seed=seed, device=device)
        return combine_results(x, new)
    for name, v in (("n_samples", n_samples), ("thin_interval", thin_interval),
                    ("burnin_interval", burnin_interval)):
        if np.ndim(v) != 0:
            raise ValueError(f"{name} must be a scalar")
    if n_samples < 0:
        raise ValueError("n_samples must be positive")
    if thin_interval < 1:
        raise ValueError("thin_interval must be positive")
    if burnin_interval < 0:
        raise ValueError("burnin_interval must be non-negative")
    for name, v in (("n_samples", n_samples), ("thin_interval", thin_interval),
                    ("burnin_interval", burnin_interval)):
        if not _is_integer(v):
            raise ValueError(f"{name} must be an integer")
    n_samples, thin_interval, burnin_interval = int(n_samples), int(thin_interval), int(burnin_interval)
    if x.iteration == 0:
        start_iter = burnin_interval if burnin_interval else thin_interval
    else:
        start_iter = x.iteration + thin_interval
    pe_result = None
    with Sampler(x, seed=seed, device=device) as s:
        if point_estimate:
            from .analysis import DevicePointEstimate
            # a record can be seen in at most n_samples distinct clusters: with that many slots (up to
            # 32) the per-record tables cannot overflow; beyond that overflowed records are redone from
            # the history, or reported by a RuntimeWarning when keep_links=False leaves none
            with DevicePointEstimate(s.N, slots=int(min(max(n_samples, 8), 32)), device=device) as pe:
                history = s.run(n_samples, thin_interval, burnin_interval, keep_links=keep_links, point_estimate=pe)
                s._check(s.lib.exb200_attach_point_estimate(s.h, None))
                pe_result = pe.finish(history["links"] if keep_links and n_samples else None) if n_samples else None
        else:
            history = s.run(n_samples, thin_interval, burnin_interval, keep_links=keep_links)
        final = model_with_state(x, s.state())
    fit = ExchangERFit(history=history, state=final, mcpar=(start_iter, final.iteration, thin_interval))
    fit.point_estimate = pe_result
    return fit


def combine_results(a: ExchangERFit, b: ExchangERFit) -> ExchangERFit:
    """R/ExchangERFit.R:84-119."""
    if a.mcpar[1] - a.mcpar[0] == 0 and len(a.history["links"]) == 0:
        raise ValueError("resultA has no history")
    if len(b.history["links"]) == 0:
        raise ValueError("resultB has no history")
    if a.mcpar[1] >= b.mcpar[0]:
        raise ValueError("resultB must occur chronologically after resultA")
    if a.mcpar[2] != b.mcpar[2]:
        raise ValueError("results must have the same thin_interval")
    if set(a.history) != set(b.history):
        raise ValueError("results must have the same history variables")
    hist = {}
    for k, va in a.history.items():
        hist[k] = va if isinstance(va, list) else np.concatenate([va, b.history[k]], axis=0)
    return ExchangERFit(history=hist, state=b.state, mcpar=(a.mcpar[0], b.state.iteration, a.mcpar[2]))

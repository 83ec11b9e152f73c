"""Host-side model preparation: the Python mirror of the reference's R layer.

The reference prepares an ``ExchangERModel`` in R (``exchanger()``,
R/ExchangERModel.R:239-351) from ``Attribute`` specs (R/Attribute.R:112-149), a
random-partition prior (R/PitmanYorRP.R, R/GeneralizedCouponRP.R) and per-attribute
``AttributeIndex`` objects (R/AttributeIndex.R:68-137).  R is not available in the
build image, so the same constructors are mirrored here with the same names,
argument meaning and validation errors; the result is a plain container of numpy
arrays whose fields are exactly the S4 slots that ``.sample`` reads
(src/sample.cpp:71-105) and that the C ABI (include/exb200.h) takes.

Nothing in this module does sampling.  It runs once per model.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Mapping, Sequence

import numpy as np

NA_INT = np.int32(-2147483648)

# --------------------------------------------------------------------------- RVs


@dataclass(frozen=True)
class BetaRV:
    """R/BetaRV.R: Beta(shape1, shape2)."""

    shape1: float
    shape2: float

    def __post_init__(self):
        if not (self.shape1 > 0 and self.shape2 > 0):
            raise ValueError("shape1 and shape2 must be positive")

    def mean(self) -> float:
        return self.shape1 / (self.shape1 + self.shape2)


@dataclass(frozen=True)
class GammaRV:
    """R/GammaRV.R: Gamma(shape, rate)."""

    shape: float
    rate: float

    def __post_init__(self):
        if not (self.shape > 0 and self.rate > 0):
            raise ValueError("shape and rate must be positive")

    def mean(self) -> float:
        return self.shape / self.rate


@dataclass(frozen=True)
class ShiftedNegBinomRV:
    """R/ShiftedNegBinomRV.R: 1 + NegBinom(size, prob)."""

    size: float
    prob: float

    def __post_init__(self):
        if not self.size > 0:
            raise ValueError("size must be positive")
        if not (0 < self.prob <= 1):
            raise ValueError("prob must be on the interval (0, 1]")
        if not math.isfinite(self.size):
            raise ValueError("size must be finite")

    def mean(self) -> float:
        return self.size * (1 - self.prob) / self.prob + 1


@dataclass(frozen=True)
class DirichletRV:
    """R/DirichletRV.R: Dirichlet(alpha); only scalar alpha is accepted by the
    sampler (src/entities.cpp:192-196)."""

    alpha: float

    def __post_init__(self):
        if not (np.ndim(self.alpha) == 0 and self.alpha > 0):
            raise ValueError("alpha must be a positive scalar")


@dataclass(frozen=True)
class DirichletProcess:
    """R/DirichletProcess.R: DP over the distortion distribution with
    concentration ``alpha`` (a positive number, ``inf`` or a GammaRV)."""

    alpha: object = math.inf

    def __post_init__(self):
        if not isinstance(self.alpha, GammaRV) and not self.alpha > 0:
            raise ValueError("alpha must be positive or a GammaRV")

    def mean_alpha(self) -> float:
        return self.alpha.mean() if isinstance(self.alpha, GammaRV) else float(self.alpha)


def _is_rv(x) -> bool:
    return isinstance(x, (BetaRV, GammaRV, ShiftedNegBinomRV, DirichletRV))


# ----------------------------------------------------------------- distances


def levenshtein_one_to_many(y: str, codes: np.ndarray, lens: np.ndarray) -> np.ndarray:
    """Unit-cost edit distance from ``y`` to every row of ``codes``.

    ``codes`` is a [D, L] uint32 matrix of code points padded with 0 and ``lens``
    the true lengths.  Restates what ``comparator::Levenshtein()`` computes when
    called as ``dist_fn(y, domain)`` (R/AttributeIndex.R:119; third-party package,
    unit insertion/deletion/substitution costs, not normalised).
    The DP is vectorised over the D targets.
    """
    ycodes = np.array([ord(c) for c in y], dtype=np.uint32)
    D, L = codes.shape
    prev = np.broadcast_to(np.arange(L + 1, dtype=np.int32), (D, L + 1)).copy()
    cur = np.empty_like(prev)
    for i, yc in enumerate(ycodes, start=1):
        cur[:, 0] = i
        sub = prev[:, :-1] + (codes != yc)
        dele = prev[:, 1:] + 1
        best = np.minimum(sub, dele)
        # insertion needs a running minimum along the row
        for j in range(1, L + 1):
            cur[:, j] = np.minimum(best[:, j - 1], cur[:, j - 1] + 1)
        prev, cur = cur, prev
    return prev[np.arange(D), lens].astype(np.float64)


class Levenshtein:
    """Stand-in for ``comparator::Levenshtein()`` (unit costs)."""

    def __call__(self, y: str, domain_codes: np.ndarray, domain_lens: np.ndarray) -> np.ndarray:
        return levenshtein_one_to_many(y, domain_codes, domain_lens)


class Constant:
    """Stand-in for ``comparator::Constant()`` (R/Attribute.R:144)."""

    def __call__(self, y, domain_codes, domain_lens):
        return np.zeros(len(domain_lens))


def transform_dist_fn(dist_fn: Callable, threshold: float, scaling_factor: float = 1.0):
    """R/transform_dist_fn.R:27-40: scale, then map distances above the threshold
    to infinity."""
    if scaling_factor < 0:
        raise ValueError("scaling_factor must be non-negative")
    if threshold < 0:
        raise ValueError("threshold must be non-negative")
    return TransformedDistFn(dist_fn, threshold, scaling_factor)


@dataclass
class TransformedDistFn:
    """What ``transform_dist_fn`` returns: a callable object rather than a closure, so that a model that
    holds it can be pickled (``bench.py`` hands the model built by rank 0 to the other ranks of the
    node through /dev/shm)."""

    dist_fn: Callable
    threshold: float
    scaling_factor: float = 1.0

    def __call__(self, y, domain_codes, domain_lens):
        d = self.scaling_factor * self.dist_fn(y, domain_codes, domain_lens)
        return np.where(d <= self.threshold, d, np.inf)


# ------------------------------------------------------------ attribute specs


@dataclass
class Attribute:
    """R/Attribute.R:112-119."""

    dist_fn: Callable
    distort_prob_prior: object  # BetaRV or a fixed probability
    distort_dist_prior: DirichletProcess = field(default_factory=DirichletProcess)
    exclude_entity_value: bool = True
    entity_dist_prior: object = None  # None | Mapping[str, float] | DirichletRV
    categorical: bool = False

    def __post_init__(self):
        p = self.distort_prob_prior
        if not isinstance(p, BetaRV) and not (0 <= float(p) <= 1):
            raise ValueError("distort_prob_prior must be a BetaRV or a probability")
        if not self.exclude_entity_value and math.isfinite(self.distort_dist_prior.mean_alpha()):
            # the C++ assumes exclude_entity_value whenever the concentration is finite
            # (comment at src/links.cpp:251)
            raise ValueError("a finite distort_dist_prior requires exclude_entity_value=TRUE")


def CategoricalAttribute(distort_prob_prior, distort_dist_prior=None,
                         exclude_entity_value=True, entity_dist_prior=None) -> Attribute:
    """R/Attribute.R:140-149: an attribute with a constant distance function."""
    return Attribute(Constant(), distort_prob_prior,
                     distort_dist_prior or DirichletProcess(),
                     exclude_entity_value, entity_dist_prior, categorical=True)


# ------------------------------------------------------------------- RP priors


@dataclass
class PitmanYorRP:
    """R/PitmanYorRP.R: alpha (number or GammaRV), d (number in [0,1) or BetaRV)."""

    alpha: object
    d: object

    def __post_init__(self):
        if not isinstance(self.alpha, GammaRV):
            if isinstance(self.d, BetaRV):
                if self.alpha <= 0:
                    raise ValueError("alpha must be > 0")
            elif self.alpha <= -self.d:
                raise ValueError("alpha must be > -d")
            if not math.isfinite(self.alpha):
                raise ValueError("alpha must be finite")
        if not isinstance(self.d, BetaRV) and not (0.0 <= self.d < 1.0):
            raise ValueError("d must be on the interval [0, 1)")


def EwensRP(alpha) -> PitmanYorRP:
    """R/PitmanYorRP.R:107-109."""
    return PitmanYorRP(alpha, 0.0)


@dataclass
class GeneralizedCouponRP:
    """R/GeneralizedCouponRP.R: m (integer or ShiftedNegBinomRV), kappa (number,
    inf or GammaRV)."""

    m: object
    kappa: object

    def __post_init__(self):
        if not isinstance(self.m, ShiftedNegBinomRV):
            if self.m <= 0:
                raise ValueError("m must be positive")
            if int(self.m) != self.m:
                raise ValueError("m must be an integer")
        if not isinstance(self.kappa, GammaRV) and not self.kappa > 0:
            raise ValueError("kappa must be strictly positive")


# -------------------------------------------------------------- attribute index


@dataclass
class AttributeIndex:
    """R/AttributeIndex.R:3-40 slots.  ``close_*`` is the CSR form of the R list
    ``close_values`` (row v = entity value, entries = record values w with
    p(w | v))."""

    domain: list
    probs: np.ndarray
    close_row_ptr: np.ndarray | None
    close_val_ids: np.ndarray | None
    close_probs: np.ndarray | None
    max_exp_factor: np.ndarray | None
    n_missing: int

    @property
    def is_constant(self) -> bool:
        return self.close_row_ptr is None


def _encode_domain(domain: Sequence[str]):
    L = max((len(s) for s in domain), default=0)
    codes = np.zeros((len(domain), max(L, 1)), dtype=np.uint32)
    lens = np.zeros(len(domain), dtype=np.int64)
    for i, s in enumerate(domain):
        lens[i] = len(s)
        codes[i, :len(s)] = [ord(c) for c in s]
    return codes, lens


def compute_close_values(domain: Sequence[str], dist_fn: Callable, include_self: bool):
    """R/AttributeIndex.R:117-137.  For every domain value y: thresholded distances
    to the whole domain, self set to Inf unless ``include_self``; the finite ones get
    ``exp(-d) / sum`` and ``max_exp_factor[y] = exp(-0.5 * min_d / max_d_overall)``
    (0 when y has no finite neighbour)."""
    codes, lens = _encode_domain(domain)
    D = len(domain)
    row_ptr = np.zeros(D + 1, dtype=np.int64)
    cols, probs = [], []
    min_d = np.full(D, np.inf)
    max_d = np.full(D, np.nan)
    for i, y in enumerate(domain):
        d = np.asarray(dist_fn(y, codes, lens), dtype=np.float64).copy()
        if not include_self:
            d[i] = np.inf
        min_d[i] = d.min() if D else np.inf
        fin = np.flatnonzero(np.isfinite(d))
        if fin.size:
            df = d[fin]
            max_d[i] = df.max()
            p = np.exp(-df)
            p = p / p.sum()
            cols.append(fin.astype(np.int32))
            probs.append(p)
        row_ptr[i + 1] = row_ptr[i] + fin.size
    overall = 0.0 if np.all(np.isnan(max_d)) else np.nanmax(max_d)
    with np.errstate(divide="ignore", invalid="ignore"):
        mef = np.where(np.isfinite(min_d), np.exp(-0.5 * min_d / overall), 0.0)
    cols = np.concatenate(cols) if cols else np.zeros(0, np.int32)
    probs = np.concatenate(probs) if probs else np.zeros(0)
    return row_ptr, cols, probs, mef


def build_attribute_index(codes: np.ndarray, domain: Sequence[str], dist_fn,
                          exclude_entity_value: bool, neighbours=None) -> AttributeIndex:
    """R/AttributeIndex.R:68-88.  ``codes`` are 0-based level codes with NA < 0.
    ``neighbours`` lets a caller supply a precomputed (row_ptr, cols, probs, mef)
    tuple (used by the synthetic generator, whose neighbourhoods are computed by
    the native helper)."""
    D = len(domain)
    counts = np.bincount(codes[codes >= 0], minlength=D).astype(np.float64) + 1.0
    probs = counts / counts.sum()
    n_missing = int((codes < 0).sum())
    if dist_fn is None:
        return AttributeIndex(list(domain), probs, None, None, None, None, n_missing)
    if neighbours is None:
        neighbours = compute_close_values(domain, dist_fn, not exclude_entity_value)
    rp, ci, pr, mef = neighbours
    return AttributeIndex(list(domain), probs, np.asarray(rp, np.int64), np.asarray(ci, np.int32),
                          np.asarray(pr, np.float64), np.asarray(mef, np.float64), n_missing)


# ------------------------------------------------------------------- the model


@dataclass
class ExchangERModel:
    """The S4 slots of ``ExchangERModel`` (R/ExchangERModel.R:166-182) with 0-based
    integer codes (the shift of src/sample.cpp:89-93 already applied)."""

    iteration: int
    attr_names: list
    attr_params: list            # list[Attribute], categorical first
    file_ids: np.ndarray         # [N] int32, 0-based
    file_levels: list
    rec_ids: list
    rec_attrs: np.ndarray        # [A, N] int32, NA = INT32_MIN
    rec_distortions: np.ndarray  # [A, N] int32
    ent_attrs: np.ndarray        # [A, K] int32
    ent_ids: np.ndarray          # [K] int32 labels
    links: np.ndarray            # [N] int32 labels
    distort_probs: np.ndarray    # [F, A] float64
    distort_dist_concs: np.ndarray  # [A]
    clust_params: object         # RP with numbers only
    clust_prior: object          # RP possibly holding RVs
    attr_indices: list           # list[AttributeIndex]

    def __post_init__(self):
        # matrices are held column-major, as R holds them: the C ABI reads them in place
        for name in ("rec_attrs", "rec_distortions", "ent_attrs"):
            setattr(self, name, np.asfortranarray(getattr(self, name), dtype=np.int32))
        self.distort_probs = np.asfortranarray(self.distort_probs, dtype=np.float64)

    @property
    def n_records(self) -> int:
        return self.rec_attrs.shape[1]

    @property
    def n_attrs(self) -> int:
        return self.rec_attrs.shape[0]

    @property
    def n_files(self) -> int:
        return len(self.file_levels)


def _first_appearance_domain(col: Sequence) -> list:
    seen, out = set(), []
    for v in col:
        if v is None:
            continue
        s = _as_character(v)
        if s not in seen:
            seen.add(s)
            out.append(s)
    return out


def _as_character(v) -> str:
    if isinstance(v, float) and v == int(v):
        return str(int(v))
    return str(v)


def clust_params_from_prior(clust_prior, n_records: int):
    """R/ExchangERModel.R:424-451: random hyper-parameters start at their prior
    mean; the coupon count m is raised to at least n_records."""
    if isinstance(clust_prior, GeneralizedCouponRP):
        m = clust_prior.m.mean() if _is_rv(clust_prior.m) else clust_prior.m
        if not math.isfinite(m):
            m = n_records
        m = max(m, n_records)
        kappa = clust_prior.kappa.mean() if _is_rv(clust_prior.kappa) else clust_prior.kappa
        out = GeneralizedCouponRP.__new__(GeneralizedCouponRP)
        out.m, out.kappa = float(m), float(kappa)
        return out
    alpha = clust_prior.alpha.mean() if _is_rv(clust_prior.alpha) else clust_prior.alpha
    d = clust_prior.d.mean() if _is_rv(clust_prior.d) else clust_prior.d
    out = PitmanYorRP.__new__(PitmanYorRP)
    out.alpha, out.d = float(alpha), float(d)
    return out


def exchanger(data: Mapping[str, Sequence], attr_params: Mapping[str, Attribute], clust_prior,
              file_id_colname: str | None = None, rec_id_colname: str | None = None,
              *, na_fill_seed: int = 0, neighbours: Mapping[str, tuple] | None = None
              ) -> ExchangERModel:
    """Initialise the model: mirror of ``exchanger()`` (R/ExchangERModel.R:239-351).

    ``data`` maps column names to sequences (``None`` = NA).  Attributes are sorted so
    that categorical ones come first (:252-256); every column is factor-coded with its
    domain in order of first appearance (:468-504); every record starts linked to its
    own entity whose attributes equal the record's, missing ones drawn from the
    empirical pmf (:318-334; the reference uses R's RNG there, here a numpy generator
    seeded with ``na_fill_seed``).
    """
    if not isinstance(attr_params, Mapping) or not attr_params:
        raise ValueError("attr_params must be a named list")
    names = list(attr_params)
    n_records = len(next(iter(data.values())))
    if n_records < 2:
        raise ValueError("data must contain at least two records for ER to be feasible")
    # order(class, decreasing=TRUE) is stable: "CategoricalAttribute" > "Attribute"
    names = [n for n in names if attr_params[n].categorical] + \
            [n for n in names if not attr_params[n].categorical]
    missing = [n for n in names if n not in data]
    if missing:
        raise ValueError("`data` is missing the following attributes: " + " ".join(missing))

    if file_id_colname is not None:
        if file_id_colname not in data:
            raise ValueError("`file_id_colname` does not match any columns in data")
        raw_files = list(data[file_id_colname])
        if any(f is None for f in raw_files):
            raise ValueError("file identifiers cannot contain missing values")
    else:
        raw_files = [0] * n_records
    file_levels = sorted(set(raw_files), key=lambda v: (str(type(v)), v))
    fmap = {v: i for i, v in enumerate(file_levels)}
    file_ids = np.array([fmap[v] for v in raw_files], dtype=np.int32)

    if rec_id_colname is not None:
        if rec_id_colname not in data:
            raise ValueError("rec_id_colname does not match any columns in data")
        rec_ids = list(data[rec_id_colname])
        if len(set(rec_ids)) != n_records:
            raise ValueError("record identifiers must be unique")
    else:
        rec_ids = list(range(1, n_records + 1))

    if isinstance(clust_prior, GeneralizedCouponRP):
        if not _is_rv(clust_prior.m) and clust_prior.m < n_records:
            raise ValueError("Coupon collecting parameter m < n_records is currently not supported.")
        if (not _is_rv(clust_prior.kappa) and math.isinf(clust_prior.kappa)
                and _is_rv(clust_prior.m)):
            raise ValueError("Random parameter m is currently not supported when kappa = Inf")
    elif not isinstance(clust_prior, PitmanYorRP):
        raise ValueError("clust_prior must be a RP")

    A = len(names)
    rec_attrs = np.full((A, n_records), NA_INT, dtype=np.int32)
    indices = []
    for a, name in enumerate(names):
        spec = attr_params[name]
        col = data[name]
        prior = spec.entity_dist_prior
        emp_domain = _first_appearance_domain(col)
        if isinstance(prior, Mapping):
            domain = list(prior)
            if set(emp_domain) - set(domain):
                raise ValueError(f"attribute '{name}' in data contains values not in entity_dist_prior")
        else:
            domain = emp_domain
        if len(emp_domain) <= 1:
            raise ValueError(f"attribute '{name}' has a domain of size 1 and cannot be used "
                             "for entity resolution")
        lut = {s: i for i, s in enumerate(domain)}
        rec_attrs[a] = [NA_INT if v is None else lut[_as_character(v)] for v in col]
        nb = None if neighbours is None else neighbours.get(name)
        indices.append(build_attribute_index(
            rec_attrs[a], domain, None if spec.categorical else spec.dist_fn,
            spec.exclude_entity_value, nb))

    specs = [attr_params[n] for n in names]
    theta0 = np.array([s.distort_prob_prior.mean() if isinstance(s.distort_prob_prior, BetaRV)
                       else float(s.distort_prob_prior) for s in specs])
    distort_probs = np.tile(theta0, (len(file_levels), 1))
    concs = np.array([s.distort_dist_prior.mean_alpha() for s in specs], dtype=np.float64)

    ent_attrs = rec_attrs.copy()
    rng = np.random.default_rng(na_fill_seed)
    for a, idx in enumerate(indices):
        if idx.n_missing > 0:
            miss = ent_attrs[a] < 0
            ent_attrs[a, miss] = rng.choice(len(idx.probs), size=int(miss.sum()), p=idx.probs)

    ids = np.arange(n_records, dtype=np.int32)
    return ExchangERModel(
        iteration=0, attr_names=names, attr_params=specs, file_ids=file_ids,
        file_levels=list(file_levels), rec_ids=rec_ids, rec_attrs=rec_attrs,
        rec_distortions=np.zeros((A, n_records), dtype=np.int32), ent_attrs=ent_attrs,
        ent_ids=ids.copy(), links=ids.copy(), distort_probs=distort_probs,
        distort_dist_concs=concs, clust_params=clust_params_from_prior(clust_prior, n_records),
        clust_prior=clust_prior, attr_indices=indices)


def model_with_state(model: ExchangERModel, state: dict) -> ExchangERModel:
    """A copy of ``model`` whose mutable slots are replaced by a sampler state - what
    ``buildFinalS4State`` assembles (src/sample.cpp:14-63).  ``state`` holds ``iteration``,
    ``links`` [N], ``ent_ids`` [K], ``ent_attrs`` [A, K], ``rec_distortions`` [A, N],
    ``distort_probs`` [F, A] and ``clust`` (dict with alpha/d or m/kappa)."""
    import copy

    out = copy.copy(model)
    out.iteration = int(state["iteration"])
    out.links = np.asarray(state["links"], np.int32).copy()
    out.ent_ids = np.asarray(state["ent_ids"], np.int32).copy()
    out.ent_attrs = np.array(state["ent_attrs"], np.int32, order="F")
    out.rec_distortions = np.array(state["rec_distortions"], np.int32, order="F")
    out.distort_probs = np.array(state["distort_probs"], np.float64, order="F")
    if "distort_dist_concs" in state:
        out.distort_dist_concs = np.asarray(state["distort_dist_concs"], np.float64).copy()
    cp = copy.copy(model.clust_params)
    cl = state["clust"]
    if isinstance(cp, PitmanYorRP):
        cp.alpha, cp.d = float(cl["alpha"]), float(cl["d"])
    else:
        cp.m, cp.kappa = float(cl["m"]), float(cl["kappa"])
    out.clust_params = cp
    return out


def exchanger_coded(codes: np.ndarray, domains: Sequence[Sequence[str]], attr_names: Sequence[str],
                    attr_params: Sequence[Attribute], clust_prior, file_ids: np.ndarray | None = None,
                    neighbours: Mapping[str, tuple] | None = None, na_fill_seed: int = 0
                    ) -> ExchangERModel:
    """``exchanger()`` for data that is already integer coded (at 10 M rows
    R's ``factor`` / ``data.matrix`` path, R/ExchangERModel.R:306-311, is the slow part of model
    construction).  ``codes`` is ``[A, N]`` with 0-based level codes and negative = NA; attributes
    must already be ordered categorical first.  Everything else follows ``exchanger()``:
    singleton initial links, prior-mean hyper-parameters, NA entity values drawn from the
    empirical pmf."""
    codes = np.ascontiguousarray(codes, dtype=np.int32)
    A, N = codes.shape
    if N < 2:
        raise ValueError("data must contain at least two records for ER to be feasible")
    if any(not s.categorical for s in attr_params[:sum(s.categorical for s in attr_params)]):
        raise ValueError("categorical attributes must come first")
    if isinstance(clust_prior, GeneralizedCouponRP) and not _is_rv(clust_prior.m) and clust_prior.m < N:
        raise ValueError("Coupon collecting parameter m < n_records is currently not supported.")
    rec_attrs = np.where(codes < 0, NA_INT, codes).astype(np.int32)
    indices = []
    for a, (name, spec) in enumerate(zip(attr_names, attr_params)):
        nb = None if neighbours is None else neighbours.get(name)
        indices.append(build_attribute_index(rec_attrs[a], domains[a], None if spec.categorical else spec.dist_fn,
                                             spec.exclude_entity_value, nb))
    theta0 = np.array([s.distort_prob_prior.mean() if isinstance(s.distort_prob_prior, BetaRV)
                       else float(s.distort_prob_prior) for s in attr_params])
    if file_ids is None:
        file_ids = np.zeros(N, np.int32)
    n_files = int(file_ids.max()) + 1
    ent_attrs = rec_attrs.copy()
    rng = np.random.default_rng(na_fill_seed)
    for a, idx in enumerate(indices):
        if idx.n_missing > 0:
            miss = ent_attrs[a] < 0
            ent_attrs[a, miss] = rng.choice(len(idx.probs), size=int(miss.sum()), p=idx.probs)
    ids = np.arange(N, dtype=np.int32)
    return ExchangERModel(
        iteration=0, attr_names=list(attr_names), attr_params=list(attr_params),
        file_ids=np.asarray(file_ids, np.int32), file_levels=list(range(n_files)), rec_ids=None,
        rec_attrs=rec_attrs, rec_distortions=np.zeros((A, N), np.int32), ent_attrs=ent_attrs,
        ent_ids=ids.copy(), links=ids.copy(), distort_probs=np.tile(theta0, (n_files, 1)),
        distort_dist_concs=np.array([s.distort_dist_prior.mean_alpha() for s in attr_params]),
        clust_params=clust_params_from_prior(clust_prior, N), clust_prior=clust_prior, attr_indices=indices)

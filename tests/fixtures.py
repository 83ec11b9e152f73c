"""Model builders shared by the tests, the smoke test and the benchmark (no reference access)."""
from __future__ import annotations

import os

import numpy as np

from exchanger_b200 import (Attribute, BetaRV, CategoricalAttribute, EwensRP, GammaRV,
                            GeneralizedCouponRP, Levenshtein, PitmanYorRP, exchanger,
                            transform_dist_fn)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_rldata(name="rldata500", n=None):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    cols = {}
    for k in z.files:
        if k.endswith("__na") or k == "identity":
            continue
        na = z[k + "__na"]
        cols[k] = [None if m else v for v, m in zip(z[k].tolist(), na.tolist())]
    ident = z["identity"]
    if n is not None:
        cols = {k: v[:n] for k, v in cols.items()}
        ident = ident[:n]
    return cols, ident


def readme_attr_params(prior=None, with_c2=False, exclude=True):
    """The attribute specification of README.md:76-89 (Levenshtein thr=3 names, categorical dates)."""
    prior = prior or BetaRV(1, 1)
    lev = transform_dist_fn(Levenshtein(), threshold=3)
    ap = dict(
        fname_c1=Attribute(lev, prior, exclude_entity_value=exclude),
        lname_c1=Attribute(lev, prior, exclude_entity_value=exclude),
        by=CategoricalAttribute(prior, exclude_entity_value=exclude),
        bm=CategoricalAttribute(prior, exclude_entity_value=exclude),
        bd=CategoricalAttribute(prior, exclude_entity_value=exclude),
    )
    if with_c2:  # the mostly-missing second name components exercise the NA paths
        ap["fname_c2"] = Attribute(lev, prior, exclude_entity_value=exclude)
        ap["lname_c2"] = CategoricalAttribute(prior, exclude_entity_value=exclude)
    return ap


def readme_model(name="rldata500", n=None, clust_prior=None, **kw):
    cols, ident = load_rldata(name, n)
    clust_prior = clust_prior or PitmanYorRP(GammaRV(1, 1), BetaRV(1, 1))
    return exchanger(cols, readme_attr_params(**kw), clust_prior), ident


def model_with_files(name="rldata500", n=None, n_files=3, **kw):
    cols, ident = load_rldata(name, n)
    rng = np.random.default_rng(5)
    cols["file"] = [f"f{int(i)}" for i in rng.integers(0, n_files, len(ident))]
    return exchanger(cols, readme_attr_params(**kw), PitmanYorRP(GammaRV(1, 1), BetaRV(1, 1)),
                     file_id_colname="file"), ident


PRIORS = {
    "pitman_yor": lambda n: PitmanYorRP(GammaRV(1, 1), BetaRV(1, 1)),
    "pitman_yor_fixed": lambda n: PitmanYorRP(2.0, 0.5),
    "ewens": lambda n: EwensRP(GammaRV(1, 1)),
    "gen_coupon": lambda n: GeneralizedCouponRP(n, GammaRV(1, 1)),
    "coupon": lambda n: GeneralizedCouponRP(2 * n, float("inf")),
}

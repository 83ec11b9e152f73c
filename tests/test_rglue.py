"""The R-side binding (exchanger_b200/rglue/sample_b200.cpp) against the reference's own
src/sample.cpp.  R is not in the image: both are compiled against the stand-in Rcpp headers of
oracle/ref_shim/ (oracle/Makefile, target `glue`), fed the same `ExchangERModel` S4 object and
their `ExchangERFit` results compared slot for slot - names, order, types, dimensions and dimnames
must be identical (values differ: the reference draws from its own RNG stream).  The glue is
linked with a test double of libexb200 here (tests/rglue/mock_exb200.cpp); with the real library
in test_gpu_parity.py::test_r_glue_end_to_end_on_the_gpu."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from exchanger_b200 import BetaRV, DirichletRV, GammaRV, GeneralizedCouponRP, PitmanYorRP, ShiftedNegBinomRV, exchanger
from exchanger_b200.capi import CAttr, CModel, FlatModel
from fixtures import load_rldata, model_with_files, readme_attr_params, readme_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOCK_SO = os.path.join(ROOT, "oracle", "_build", "librglue_mock.so")

pytestmark = pytest.mark.skipif(not os.path.exists(MOCK_SO), reason="glue harness not built (no reference sources here)")


def harness(path=MOCK_SO):
    lib = C.CDLL(path)
    lib.rglue_run.restype = C.c_char_p
    lib.rglue_run.argtypes = [C.POINTER(CModel), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int]
    lib.rglue_seed_for.restype = C.c_uint64
    lib.rglue_seed_for.argtypes = [C.c_uint64]
    return lib


def run(lib, model, n_samples, thin, burnin, which, abort_after=-1, seed=5, max_values=200000):
    flat = FlatModel(model, seed)
    out = json.loads(lib.rglue_run(C.byref(flat.c), n_samples, thin, burnin, which, abort_after, seed, max_values).decode())
    assert "error" not in out, out
    return out


def shape(node):
    """Everything but the values: type, class, length, dimensions, names, dimnames, slot / attribute names."""
    if node is None:
        return None
    out = {k: node[k] for k in ("type", "class", "len", "dim", "names") if k in node}
    for key in ("slots", "attrs"):
        if key in node:
            out[key] = {k: (v if k in ("rownames", "colnames", "names", "levels") else shape(v)) for k, v in node[key].items()}
    if "items" in node:
        out["items"] = [shape(x) for x in node["items"]]
    return out


MODELS = {
    "readme_pitman_yor": lambda: readme_model("rldata500", n=60)[0],
    "three_files_gen_coupon_random_m": lambda: _with_prior(model_with_files("rldata500", n=60)[0],
                                                           GeneralizedCouponRP(ShiftedNegBinomRV(100, 0.5), GammaRV(1, 1))),
    "coupon_fixed": lambda: readme_model("rldata500", n=60, clust_prior=GeneralizedCouponRP(120, float("inf")))[0],
    "fixed_hyperparameters": lambda: readme_model("rldata500", n=60, clust_prior=PitmanYorRP(2.0, 0.5))[0],
    "dirichlet_and_gamma_concentration": lambda: _vignette(),
}


def _with_prior(model, prior):
    from exchanger_b200.model import clust_params_from_prior
    model.clust_prior = prior
    model.clust_params = clust_params_from_prior(prior, model.n_records)
    return model


def _vignette():
    from exchanger_b200 import DirichletProcess
    cols, _ = load_rldata("rldata500", 60)
    ap = readme_attr_params(prior=BetaRV(1, 4))
    for k in ap:
        ap[k].distort_dist_prior = DirichletProcess(GammaRV(2, 1e-2))
    ap["bm"].entity_dist_prior = DirichletRV(1.0)
    return exchanger(cols, ap, PitmanYorRP(GammaRV(1, 1), BetaRV(1, 1)))


@pytest.mark.parametrize("name", list(MODELS))
def test_glue_and_reference_return_the_same_fit_structure(name):
    """src/sample.cpp:14-63,133-166,205-218: history elements (names, order, matrix shapes, column names
    "attr[file]" in f + a F order), the final ExchangERModel (slots, dimnames) - identical for both."""
    lib, model = harness(), MODELS[name]()
    ref = run(lib, model, 4, 2, 3, which=0)
    glue = run(lib, model, 4, 2, 3, which=1)
    A = model.n_attrs
    for out in (ref, glue):   # the number of entities is a sampled value (the test double keeps K = N): compare it symbolically
        st = out["fit"]["slots"]["state"]["slots"]
        K = st["ent_ids"]["len"]
        assert st["ent_attrs"]["dim"] == [A, K] and st["ent_attrs"]["len"] == A * K and 1 <= K <= model.n_records
        st["ent_ids"]["len"], st["ent_attrs"]["len"], st["ent_attrs"]["dim"] = "K", "A*K", [A, "K"]
    assert shape(glue["fit"]) == shape(ref["fit"])
    hist = {n: x for n, x in zip(glue["fit"]["slots"]["history"]["names"], glue["fit"]["slots"]["history"]["items"])}
    assert list(hist) == ["links", "distort_probs", "distort_counts", "distort_dist_concs", "n_linked_ents"] + \
        (["clust_params"] if name not in ("coupon_fixed", "fixed_hyperparameters") else [])
    N, A, F = model.n_records, model.n_attrs, model.n_files
    assert hist["links"]["dim"] == [4, N] and hist["distort_probs"]["dim"] == [4, F * A]
    assert hist["distort_probs"]["attrs"]["colnames"]["values"][:F + 1] == [f"attr1[file{f + 1}]" for f in range(F)] + ["attr2[file1]"]
    # both count the same sweeps: the progress bar advanced once per sweep, 3 + 2 * 3 of them
    assert glue["increments"] == ref["increments"] == 9
    assert glue["fit"]["slots"]["state"]["slots"]["iteration"]["values"] == ref["fit"]["slots"]["state"]["slots"]["iteration"]["values"] == [9]


def test_glue_marshals_values_both_ways():
    """What crosses the ABI: 0-based ids and NA on the way in (src/sample.cpp:89-93), the save schedule,
    R-style 1-based labels, random cluster parameters only and column-major matrices on the way out."""
    lib = harness()
    cols, _ = load_rldata("rldata500", 50)
    cols["fname_c1"] = [None if i % 9 == 0 else v for i, v in enumerate(cols["fname_c1"])]
    model = exchanger(cols, readme_attr_params(), PitmanYorRP(GammaRV(1, 1), 0.3))
    glue = run(lib, model, 3, 2, 0, which=1)
    lib.mock_seen_model.restype = C.POINTER(CModel)
    lib.mock_seen_rec_attrs.restype = C.POINTER(C.c_int)
    lib.mock_seen_attrs.restype = C.POINTER(CAttr)
    seen = lib.mock_seen_model().contents
    N, A = model.n_records, model.n_attrs
    assert (seen.n_records, seen.n_attrs, seen.n_files, seen.n_entities, seen.iteration) == (N, A, 1, N, 0)
    got = np.ctypeslib.as_array(lib.mock_seen_rec_attrs(), shape=(N, A))
    want = model.rec_attrs.T
    assert np.array_equal(got >= 0, want >= 0) and np.array_equal(got[want >= 0], want[want >= 0])   # NA stays negative
    assert seen.clust.kind == 0 and seen.clust.alpha_prior_shape == 1.0 and seen.clust.d_prior_shape1 == 0.0 and seen.clust.d == 0.3
    attrs = lib.mock_seen_attrs()
    for a, idx in enumerate(model.attr_indices):
        assert attrs[a].domain_size == len(idx.probs) and attrs[a].is_constant == int(idx.is_constant)
        assert attrs[a].distort_prob_shape1 == 1.0 and np.isinf(attrs[a].distort_dist_conc)
    assert seen.seed == lib.rglue_seed_for(5)          # drawn from R's RNG: set.seed() fixes the chain
    fit = glue["fit"]["slots"]
    hist = dict(zip(fit["history"]["names"], fit["history"]["items"]))
    links = np.array(hist["links"]["values"]).reshape(N, 3).T          # column-major n_samples x N
    assert np.array_equal(links, np.tile(np.arange(N) % 7 + 1, (3, 1)))
    assert hist["n_linked_ents"]["values"] == [1002, 1004, 1006]        # thin 2, no burn-in: sweeps 2, 4, 6
    assert hist["clust_params"]["dim"] == [3, 1] and hist["clust_params"]["attrs"]["colnames"]["values"] == ["alpha"]
    assert hist["clust_params"]["values"] == [2.5, 4.5, 6.5]
    theta = np.array(hist["distort_probs"]["values"]).reshape(A, 3).T
    assert np.allclose(theta[1], 4 + 0.001 * np.arange(A))
    state = fit["state"]["slots"]
    # labels come back from the library 0-based ("smallest member record") and leave the glue 1-based;
    # the test double echoes the labels it was given (1..N), hence 2..N+1 here
    assert state["ent_attrs"]["dim"] == [A, N] and state["links"]["values"] == list(range(2, N + 2))
    assert state["rec_distortions"]["attrs"]["rownames"] == state["ent_attrs"]["attrs"]["rownames"]
    assert state["clust_params"]["class"] == "PitmanYorRP" and state["clust_params"]["slots"]["d"]["values"] == [0.3]


def test_interrupt_returns_the_partial_history_like_the_reference():
    """Progress::check_abort() (src/sample.cpp:200): both stop after the same sweep and return the
    history allocated for all samples with the rows saved so far."""
    lib, model = harness(), readme_model("rldata500", n=40)[0]
    ref = run(lib, model, 5, 2, 0, which=0, abort_after=4)     # the fifth poll answers "interrupted"
    glue = run(lib, model, 5, 2, 0, which=1, abort_after=4)
    assert shape(glue["fit"]) == shape(ref["fit"])
    for out in (ref, glue):
        assert out["increments"] == 4 and out["fit"]["slots"]["state"]["slots"]["iteration"]["values"] == [5]
        hist = dict(zip(out["fit"]["slots"]["history"]["names"], out["fit"]["slots"]["history"]["items"]))
        n_linked = hist["n_linked_ents"]["values"]
        assert hist["links"]["dim"] == [5, 40] and n_linked[2:] == [0, 0, 0] and all(n_linked[:2])


def test_glue_reports_library_errors_as_r_errors():
    lib, model = harness(), readme_model("rldata500", n=40)[0]
    model.clust_prior = "nonsense"
    flat = FlatModel.__new__(FlatModel)
    # an S4 object of an unknown prior class never reaches the library: Rcpp::stop like clust_params.cpp:279
    m = readme_model("rldata500", n=40)[0]
    f = FlatModel(m, 1)
    f.c.clust.kind = 1          # GeneralizedCouponRP ...
    f.c.clust.m = 10.0          # ... with m < n_records: the mock accepts it, so only check the call path works
    out = json.loads(lib.rglue_run(C.byref(f.c), 1, 1, 0, 1, -1, 1, 10).decode())
    assert "fit" in out
    f.c.n_records = 1           # the library refuses: its message becomes the R error
    out = json.loads(lib.rglue_run(C.byref(f.c), 1, 1, 0, 1, -1, 1, 10).decode())
    assert out == {"error": "mock: need two records"}

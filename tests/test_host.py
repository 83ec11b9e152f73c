"""CPU tests of the host side: model preparation (mirror of R/ExchangERModel.R and
R/AttributeIndex.R), argument validation of run_inference, the point estimates, and the C-ABI
library (loads, exports every symbol of include/exb200.h, refuses to run without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from exchanger_b200 import (Attribute, BetaRV, CategoricalAttribute, GammaRV, GeneralizedCouponRP,
                            Levenshtein, PitmanYorRP, exchanger, transform_dist_fn)
from exchanger_b200 import capi
from exchanger_b200.analysis import (clusters_to_membership, mp_clusters, pairwise_measures,
                                     smp_clusters)
from exchanger_b200.model import levenshtein_one_to_many, _encode_domain
from exchanger_b200.sampler import run_inference
from fixtures import load_rldata, readme_attr_params, readme_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_levenshtein_known_distances():
    dom = ["kitten", "sitting", "", "kitten", "mitten", "abc"]
    codes, lens = _encode_domain(dom)
    d = levenshtein_one_to_many("kitten", codes, lens)
    assert d.tolist() == [0, 3, 6, 0, 1, 6]


@pytest.mark.parametrize("name,doms,nnz", [
    ("rldata500", [86, 12, 31, 146, 108], [672, 584]),
    ("rldata10000", [107, 20, 45, 539, 354], [5890, 6170]),
])
def test_model_matches_reference_dataset_statistics(name, doms, nnz):
    """Domain sizes and thresholded-Levenshtein neighbourhood sizes measured on the reference's
    datasets."""
    m, ident = readme_model(name)
    assert m.attr_names == ["by", "bm", "bd", "fname_c1", "lname_c1"]  # categorical first (:252-256)
    assert [len(i.probs) for i in m.attr_indices] == doms
    assert [len(i.close_val_ids) for i in m.attr_indices if not i.is_constant] == nnz
    for i in m.attr_indices:
        assert abs(i.probs.sum() - 1) < 1e-12
        if not i.is_constant:
            rp = i.close_row_ptr
            for v in range(0, len(i.probs), 17):
                if rp[v + 1] > rp[v]:
                    assert abs(i.close_probs[rp[v]:rp[v + 1]].sum() - 1) < 1e-12
            assert np.all((i.max_exp_factor >= 0) & (i.max_exp_factor <= 1))
    assert np.array_equal(m.links, np.arange(m.n_records)) and np.all(m.rec_distortions == 0)
    assert np.allclose(m.distort_probs, 0.5)  # Beta(1,1) prior mean


def test_exchanger_validation_errors():
    cols, _ = load_rldata("rldata500", 50)
    ap = readme_attr_params()
    with pytest.raises(ValueError, match="missing the following attributes"):
        exchanger({k: v for k, v in cols.items() if k != "by"}, ap, PitmanYorRP(1.0, 0.5))
    with pytest.raises(ValueError, match="m < n_records"):
        exchanger(cols, ap, GeneralizedCouponRP(10, 1.0))
    with pytest.raises(ValueError, match="clust_prior must be a RP"):
        exchanger(cols, ap, "nope")
    with pytest.raises(ValueError, match="d must be on the interval"):
        PitmanYorRP(1.0, 1.5)
    one = dict(cols)
    one["bm"] = ["1"] * 50
    with pytest.raises(ValueError, match="domain of size 1"):
        exchanger(one, ap, PitmanYorRP(1.0, 0.5))


def test_missing_values_are_filled_from_the_empirical_distribution():
    cols, _ = load_rldata("rldata500", 200)
    m = exchanger(cols, readme_attr_params(with_c2=True), PitmanYorRP(1.0, 0.5))
    a = m.attr_names.index("fname_c2")
    assert (m.rec_attrs[a] < 0).sum() == m.attr_indices[a].n_missing > 0
    assert np.all(m.ent_attrs >= 0)


def test_run_inference_argument_validation():
    m, _ = readme_model("rldata500", n=50)
    for kw, msg in ((dict(n_samples=-1), "n_samples must be positive"),
                    (dict(n_samples=1, thin_interval=0), "thin_interval must be positive"),
                    (dict(n_samples=1, burnin_interval=-2), "burnin_interval must be non-negative"),
                    (dict(n_samples=1.5), "n_samples must be an integer")):
        with pytest.raises(ValueError, match=msg):
            run_inference(m, **kw)


def test_point_estimates():
    hist = np.array([[0, 0, 2, 3], [0, 0, 2, 2], [5, 5, 7, 7], [1, 1, 1, 3]])
    cl, pr = mp_clusters(hist)
    assert cl[0] == (0, 1) and pr[0] == 0.75
    assert cl[3] == (2, 3) and pr[3] == 0.5
    assert sorted(map(sorted, smp_clusters(hist))) == [[0, 1], [2, 3]]
    pm = pairwise_measures(np.array([0, 0, 1, 2]), clusters_to_membership(smp_clusters(hist), 4))
    assert pm["tp"] == 1 and pm["pred_pairs"] == 2 and pm["recall"] == 1.0 and pm["precision"] == 0.5


def test_library_exports_every_declared_symbol():
    lib = capi.load_library()
    header = open(os.path.join(ROOT, "include", "exb200.h")).read()
    declared = set(re.findall(r"\b(exb200_[a-z_]+)\s*\(", header))
    assert declared == set(capi.EXPORTED_SYMBOLS)
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert lib.exb200_abi_version() == 4


def test_struct_layouts_match_the_header():
    """ctypes mirrors vs the C compiler's view of include/exb200.h."""
    import subprocess, tempfile
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "exb200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(exb200_attr), sizeof(exb200_clust), sizeof(exb200_model),
         sizeof(exb200_history), sizeof(exb200_state), sizeof(exb200_stats));
  printf("%zu %zu %zu\n", offsetof(exb200_model, clust), offsetof(exb200_model, seed), offsetof(exb200_state, clust));
  return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")], check=True)
        out = subprocess.run([os.path.join(d, "t")], capture_output=True, text=True, check=True).stdout.split()
    sizes = [C.sizeof(x) for x in (capi.CAttr, capi.CClust, capi.CModel, capi.CHistory, capi.CState, capi.CStats)]
    assert [int(x) for x in out[:6]] == sizes
    assert [int(x) for x in out[6:]] == [capi.CModel.clust.offset, capi.CModel.seed.offset, capi.CState.clust.offset]


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run (it must never route through the oracle)."""
    lib = capi.load_library()
    if lib.exb200_device_count() > 0:
        pytest.skip("a GPU is present")
    m, _ = readme_model("rldata500", n=50)
    with pytest.raises(capi.Exb200Error) as e:
        run_inference(m, n_samples=1)
    assert e.value.status == -3
    from exchanger_b200.analysis import DevicePointEstimate
    with pytest.raises(capi.Exb200Error) as e:   # the device point estimates have no host path either
        DevicePointEstimate(50)
    assert e.value.status == -3
    src = ""
    for root, _, files in os.walk(os.path.join(ROOT, "exchanger_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                src += open(os.path.join(root, f)).read()
    assert "exoracle" not in src and "libexref" not in src and "import checkers" not in src

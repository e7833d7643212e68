import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    # belt and braces: a gpu test collected on a GPU-less machine without `-m "not gpu"` is skipped, not failed
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


import ace_jl_b200 as ace  # noqa: E402
from ace_jl_b200.utils import RnYlm_1pbasis, philox, rand_envs  # noqa: E402


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def make_basis(kind: str):
    """The model zoo shared by the CPU and GPU tests (small versions of BASELINE.json's configs)."""
    if kind == "inv_simple_3_6":         # the reference's unit-test basis (test_symmbasis.jl, test_linearmodel.jl)
        return ace.SymmetricBasis(ace.Invariant(), RnYlm_1pbasis(maxdeg=6), ace.SimpleSparseBasis(3, 6))
    if kind == "inv_sparse_3_10":        # config 1 (bm_basis.jl:58-62 with maxorder = ord)
        Bsel = ace.SparseBasis(maxorder=3, p=1, default_maxdeg=10, weight={"n": 1.0, "l": 1.5})
        return ace.SymmetricBasis(ace.Invariant(), RnYlm_1pbasis(maxdeg=10, maxL=math.ceil(10 / 1.5), Bsel=Bsel), Bsel)
    if kind == "inv_sparse_3_12":        # config 2 (bm_linear.jl:151-155)
        Bsel = ace.SparseBasis(maxorder=3, p=1, default_maxdeg=12, weight={"n": 1.0, "l": 1.5})
        return ace.SymmetricBasis(ace.Invariant(), RnYlm_1pbasis(maxdeg=12, maxL=math.ceil(12 / 1.5), Bsel=Bsel), Bsel)
    if kind == "inv_sparse_4_8":         # small cousin of config 3 (profile_linearmodel.jl:13-25)
        Bsel = ace.SparseBasis(maxorder=4, p=1, default_maxdeg=8, weight={"n": 1.0, "l": 1.5})
        return ace.SymmetricBasis(ace.Invariant(), RnYlm_1pbasis(maxdeg=8, maxL=math.ceil(8 / 1.5), Bsel=Bsel), Bsel)
    if kind == "euclvec_3_5":            # config 4 (test_euclvec.jl)
        return ace.SymmetricBasis(ace.EuclideanVector(), RnYlm_1pbasis(maxdeg=5), ace.SimpleSparseBasis(3, 5))
    if kind == "euclmat_2_5":            # config 4 (test_EuclideanMatrix.jl: ord 2, deg 5)
        return ace.SymmetricBasis(ace.EuclideanMatrix(), RnYlm_1pbasis(maxdeg=5), ace.SimpleSparseBasis(2, 5))
    if kind == "species_3_5":            # small cousin of config 5 (test_discrete.jl:67-87)
        Bsel = ace.SparseBasis(maxorder=3, p=1, default_maxdeg=5, weight={"n": 1.0, "l": 1.5})
        RnYlm = RnYlm_1pbasis(maxdeg=5, maxL=4)
        B1p = ace.Product1pBasis((ace.Categorical1pBasis(["a", "b", "c", "d"], varsym="mu", idxsym="q"),) + RnYlm.bases)
        return ace.SymmetricBasis(ace.Invariant(), B1p, Bsel)
    if kind == "inv_complexB_2_5":       # SymmetricBasis(...; isreal = false): B, dB stay complex (symmbasis.jl:84-86)
        return ace.SymmetricBasis(ace.Invariant(), RnYlm_1pbasis(maxdeg=5), ace.SimpleSparseBasis(2, 5), isreal=False)
    if kind == "inv_morse_2_6":
        return ace.SymmetricBasis(ace.Invariant(), RnYlm_1pbasis(maxdeg=6, trans=ace.morsetransform(1.3, 1.1), pin=2),
                                  ace.SimpleSparseBasis(2, 6))
    if kind == "inv_agnesi_2_6":
        return ace.SymmetricBasis(ace.Invariant(), RnYlm_1pbasis(maxdeg=6, trans=ace.agnesitransform(1.0, 3)),
                                  ace.SimpleSparseBasis(2, 6))
    if kind == "inv_highL_2_12":         # l up to 10: beyond the statically unrolled harmonics walk (ace_math.cuh kStaticL)
        Bsel = ace.SparseBasis(maxorder=2, p=1, default_maxdeg=12, weight={"n": 1.0, "l": 0.6})
        return ace.SymmetricBasis(ace.Invariant(), RnYlm_1pbasis(maxdeg=12, maxL=10, Bsel=Bsel), Bsel)
    raise KeyError(kind)


_CACHE = {}


@pytest.fixture
def zoo():
    def get(kind):
        if kind not in _CACHE:
            _CACHE[kind] = make_basis(kind)
        return _CACHE[kind]
    return get


def rn_of(basis):
    return basis.pibasis.basis1p.component(0)


def nspecies_of(basis):
    cat = basis.pibasis.basis1p.component(2)
    return 0 if cat is None else len(cat)

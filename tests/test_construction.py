"""Host-side table construction: integer tables and coupling coefficients.

  - table sizes of the five BASELINE configs
  - the tiny worked example of Appendix B (1p ordering, AA rows)
  - Clebsch-Gordan vs sympy on random samples               (test/test_cg.jl:21-44)
  - the Y.Y product-expansion identity                        (test/test_cg.jl:53-77)
  - derived Euclidean seeds vs the reference's tables         (src/eucl/*.jl; only where /root/reference exists)
"""
import math
import os
import re

import numpy as np
import pytest

import ace_jl_b200 as ace
from ace_jl_b200.properties import crmatrix, mrmatrix
from ace_jl_b200.rotations3d import clebschgordan
from ace_jl_b200.utils import RnYlm_1pbasis, philox
from conftest import make_basis


def _sizes(basis):
    o = basis.pibasis.spec.orders
    return (len(basis.pibasis.basis1p), len(basis.pibasis), [int((o == k).sum()) for k in range(o.max() + 1)],
            len(basis), basis.A2Bmap.nnz)


def test_sizes_simple_3_6():
    assert _sizes(make_basis("inv_simple_3_6")) == (43, 247, [1, 6, 59, 181], 97, 247)


def test_sizes_config1():
    assert _sizes(make_basis("inv_sparse_3_10")) == (73, 762, [1, 10, 141, 610], 266, 762)


def test_sizes_config2():
    b = make_basis("inv_sparse_3_12")
    assert _sizes(b) == (114, 1827, [1, 12, 246, 1568], 484, 1827)
    assert max(x[1] for x in b.pibasis.basis1p.spec) == 4          # only l <= 4 survives cleaning


def test_sizes_config3():
    Bsel = ace.SparseBasis(maxorder=4, p=1, default_maxdeg=14, weight={"n": 1.0, "l": 1.5})
    b = ace.SymmetricBasis(ace.Invariant(), RnYlm_1pbasis(maxdeg=14, maxL=math.ceil(14 / 1.5), Bsel=Bsel), Bsel)
    nA, nAA, byord, nB, nnz = _sizes(b)
    assert (nA, byord[:4], nB) == (164, [1, 14, 401, 3689], 2566)
    assert abs(nAA - 15945) <= 20 and abs(nnz - 19512) <= 800      # order-4 count is round-off sensitive (Appendix C.2)


def test_worked_example_orderings():
    b1p = RnYlm_1pbasis(maxdeg=3)
    Bsel = ace.SimpleSparseBasis(2, 3)
    spec = ace.pibasis.build_pibasis_spec(b1p, ace.O3(), Bsel, property=ace.Invariant())
    assert b1p.spec[:9] == [(1, 0, 0), (1, 1, -1), (2, 0, 0), (1, 1, 0), (1, 1, 1), (1, 2, -2), (2, 1, -1), (1, 2, -1), (3, 0, 0)]
    rows = [tuple(int(v) for v in r) for r in spec.iAA2iA]
    assert rows[0] == (0, 0) and rows[1] == (1, 0) and rows[29] == (29, 0)
    assert rows[30:35] == [(1, 1), (3, 1), (9, 1), (11, 1), (22, 1)]
    assert rows[35:38] == [(5, 2), (12, 2), (3, 3)]


def test_index_tables_are_canonical():
    b = make_basis("inv_simple_3_6")
    spec = b.pibasis.spec
    for i in range(len(spec)):
        row = spec.iAA2iA[i]
        o = spec.orders[i]
        assert np.all(row[:o] > 0) and np.all(row[o:] == 0)
        assert np.all(np.diff(row[:o]) <= 0)                       # descending (pibasis.jl:97)
    assert b.A2Bmap.colptr[0] == 1 and b.A2Bmap.colptr[-1] == b.A2Bmap.nnz + 1
    assert np.all(b.A2Bmap.col_norms() > 0)                        # clean_pibasis! left no zero column


def test_cg_vs_sympy():
    from sympy.physics.quantum.cg import CG
    rng = philox(5)
    n = 0
    while n < 60:
        j1, j2 = int(rng.integers(0, 5)), int(rng.integers(0, 5))
        J = int(rng.integers(abs(j1 - j2), j1 + j2 + 1))
        m1, m2 = int(rng.integers(-j1, j1 + 1)), int(rng.integers(-j2, j2 + 1))
        M = m1 + m2
        if abs(M) > J:
            continue
        ref = float(CG(j1, m1, j2, m2, J, M).doit())
        assert abs(clebschgordan(j1, m1, j2, m2, J, M) - ref) < 1e-12
        n += 1
    assert clebschgordan(1, 1, 1, 1, 1, 2) == 0.0


def test_ylm_product_expansion():
    """Y_l1^m1 Y_l2^m2 = sum_L sqrt((2l1+1)(2l2+1)/(4 pi (2L+1))) C(l1 0 l2 0|L 0) C(l1 m1 l2 m2|L M) Y_L^M
    (test/test_cg.jl:53-77), evaluated with the oracle's harmonics."""
    from ace_jl_b200.descriptor import basis_descriptor
    from ace_jl_b200.onepbasis import index_y
    from oracle import Oracle
    o = Oracle(basis_descriptor(make_basis("inv_simple_3_6"), None))
    rng = philox(6)
    for _ in range(20):
        l1, l2 = int(rng.integers(0, 4)), int(rng.integers(0, 4))
        m1, m2 = int(rng.integers(-l1, l1 + 1)), int(rng.integers(-l2, l2 + 1))
        R = rng.standard_normal(3)
        Y = o.ylm(l1 + l2, R)
        lhs = Y[index_y(l1, m1) - 1] * Y[index_y(l2, m2) - 1]
        rhs = 0.0
        for Lc in range(abs(l1 - l2), l1 + l2 + 1):
            if abs(m1 + m2) > Lc:
                continue
            rhs += (math.sqrt((2 * l1 + 1) * (2 * l2 + 1) / (4 * math.pi * (2 * Lc + 1)))
                    * clebschgordan(l1, 0, l2, 0, Lc, 0) * clebschgordan(l1, m1, l2, m2, Lc, m1 + m2)
                    * Y[index_y(Lc, m1 + m2) - 1])
        assert abs(lhs - rhs) < 1e-12


def _parse_jl_complex(s):
    return complex(s.replace("im", "j").replace("+-", "-").replace(" ", ""))


@pytest.mark.skipif(not os.path.exists("/root/reference/src/eucl/cov_coeffs_dict.jl"), reason="reference tree not present")
def test_derived_seeds_match_reference_tables():
    txt = open("/root/reference/src/eucl/cov_coeffs_dict.jl").read()
    n = 0
    for m in re.finditer(r"\(l=(-?\d+), m=(-?\d+), mu=(-?\d+), i=(\d)\) => SVector\{3,ComplexF64\}\((.*?)\),?\n", txt):
        l, mm, mu, i = map(int, m.groups()[:4])
        ref = np.array([_parse_jl_complex(s) for s in m.group(5).split(", ")])
        assert l == 1 and np.abs(crmatrix(mm, mu, i) - ref).max() < 1e-15
        n += 1
    assert n == 27
    txt = open("/root/reference/src/eucl/equi_coeffs_dict.jl").read()
    n = 0
    for m in re.finditer(r"\(l=(-?\d+), m=(-?\d+), mu=(-?\d+), i=(\d), j=(\d)\) => SMatrix\{3, 3, ComplexF64, 9\}\((.*?)\),?\n", txt):
        l, mm, mu, i, j = map(int, m.groups()[:5])
        ref = np.array([_parse_jl_complex(s) for s in m.group(6).split(", ")]).reshape(3, 3).T
        assert np.abs(mrmatrix(l, mm, mu, i, j) - ref).max() < 1e-15
        n += 1
    assert n == 315


def test_discrete_jacobi_recursion_is_consistent():
    """The three-term coefficients reproduce the polynomials they were built from (orthpolys.jl:186-215)."""
    J = ace.discrete_jacobi(8, pcut=2, xcut=2.5, pin=0, xin=0.5, trans=ace.polytransform(2, 1.0))
    assert J.pl == 2 and J.pr == 0 and J.tl < J.tr              # decreasing transform: outer cutoff is the LEFT end
    t = J.tdf
    P = [J.A[0] * (t - J.tl) ** J.pl * (t - J.tr) ** J.pr]
    P.append((J.A[1] * t + J.B[1]) * P[0])
    for n in range(2, 8):
        P.append((J.A[n] * t + J.B[n]) * P[n - 1] + J.C[n] * P[n - 2])
    G = np.array([[np.dot(a, J.ww * b) for b in P] for a in P])
    assert np.abs(G - np.eye(8)).max() < 1e-10


def test_shared_O3_does_not_serve_stale_coupling_coefficients():
    """One O3() reused for bases of different properties (the reference's O3 is a stateless singleton,
    src/symmetrygroups.jl:66-94): the rpe cache must be keyed on the property, not on a recyclable object id."""
    import gc
    from ace_jl_b200.symmetrygroups import O3
    from ace_jl_b200.rotations3d import Rot3DCoeffs
    grp = O3()
    U1, _ = grp.rpe_basis(Rot3DCoeffs(ace.EuclideanMatrix()), (1, 1), (1, 1))
    gc.collect()
    U2, M2 = grp.rpe_basis(Rot3DCoeffs(ace.Invariant()), (1, 1), (1, 1))
    assert U1.shape[2] == 9 and U2.shape == (1, len(M2), 1)
    # and whole bases built through one shared group object equal the ones built with private groups
    B1p = RnYlm_1pbasis(maxdeg=4)
    Bsel = ace.SimpleSparseBasis(2, 4)
    shared = O3()
    a = ace.SymmetricBasis(ace.EuclideanVector(), RnYlm_1pbasis(maxdeg=4), shared, Bsel)
    b = ace.SymmetricBasis(ace.Invariant(), RnYlm_1pbasis(maxdeg=4), shared, Bsel)
    b0 = ace.SymmetricBasis(ace.Invariant(), B1p, Bsel)
    assert a.A2Bmap.ncomp == 3 and b.A2Bmap.ncomp == 1
    assert np.array_equal(b.A2Bmap.nzval, b0.A2Bmap.nzval) and np.array_equal(b.A2Bmap.rowval, b0.A2Bmap.rowval)


def test_mutating_a_basis_invalidates_its_device_handle_stamp():
    """sparsify! / clean_pibasis! change the tables in place (src/symmbasis.jl:204-236): a cached device handle is
    only reused while the table stamp is unchanged (api._handle_of)."""
    from ace_jl_b200.api import _table_stamp
    basis = ace.SymmetricBasis(ace.Invariant(), RnYlm_1pbasis(maxdeg=4), ace.SimpleSparseBasis(2, 4))
    s0 = _table_stamp(basis)
    assert _table_stamp(basis) == s0
    n0 = len(basis)
    basis.sparsify(keep=list(range(1, n0 // 2)))
    assert len(basis) < n0 and _table_stamp(basis) != s0
    s1 = _table_stamp(basis.pibasis)
    basis.pibasis.basis1p.set_spec(basis.pibasis.basis1p.spec)
    assert _table_stamp(basis.pibasis) != s1

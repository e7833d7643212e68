"""Pin the CPU oracle with every known-answer test the reference holds for this path.

  - closed-form Y_l^m, l <= 3, at random points and near the pole   (test/polynomials/test_ylm.jl:14-67)
  - gradient of Y_l^m by finite differences, values of evaluate == evaluate_ed (test_ylm.jl:158-183)
  - closed-form distance transforms, derivative by finite differences (test/transforms/test_transforms.jl:16-52)
  - radial basis derivative                                          (test/polynomials/test_orthpolys.jl:32-52)
  - one-hot categorical basis                                        (test/test_discrete.jl:38-45)
  - A(cfg) = sum_j A(X_j), permutation invariance                    (test/test_1pbasis.jl:45-53)
  - AA = naive prod A; first spec is the empty tuple                 (test/test_pibasis.jl:43-66)
  - B = A2Bmap * AA; invariance under O(3) x permutations; rank      (test/test_symmbasis.jl:42-107)
  - naive == product evaluator for evaluate / grad_config            (test/test_linearmodel.jl:47-78)
"""
import math

import numpy as np
import pytest

import ace_jl_b200 as ace
from ace_jl_b200.descriptor import basis_descriptor
from ace_jl_b200.utils import philox, rand_envs, rand_rot
from conftest import nspecies_of, relerr, rn_of
from oracle import Oracle


def explicit_shs(th, ph):
    s, c = math.sin, math.cos
    e = lambda a: complex(math.cos(a), math.sin(a))  # noqa: E731
    pi = math.pi
    return np.array([
        0.5 * math.sqrt(1 / pi),
        0.5 * math.sqrt(3 / (2 * pi)) * s(th) * e(-ph), 0.5 * math.sqrt(3 / pi) * c(th), -0.5 * math.sqrt(3 / (2 * pi)) * s(th) * e(ph),
        0.25 * math.sqrt(15 / (2 * pi)) * s(th) ** 2 * e(-2 * ph), 0.5 * math.sqrt(15 / (2 * pi)) * s(th) * c(th) * e(-ph),
        0.25 * math.sqrt(5 / pi) * (3 * c(th) ** 2 - 1), -0.5 * math.sqrt(15 / (2 * pi)) * s(th) * c(th) * e(ph),
        0.25 * math.sqrt(15 / (2 * pi)) * s(th) ** 2 * e(2 * ph),
        1 / 8 * math.sqrt(35 / pi) * s(th) ** 3 * e(-3 * ph), 1 / 4 * math.sqrt(105 / (2 * pi)) * s(th) ** 2 * c(th) * e(-2 * ph),
        1 / 8 * math.sqrt(21 / pi) * s(th) * (5 * c(th) ** 2 - 1) * e(-ph), 1 / 4 * math.sqrt(7 / pi) * (5 * c(th) ** 3 - 3 * c(th)),
        -1 / 8 * math.sqrt(21 / pi) * s(th) * (5 * c(th) ** 2 - 1) * e(ph), 1 / 4 * math.sqrt(105 / (2 * pi)) * s(th) ** 2 * c(th) * e(2 * ph),
        -1 / 8 * math.sqrt(35 / pi) * s(th) ** 3 * e(3 * ph)])


@pytest.fixture(scope="module")
def orc(request):
    from conftest import make_basis
    basis = make_basis("inv_simple_3_6")
    rng = philox(11)
    c = rng.random(len(basis)) - 0.5
    return basis, c, Oracle(basis_descriptor(basis, c.reshape(-1, 1)))


def test_ylm_closed_forms(orc):
    _, _, o = orc
    rng = philox(1)
    for _ in range(30):
        th, ph, r = rng.random() * math.pi, (rng.random() - 0.5) * 2 * math.pi, 0.1 + rng.random()
        R = r * np.array([math.sin(th) * math.cos(ph), math.sin(th) * math.sin(ph), math.cos(th)])
        assert np.abs(o.ylm(3, R) - explicit_shs(th, ph)).max() < 1e-14
        assert np.abs(o.ylm_ed(3, R)[0] - explicit_shs(th, ph)).max() < 1e-14


def test_ylm_near_pole(orc):
    _, _, o = orc
    rng = philox(2)
    for _ in range(30):
        th, ph = rng.random() * 1e-9, (rng.random() - 0.5) * 2 * math.pi
        if th < 1e-13:
            th = 0.0
        R = np.array([math.sin(th) * math.cos(ph), math.sin(th) * math.sin(ph), math.cos(th)])
        assert np.abs(o.ylm(3, R) - explicit_shs(th, ph)).max() < 1e-12


def test_ylm_gradients_fd(orc):
    _, _, o = orc
    rng = philox(3)
    for _ in range(10):
        R = rng.standard_normal(3)
        Y, dY = o.ylm_ed(5, R)
        assert np.abs(Y - o.ylm(5, R)).max() < 1e-14
        h = 1e-6
        fd = np.stack([(o.ylm(5, R + h * np.eye(3)[k]) - o.ylm(5, R - h * np.eye(3)[k])) / (2 * h) for k in range(3)], axis=1)
        assert np.abs(fd - dY).max() < 1e-8


def test_transform_closed_forms():
    # test/transforms/test_transforms.jl:21, 34, 49
    from conftest import make_basis
    for kind, f in (("inv_morse_2_6", lambda r: math.exp(-1.3 * (r / 1.1 - 1))),
                    ("inv_agnesi_2_6", lambda r: 1 / (1 + 0.5 * (r / 1.0) ** 3)),
                    ("inv_simple_3_6", lambda r: ((1 + 1.0) / (1 + r)) ** 2)):
        basis = make_basis(kind)
        o = Oracle(basis_descriptor(basis, None))
        tr = rn_of(basis).trans
        for r in np.linspace(0.3, 3.0, 17):
            assert abs(o.transform(r) - f(r)) < 1e-15 * max(1, abs(f(r)))
            assert abs(tr(r) - f(r)) < 1e-15 * max(1, abs(f(r)))
            h = 1e-6
            assert abs(o.transform_d(r) - (f(r + h) - f(r - h)) / (2 * h)) < 1e-8
            assert abs(tr.deriv(r) - o.transform_d(r)) < 1e-13
            assert abs(tr.inv(tr(r)) - r) < 1e-12


def test_radial_derivative_and_cutoff(orc):
    basis, _, o = orc
    rng = philox(4)
    for _ in range(10):
        R = rng.standard_normal(3)
        R *= (0.6 + 1.8 * rng.random()) / np.linalg.norm(R)
        P, dP = o.rn_ed(R)
        h = 1e-6
        fd = np.stack([(o.rn(R + h * np.eye(3)[k]) - o.rn(R - h * np.eye(3)[k])) / (2 * h) for k in range(3)], axis=1)
        assert np.abs(fd - dP).max() < 1e-7
    # beyond the cutoff every R_n is exactly zero (orthpolys.jl:41-46)
    assert np.all(o.rn(np.array([0.0, 0.0, rn_of(basis).meta["rcut"] + 0.1])) == 0.0)


def test_radial_basis_is_orthonormal(orc):
    """The construction (orthpolys.jl:153-218) yields <J_m, J_n> = delta_mn on its quadrature grid."""
    basis, _, o = orc
    Rn = rn_of(basis)
    vals = np.array([o.rn(np.array([0.0, 0.0, Rn.trans.inv(t)])) for t in Rn.R.tdf])
    G = vals.T @ (Rn.R.ww[:, None] * vals)
    assert np.abs(G - np.eye(len(Rn.R))).max() < 1e-9


def test_onehot_categorical():
    from conftest import make_basis
    basis = make_basis("species_3_5")
    o = Oracle(basis_descriptor(basis, None))
    b1p = basis.pibasis.basis1p
    R = np.array([[0.3, -0.5, 0.9]])
    qk = b1p.sym_index("q")
    cats = b1p.component(2).categories
    for q in range(1, 5):
        A = o.eval_A(R, [0, 1], [q])[0]
        for iA, b in enumerate(b1p.spec):
            if b[qk] != cats[q - 1]:
                assert A[iA] == 0
        assert np.abs(A).max() > 0
    with pytest.raises(Exception):
        o.eval_A(R, [0, 1], [7])


def test_A_is_sum_over_neighbours_and_permutation_invariant(orc):
    basis, _, o = orc
    rng = philox(5)
    R, off, _ = rand_envs(rng, rn_of(basis), 1, 9)
    A = o.eval_A(R, off)[0]
    Asum = sum(o.eval_A(R[j:j + 1], [0, 1])[0] for j in range(9))
    assert relerr(A, Asum) < 1e-14
    p = rng.permutation(9)
    assert relerr(o.eval_A(R[p], off)[0], A) < 1e-14


def test_AA_is_naive_product(orc):
    basis, _, o = orc
    rng = philox(6)
    R, off, _ = rand_envs(rng, rn_of(basis), 3, [4, 17, 30])
    A, AA = o.eval_A(R, off), o.eval_AA(R, off)
    spec = basis.pibasis.spec
    assert spec.orders[0] == 0 and spec.get_spec(1) == ()      # test_pibasis.jl:43-44
    for e in range(3):
        naive = np.array([np.prod([A[e, v - 1] for v in spec.get_spec(i)]) if spec.orders[i - 1] else 1.0
                          for i in range(1, len(spec) + 1)])
        assert relerr(AA[e], naive.real) < 1e-14


@pytest.mark.parametrize("kind", ["inv_simple_3_6", "inv_sparse_4_8", "species_3_5"])
def test_B_matches_A2B_and_is_invariant(kind, zoo):
    basis = zoo(kind)
    o = Oracle(basis_descriptor(basis, None))
    rng = philox(7)
    ns = nspecies_of(basis)
    R, off, sp = rand_envs(rng, rn_of(basis), 6, [3, 12, 30, 1, 8, 20], ns)
    B, AA = o.eval_B(R, off, sp), o.eval_AA(R, off, sp)
    A2B = basis.A2Bmap.todense()[:, :, 0]
    assert relerr(B[:, :, 0], (AA @ A2B.T).real) < 1e-10                 # test_symmbasis.jl:42-44
    for _ in range(3):
        Q = rand_rot(rng) * rng.choice([-1, 1])
        perm = np.concatenate([off[e] + rng.permutation(off[e + 1] - off[e]) for e in range(6)])
        Bq = o.eval_B((R @ Q.T)[perm], off, None if sp is None else sp[perm])
        assert relerr(Bq, B) < 1e-10                                      # test_symmbasis.jl:60-64


def test_B_rank_is_full(orc):
    basis, _, o = orc
    rng = philox(8)
    R, off, _ = rand_envs(rng, rn_of(basis), 3 * len(basis), 12)
    B = o.eval_B(R, off)[:, :, 0]
    assert np.linalg.matrix_rank(B) == len(basis)                        # test_symmbasis.jl:101-107


@pytest.mark.parametrize("kind,nprop", [("inv_simple_3_6", 1), ("inv_sparse_4_8", 3), ("euclvec_3_5", 1),
                                        ("euclmat_2_5", 2), ("species_3_5", 2)])
def test_naive_equals_product_evaluator(kind, nprop, zoo):
    basis = zoo(kind)
    rng = philox(9)
    c = rng.random((len(basis), nprop)) - 0.5
    o = Oracle(basis_descriptor(basis, c))
    R, off, sp = rand_envs(rng, rn_of(basis), 4, [3, 10, 1, 25], nspecies_of(basis))
    E, G = o.energy_forces(R, off, sp)
    En, Gn = o.naive_energy_forces(R, off, sp)
    assert relerr(E, En) < 1e-12 and relerr(G, Gn) < 1e-12               # test_linearmodel.jl:47-78
    # finite-difference check of the forces (test_linearmodel.jl, test_euclvec.jl:106-117)
    h, j, k = 1e-6, 5, 2
    Rp, Rm = R.copy(), R.copy()
    Rp[j, k] += h
    Rm[j, k] -= h
    fd = (o.energy(Rp, off, sp)[1] - o.energy(Rm, off, sp)[1]) / (2 * h)
    assert np.abs(fd - G[j, :, k, :]).max() < 1e-6 * max(1.0, np.abs(G).max())


@pytest.mark.parametrize("kind", ["euclvec_3_5", "euclmat_2_5"])
def test_equivariance(kind, zoo):
    """Q' B(QX) = B(X) for vectors (test_euclvec.jl:55-66), Q' B(QX) Q = B(X) for matrices
    (test_EuclideanMatrix.jl:58-72)."""
    basis = zoo(kind)
    o = Oracle(basis_descriptor(basis, None))
    rng = philox(10)
    R, off, _ = rand_envs(rng, rn_of(basis), 3, [5, 12, 20])
    B = o.eval_B(R, off)
    assert np.abs(B).max() > 1e-3
    for _ in range(3):
        Q = rand_rot(rng) * rng.choice([-1, 1])
        Bq = o.eval_B(R @ Q.T, off)
        if basis.phi.ncomp == 3:
            back = np.einsum("ji,ebj->ebi", Q, Bq)
        else:
            M = Bq.reshape(Bq.shape[0], Bq.shape[1], 3, 3).transpose(0, 1, 3, 2)   # column-major 3x3
            M = np.einsum("ai,enab,bj->enij", Q, M, Q)
            back = M.transpose(0, 1, 3, 2).reshape(Bq.shape)
        assert relerr(back, B) < 1e-10


@pytest.mark.parametrize("kind", ["inv_simple_3_6", "species_3_5", "inv_sparse_4_8"])
def test_adjoint_eval_d_equals_contracted_jacobian(kind, zoo):
    """adjoint_EVAL_D(w)_k = sum_j w_j . dB_k/dr_j: the product-evaluator shortcut equals the naive evaluator's
    formula (src/linearmodel.jl:160-169; test/test_admodel.jl uses it through the rrule).  Invariant properties only:
    the reference takes the real part *before* applying a complex A2Bmap (src/evaluator.jl:233), so for equivariant
    properties its two evaluators are not the same function; the GPU path follows the product evaluator."""
    basis = zoo(kind)
    o = Oracle(basis_descriptor(basis, None))
    rng = philox(12)
    R, off, sp = rand_envs(rng, rn_of(basis), 3, [4, 11, 19], nspecies_of(basis))
    w = rng.standard_normal((len(R), 3))
    adj = o.adjoint_eval_d(R, off, w, sp)
    _, dB = o.eval_dB(R, off, sp)                       # (sum J, nB, 3, ncomp)
    for e in range(3):
        ref = np.einsum("jx,jkxc->kc", w[off[e]:off[e + 1]], dB[off[e]:off[e + 1]])
        got = adj[e].real if basis.real else adj[e]
        assert relerr(got, ref) < 1e-11

"""Pin the CPU oracle with answers that do NOT come from this repository's reading of ACE.jl.

The reference ships no golden vectors for R_n, A, AA, B, E or forces and Julia is not
installed, so the oracle's absolute values are pinned here by third-party implementations instead:

  - complex Y_l^m for ALL l <= 8 (81 functions) against scipy.special.sph_harm_y and against
    mpmath.spherharm at 50 digits, at random points and near both poles -- extends the l <= 3 closed
    forms of test/polynomials/test_ylm.jl:14-67 to every l the BASELINE configs use (maxL = 8);
  - grad Y_l^m against a 50-digit mpmath derivative (test_ylm.jl:158-183 uses finite differences);
  - one whole small model re-evaluated with mpmath at 50 digits FROM THE SAME TABLES (recursion
    coefficients, 1p spec, iAA2iA, A2Bmap, c): R_n by the three-term recurrence of the table, Y_l^m by
    mpmath.spherharm (hypergeometric series -- independent of the ALP recursion the oracle and the
    kernels use), A = sum_j R_n Y_l^m, AA = prod A, B = Re(A2B AA), E = c . B, and forces = dE/dr_j by
    mpmath's high-precision numerical derivative.  A, AA, B, E and forces of the oracle must agree to
    1e-13 relative (test/test_linearmodel.jl:47-78 checks the same identities in Float64).

mpmath and scipy are third-party code; nothing below imports the product's kernels.
"""
import math

import mpmath as mp
import numpy as np
import pytest
from scipy.special import sph_harm_y

from ace_jl_b200.descriptor import basis_descriptor
from ace_jl_b200.utils import philox, rand_envs
from conftest import make_basis, relerr, rn_of
from oracle import Oracle

LMAX = 8


def _points(rng, n):
    pts = []
    for _ in range(n):
        th, ph, r = rng.random() * math.pi, (rng.random() - 0.5) * 2 * math.pi, 0.3 + 2 * rng.random()
        pts.append((th, ph, r))
    # near both poles (test_ylm.jl:52-67) and on the equator
    pts += [(1e-7, 0.3, 1.0), (math.pi - 1e-7, -2.0, 0.7), (math.pi / 2, 1.0, 1.3)]
    return pts


def _cart(th, ph, r):
    return r * np.array([math.sin(th) * math.cos(ph), math.sin(th) * math.sin(ph), math.cos(th)])


@pytest.fixture(scope="module")
def small():
    basis = make_basis("inv_simple_3_6")
    rng = philox(1234)
    c = rng.random(len(basis)) - 0.5
    return basis, c, Oracle(basis_descriptor(basis, c.reshape(-1, 1)))


def test_ylm_all_l_le_8_vs_scipy(small):
    _, _, o = small
    rng = philox(81)
    for th, ph, r in _points(rng, 40):
        R = _cart(th, ph, r)
        # the angles scipy sees are those of the Cartesian point the oracle sees (the round trip through sin/cos
        # moves a near-pole theta by ~1e-16, which matters at l = 8)
        th2, ph2 = math.atan2(math.hypot(R[0], R[1]), R[2]), math.atan2(R[1], R[0])
        ref = np.array([sph_harm_y(l, m, th2, ph2) for l in range(LMAX + 1) for m in range(-l, l + 1)])
        assert np.abs(o.ylm(LMAX, R) - ref).max() < 2e-14
        assert np.abs(o.ylm_ed(LMAX, R)[0] - ref).max() < 2e-14


def test_ylm_all_l_le_8_vs_mpmath_50_digits(small):
    _, _, o = small
    rng = philox(82)
    with mp.workdps(50):
        for th, ph, r in _points(rng, 6):
            R = _cart(th, ph, r)
            x, y, z = (mp.mpf(float(v)) for v in R)
            thm, phm = mp.atan2(mp.sqrt(x * x + y * y), z), mp.atan2(y, x)
            ref = np.array([complex(mp.spherharm(l, m, thm, phm)) for l in range(LMAX + 1) for m in range(-l, l + 1)])
            assert np.abs(o.ylm(LMAX, R) - ref).max() < 2e-14


def test_ylm_gradient_vs_mpmath_derivative(small):
    _, _, o = small
    rng = philox(83)
    L = 5
    with mp.workdps(40):
        for _ in range(3):
            R = rng.standard_normal(3)
            _, dY = o.ylm_ed(L, R)

            def Y(l, m, x, y, z):
                return mp.spherharm(l, m, mp.atan2(mp.sqrt(x * x + y * y), z), mp.atan2(y, x))

            x0 = [mp.mpf(float(v)) for v in R]
            for l, m in ((0, 0), (1, -1), (2, 1), (3, -3), (4, 2), (5, 0), (5, 5)):
                i = m + l + l * l
                for k in range(3):
                    def f(t, k=k, l=l, m=m):
                        xs = list(x0)
                        xs[k] = t
                        return Y(l, m, *xs)
                    ref = complex(mp.diff(f, x0[k]))
                    assert abs(dY[i, k] - ref) < 1e-12 * max(1.0, abs(ref))


def _mp_model(basis, c):
    """E(positions) of a LinearACEModel at mpmath precision, from the model's own tables."""
    b1p = basis.pibasis.basis1p
    Rn = rn_of(basis)
    P = Rn.R
    sym = b1p.symbols
    inn, il, im = sym.index("n"), sym.index("l"), sym.index("m")
    spec1 = [(b[inn], b[il], b[im]) for b in b1p.spec]
    orders = [int(v) for v in basis.pibasis.spec.orders]
    iAA2iA = np.asarray(basis.pibasis.spec.iAA2iA)
    A2B = basis.A2Bmap
    par = [mp.mpf(float(v)) for v in Rn.trans.c_params()]
    assert Rn.trans.kind == 1                   # polytransform: t = ((1 + r0) / (1 + r))^p
    rA, rB, rC = ([mp.mpf(float(v)) for v in a] for a in (P.A, P.B, P.C))
    tl, tr = mp.mpf(float(P.tl)), mp.mpf(float(P.tr))

    def radial(r):
        t = ((1 + par[1]) / (1 + r)) ** par[0]
        if (P.pl > 0 and t < tl) or (P.pr > 0 and t > tr):
            return [mp.mpf(0)] * len(rA)
        out = [rA[0] * (t - tl) ** P.pl * (t - tr) ** P.pr]
        out.append((rA[1] * t + rB[1]) * out[0])
        for n in range(2, len(rA)):
            out.append((rA[n] * t + rB[n]) * out[n - 1] + rC[n] * out[n - 2])
        return out

    def stages(X):
        """X: list of (x, y, z) mpf -> (A, AA, B, E)."""
        A = [mp.mpc(0)] * len(spec1)
        for (x, y, z) in X:
            r = mp.sqrt(x * x + y * y + z * z)
            th, ph = mp.atan2(mp.sqrt(x * x + y * y), z), mp.atan2(y, x)
            Rv = radial(r)
            ycache = {}
            for a, (n, l, m) in enumerate(spec1):
                if (l, m) not in ycache:
                    ycache[(l, m)] = mp.spherharm(l, m, th, ph)
                A[a] = A[a] + Rv[n - 1] * ycache[(l, m)]
        AA = []
        for i, o in enumerate(orders):
            v = mp.mpc(1)
            for t in range(o):
                v = v * A[int(iAA2iA[i, t]) - 1]
            AA.append(mp.re(v) if basis.pibasis.real else v)
        B = [[mp.mpf(0)] * A2B.ncomp for _ in range(A2B.m)]
        for j in range(A2B.n):
            for k in range(int(A2B.colptr[j]) - 1, int(A2B.colptr[j + 1]) - 1):
                for q in range(A2B.ncomp):
                    val = A2B.nzval[k, q]
                    B[int(A2B.rowval[k]) - 1][q] += mp.re(mp.mpc(float(val.real), float(val.imag)) * AA[j])
        E = mp.fsum(mp.mpf(float(ci)) * Bi[0] for ci, Bi in zip(c, B))
        return A, AA, B, E

    return stages


def test_small_model_vs_mpmath_50_digits(small):
    """A, AA, B, E and forces of one environment, oracle (Float64) vs a 50-digit evaluation of the same tables."""
    basis, c, o = small
    rng = philox(84)
    J = 4
    R, off, _ = rand_envs(rng, rn_of(basis), 1, J)
    with mp.workdps(50):
        stages = _mp_model(basis, c)
        X = [tuple(mp.mpf(float(v)) for v in row) for row in R]
        A, AA, B, E = stages(X)
        A = np.array([complex(v) for v in A])
        AA = np.array([float(v) for v in AA])
        B = np.array([float(v[0]) for v in B])
        assert relerr(o.eval_A(R, off)[0], A) < 1e-13
        assert relerr(o.eval_AA(R, off)[0], AA) < 1e-13
        assert relerr(o.eval_B(R, off)[0, :, 0], B) < 1e-13
        Eo, Go = o.energy_forces(R, off)
        assert abs(Eo[0, 0, 0] - float(E)) < 1e-13 * max(1.0, abs(float(E)))
        # forces: dE/dr_j by mpmath's numerical derivative at 50 digits (step 1e-12: error ~ 1e-24)
        G = np.zeros((J, 3))
        for j in range(J):
            for k in range(3):
                def f(t, j=j, k=k):
                    Y = [list(p) for p in X]
                    Y[j][k] = t
                    return stages([tuple(p) for p in Y])[3]
                G[j, k] = float(mp.diff(f, X[j][k], h=mp.mpf("1e-12")))
        assert relerr(Go[:, 0, :, 0], G) < 1e-12


def test_equivariant_basis_vs_mpmath_50_digits():
    """EuclideanVector basis (complex AA, 3-component complex A2Bmap, real B): A, AA and B vs 50-digit arithmetic."""
    basis = make_basis("euclvec_3_5")
    o = Oracle(basis_descriptor(basis, None))
    rng = philox(85)
    R, off, _ = rand_envs(rng, rn_of(basis), 1, 3)
    with mp.workdps(50):
        stages = _mp_model(basis, np.zeros(len(basis)))
        A, AA, B, _ = stages([tuple(mp.mpf(float(v)) for v in row) for row in R])
        assert relerr(o.eval_A(R, off)[0], np.array([complex(v) for v in A])) < 1e-13
        assert relerr(o.eval_AA(R, off)[0], np.array([complex(v) for v in AA])) < 1e-13
        assert relerr(o.eval_B(R, off)[0], np.array([[float(q) for q in v] for v in B])) < 1e-13

"""Clebsch-Gordan coefficients and generalised coupling coefficients (host-side, one-off).

Mirrors src/rotations3d.jl: ``clebschgordan`` (:108-137, exact integer arithmetic), the cached
``ClebschGordan`` (:145-162), the ``Rot3DCoeffs`` recursion (:191-248), the ``MRange`` iteration
(:36-66) and ``re_basis`` / ``compute_Al`` (:257-329).  The results are uploaded once per model as the
sparse ``A2Bmap``; nothing here runs on the GPU (BASELINE.json north_star).
"""
from __future__ import annotations

import itertools
import math
from decimal import Decimal, localcontext
from fractions import Fraction
from typing import Dict, List, Tuple

import numpy as np


def cg_conditions(j1, m1, j2, m2, J, M) -> bool:
    return (abs(j1 - j2) <= J <= j1 + j2) and (M == m1 + m2) and abs(m1) <= j1 and abs(m2) <= j2 and abs(M) <= J


def clebschgordan(j1, m1, j2, m2, J, M) -> float:
    """C_{j1 m1 j2 m2}^{J M}; the formula of rotations3d.jl:108-137 in exact rational arithmetic."""
    if not cg_conditions(j1, m1, j2, m2, J, M):
        return 0.0
    f = math.factorial
    N = Fraction((2 * J + 1) * f(j1 + m1) * f(j1 - m1) * f(j2 + m2) * f(j2 - m2) * f(J + M) * f(J - M),
                 f(j1 + j2 - J) * f(j1 - j2 + J) * f(-j1 + j2 + J) * f(j1 + j2 + J + 1))
    G = 0
    for k in range(max(0, j2 - J - m1, j1 - J + m2), min(j1 + j2 - J, j1 - m1, j2 + m2) + 1):
        G += (-1) ** k * math.comb(j1 + j2 - J, k) * math.comb(j1 - j2 + J, j1 - m1 - k) * math.comb(-j1 + j2 + J, j2 + m2 - k)
    # high-precision sqrt of the exact rational, one final rounding (the reference uses BigFloat)
    with localcontext() as ctx:
        ctx.prec = 80
        return float((Decimal(N.numerator) / Decimal(N.denominator)).sqrt() * G)


class ClebschGordan:
    def __init__(self):
        self.vals: Dict[tuple, float] = {}

    def __call__(self, j1, m1, j2, m2, J, M) -> float:
        if not cg_conditions(j1, m1, j2, m2, J, M):
            return 0.0
        key = (j1, m1, j2, m2, J, M)
        v = self.vals.get(key)
        if v is None:
            v = self.vals[key] = clebschgordan(j1, m1, j2, m2, J, M)
        return v


def mrange(phi, ll) -> List[tuple]:
    """All mm in prod_i [-l_i..l_i] (first index fastest) that pass coco_filter (:36-66)."""
    out = []
    for rev in itertools.product(*[range(-l, l + 1) for l in reversed(ll)]):
        mm = tuple(reversed(rev))
        if phi.coco_filter(ll, mm):
            out.append(mm)
    return out


class Rot3DCoeffs:
    """A(ll, mm, kk): recursion over the correlation order with a CG contraction (:191-248)."""

    def __init__(self, phi):
        self.phi = phi
        self.cg = ClebschGordan()
        self.vals: Dict[tuple, np.ndarray] = {}

    def __call__(self, ll: tuple, mm: tuple, kk: tuple) -> np.ndarray:
        N = len(ll)
        if N == 1:
            return self.phi.coco_init(ll[0], mm[0], kk[0])
        key = (ll, mm, kk)
        v = self.vals.get(key)
        if v is None:
            v = self.vals[key] = self._compute_val(ll, mm, kk)
        return v

    def _compute_val(self, ll, mm, kk) -> np.ndarray:
        N = len(ll)
        val = self.phi.coco_zeros()
        jmin = max(abs(ll[N - 2] - ll[N - 1]), abs(kk[N - 2] + kk[N - 1]), abs(mm[N - 2] + mm[N - 1]))
        jmax = ll[N - 2] + ll[N - 1]
        for j in range(jmin, jmax + 1):
            cgk = self.cg(ll[N - 2], kk[N - 2], ll[N - 1], kk[N - 1], j, kk[N - 2] + kk[N - 1])
            cgm = self.cg(ll[N - 2], mm[N - 2], ll[N - 1], mm[N - 1], j, mm[N - 2] + mm[N - 1])
            if cgk * cgm != 0:
                llpp = ll[:N - 2] + (j,)
                mmpp = mm[:N - 2] + (mm[N - 2] + mm[N - 1],)
                kkpp = kk[:N - 2] + (kk[N - 2] + kk[N - 1],)
                val = val + cgk * cgm * self(llpp, mmpp, kkpp)
        return val


def rank_rtol(S: np.ndarray, rtol: float) -> int:
    """rank(Diagonal(S), rtol=...) of LinearAlgebra: count of values above rtol * max."""
    if len(S) == 0:
        return 0
    tol = rtol * float(np.max(S))
    return int(np.sum(S > tol))


def compute_Al(A: Rot3DCoeffs, ll: tuple) -> Tuple[np.ndarray, List[tuple]]:
    """Rows of candidate coupling coefficients, one block of ``numcc`` rows per kk (:277-329).

    Returns CC with shape (nrows, len(Mll), ncomp) and the list Mll.
    """
    phi = A.phi
    Mll = mrange(phi, ll)
    if len(Mll) == 0:
        return np.zeros((0, 0, phi.ncomp), dtype=np.complex128), Mll
    CC = np.zeros((len(Mll) * phi.numcc, len(Mll), phi.ncomp), dtype=np.complex128)
    for ik, kk in enumerate(Mll):
        for im, mm in enumerate(Mll):
            if phi.coco_filter(ll, mm, kk):
                CC[ik * phi.numcc:(ik + 1) * phi.numcc, im, :] = A(ll, mm, kk)
    return CC, Mll


def re_basis(A: Rot3DCoeffs, ll: tuple) -> Tuple[np.ndarray, List[tuple]]:
    """Rotation-equivariant basis by SVD of the Gramian, rank cut at rtol=1e-7 (:257-273)."""
    CC, Mll = compute_Al(A, ll)
    if CC.shape[0] == 0:
        return np.zeros((0, len(Mll), A.phi.ncomp), dtype=np.complex128), Mll
    G = A.phi.coco_dot(CC, CC)
    if not np.any(G.imag):
        G = G.real  # real Gramian -> real singular vectors (Invariant couplings stay real)
    U, S, _ = np.linalg.svd(G)
    rk = rank_rtol(S, 1e-7)
    Ured = np.sqrt(S[:rk])[:, None] * np.conj(U[:, :rk]).T
    Ure = np.einsum("ij,jmc->imc", Ured, CC)
    return Ure, Mll

"""Host-side construction of the orthogonal-polynomial radial basis (coefficients only).

The recursion coefficients are what crosses the C ABI; they are evaluated on the GPU by
``csrc/aceb200.cu``.  Construction follows the discrete Stieltjes procedure of the reference
(src/polynomials/orthpolys.jl:153-218, 340-349).  It agrees with a Julia-built basis to round-off
(~1e-15 relative), not bitwise: summation order differs.  When Julia is the
host, its own ``A,B,C,tl,tr`` are passed instead and this file is not involved.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


def _fcut(pl, tl, pr, tr, t):
    """Envelope (t-tl)^pl (t-tr)^pr, zero outside the cutoff side(s) with p>0 (orthpolys.jl:32-46)."""
    t = np.asarray(t, dtype=np.float64)
    v = (t - tl) ** pl * (t - tr) ** pr
    out = np.zeros_like(v)
    inside = ~(((pl > 0) & (t < tl)) | ((pr > 0) & (t > tr)))
    out[inside] = v[inside]
    return out


@dataclass
class OrthPolyBasis:
    """Recursion J1 = A1 fcut, J2 = (A2 t + B2) J1, Jn = (An t + Bn) J(n-1) + Cn J(n-2)
    (orthpolys.jl:65-99)."""

    pl: int
    tl: float
    pr: int
    tr: float
    A: np.ndarray
    B: np.ndarray
    C: np.ndarray
    tdf: np.ndarray
    ww: np.ndarray

    def __len__(self):
        return len(self.A)


def orthpolybasis(N: int, pcut: int, tcut: float, pin: int, tin: float, tdf, ww=None) -> OrthPolyBasis:
    """orthpolys.jl:153-218."""
    assert pcut >= 0 and pin >= 0 and N > 0
    tdf = np.asarray(tdf, dtype=np.float64)
    ww = np.ones_like(tdf) if ww is None else np.asarray(ww, dtype=np.float64)
    if tcut < tin:
        tl, tr, pl, pr = tcut, tin, pcut, pin
    else:
        tl, tr, pl, pr = tin, tcut, pin, pcut
    A = np.zeros(N)
    B = np.zeros(N)
    C = np.zeros(N)
    ww = ww / ww.sum()

    def dotw(f1, f2):
        return float(np.dot(f1, ww * f2))

    _J1 = _fcut(pl, tl, pr, tr, tdf)
    a = np.sqrt(dotw(_J1, _J1))
    A[0] = 1.0 / a
    J1 = A[0] * _J1
    if N > 1:
        b = dotw(tdf * J1, J1)
        _J2 = (tdf - b) * J1
        a = np.sqrt(dotw(_J2, _J2))
        A[1] = 1.0 / a
        B[1] = -b / a
        J2 = (A[1] * tdf + B[1]) * J1
        Jprev, Jpprev = J2, J1
    for n in range(2, N):
        b = dotw(tdf * Jprev, Jprev)
        c = dotw(tdf * Jprev, Jpprev)
        _J = (tdf - b) * Jprev - c * Jpprev
        a = np.sqrt(dotw(_J, _J))
        A[n] = 1.0 / a
        B[n] = -b / a
        C[n] = -c / a
        Jprev, Jpprev = _J / a, Jprev
    return OrthPolyBasis(int(pl), float(tl), int(pr), float(tr), A, B, C, tdf.copy(), ww)


def discrete_jacobi(N: int, *, pcut=0, xcut=1.0, pin=0, xin=-1.0, Nquad=None, trans=None) -> OrthPolyBasis:
    """orthpolys.jl:340-349: uniform quadrature nodes in the transformed variable."""
    if Nquad is None:
        Nquad = max(300, 3 * N)
    t = (lambda x: x) if trans is None else trans
    tcut = t(xcut)
    tin = t(xin)
    tl, tr = min(tin, tcut), max(tin, tcut)
    dt = (tr - tl) / Nquad
    tdf = np.linspace(tl + dt / 2, tr - dt / 2, Nquad)
    return orthpolybasis(N, pcut, tcut, pin, tin, tdf)

"""Symmetry groups: which AA functions are reference functions and how they couple.

Mirrors src/symmetrygroups.jl: ``NoSym`` (:27-43), ``O3`` (:66-116), ``rpe_basis`` (:119-128) and
the permutation Gramian ``_gramian`` (:131-146).  Host-side and one-off.
"""
from __future__ import annotations

import itertools
from typing import Dict, List, Tuple

import numpy as np

from .rotations3d import Rot3DCoeffs, rank_rtol, re_basis


class NoSym:
    """No symmetrisation beyond permutations (symmetrygroups.jl:27-43)."""

    def is_refbasisfcn(self, ls, ms) -> bool:
        return True


class O3:
    """Single O(3) acting on the (l, m) channel (symmetrygroups.jl:66-94)."""

    def __init__(self, lsym: str = "l", msym: str = "m"):
        self.lsym, self.msym = lsym, msym
        self._re_cache: Dict[tuple, tuple] = {}
        self._rpe_cache: Dict[tuple, tuple] = {}

    def is_refbasisfcn(self, ls, ms) -> bool:
        return all(m == 0 for m in ms)  # :91

    # ---- rpe_basis with caching: the result depends on ll and on the equality pattern of nn only
    def rpe_basis(self, rotc: Rot3DCoeffs, nn: tuple, ll: tuple) -> Tuple[np.ndarray, List[tuple]]:
        """symmetrygroups.jl:119-128.  Returns U (nrows, nM, ncomp) and the list of mm tuples."""
        canon = {}
        pattern = tuple(canon.setdefault(n, len(canon)) for n in nn)
        # keyed on the PROPERTY (type + component count), never on id(rotc): a Rot3DCoeffs is a short-lived local of
        # SymmetricBasis._build_A2B and CPython reuses ids, so an O3() shared by two bases of different properties
        # (the reference's O3 is a stateless singleton) would otherwise be served stale coefficients
        pkey = (type(rotc.phi).__name__, getattr(rotc.phi, "ncomp", None), repr(getattr(rotc.phi, "__dict__", None)))
        key = (pkey, ll, pattern)
        hit = self._rpe_cache.get(key)
        if hit is not None:
            return hit
        rkey = (pkey, ll)
        if rkey not in self._re_cache:
            self._re_cache[rkey] = re_basis(rotc, ll)
        Ure, Mre = self._re_cache[rkey]
        if Ure.shape[0] == 0:
            out = (Ure, Mre)
        else:
            G = _gramian(rotc.phi, pattern, ll, Ure, Mre)
            if not np.any(G.imag):
                G = G.real
            U, S, _ = np.linalg.svd(G)
            rk = rank_rtol(S, 1e-7)
            Urpe = np.sqrt(S[:rk])[:, None] * np.conj(U[:, :rk]).T
            out = (np.einsum("ij,jmc->imc", Urpe, Ure), Mre)
        self._rpe_cache[key] = out
        return out


def _gramian(phi, nn, ll, Ure: np.ndarray, Mre: List[tuple]) -> np.ndarray:
    """Sum over the permutations that fix (nn, ll) of <Ure[:, mm1], Ure[:, mm1[sigma]]>
    (symmetrygroups.jl:131-146)."""
    N = len(nn)
    nre = Ure.shape[0]
    G = np.zeros((nre, nre), dtype=np.complex128)
    pos = {mm: i for i, mm in enumerate(Mre)}
    for sigma in itertools.permutations(range(N)):
        if tuple(nn[s] for s in sigma) != tuple(nn) or tuple(ll[s] for s in sigma) != tuple(ll):
            continue
        i1, i2 = [], []
        for iU1, mm1 in enumerate(Mre):
            iU2 = pos.get(tuple(mm1[s] for s in sigma))
            if iU2 is not None:
                i1.append(iU1)
                i2.append(iU2)
        if i1:
            G += phi.coco_dot(Ure[:, i1, :], Ure[:, i2, :])
    return G

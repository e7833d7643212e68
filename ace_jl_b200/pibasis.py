"""Permutation-invariant product basis AA = prod A: specification tables.

Mirrors ``PIBasisSpec`` (src/pibasis.jl:10-13, 35-101) and ``PIBasis`` (:145-207).  ``orders`` and
``iAA2iA`` are the integer tables that cross the C ABI (1-based, rows descending, zero padded).
Evaluation (:258-432) is on the GPU; see ``api.py``.
"""
from __future__ import annotations

from typing import Callable, List, Optional

import numpy as np

from .onepbasis import COMP_YLM, Product1pBasis
from .selectors import DownsetBasisSelector, gensparse, init1pspec
from .symmetrygroups import O3


class PIBasisSpec:
    def __init__(self, orders: np.ndarray, iAA2iA: np.ndarray):
        self.orders = np.asarray(orders, dtype=np.int32)
        self.iAA2iA = np.asarray(iAA2iA, dtype=np.int32).reshape(len(self.orders), -1)

    def __len__(self):
        return len(self.orders)

    @property
    def maxcorrorder(self):
        return self.iAA2iA.shape[1]

    @classmethod
    def from_tuples(cls, AAspec: List[tuple]):
        """pibasis.jl:90-101: rows are ``reverse(vv)`` (descending, zeros last)."""
        nu = len(AAspec[0])
        iAA2iA = np.zeros((len(AAspec), nu), dtype=np.int32)
        orders = np.zeros(len(AAspec), dtype=np.int32)
        for i, vv in enumerate(AAspec):
            iAA2iA[i, :] = vv[::-1]
            orders[i] = sum(1 for v in vv if v != 0)
        return cls(orders, iAA2iA)

    def get_spec(self, i: int) -> tuple:
        """1-based i -> tuple of A indices (pibasis.jl:104)."""
        return tuple(int(v) for v in self.iAA2iA[i - 1, :self.orders[i - 1]])

    def sparsify(self, Ikeep):
        Ikeep = np.asarray(Ikeep, dtype=np.int64)
        return PIBasisSpec(self.orders[Ikeep], self.iAA2iA[Ikeep, :])


def _lm_of_spec(basis1p: Product1pBasis, symgrp):
    lk = basis1p.sym_index(symgrp.lsym)
    mk = basis1p.sym_index(symgrp.msym)
    ls = np.array([b[lk] for b in basis1p.spec], dtype=np.int64)
    ms = np.array([b[mk] for b in basis1p.spec], dtype=np.int64)
    return ls, ms


def build_pibasis_spec(basis1p: Product1pBasis, symgrp, Bsel: DownsetBasisSelector, *,
                       property=None, filterfun: Callable = lambda bb: True,
                       init1pbasis: bool = True) -> PIBasisSpec:
    """pibasis.jl:35-87."""
    if init1pbasis:
        init1pspec(basis1p, Bsel)
    Aspec = basis1p.get_spec()
    lev1 = [Bsel.level(b, basis1p) for b in Aspec]
    if any(lev1[i] > lev1[i + 1] for i in range(len(lev1) - 1)):
        raise ValueError("PIBasisSpec : AAspec construction failed because Aspec is not sorted by degree.")
    if isinstance(symgrp, O3):
        ls, ms = _lm_of_spec(basis1p, symgrp)
    else:
        ls = ms = None

    def tup2b(vv):
        return [Aspec[v - 1] for v in vv if v != 0]

    def admissible(bb):
        return Bsel.level_bb(bb, basis1p) <= Bsel.maxlevel(bb, basis1p)

    # fast integer path for the selectors whose level is a p=1 sum of 1p levels
    additive = getattr(Bsel, "p", 1) == 1

    def admissible_idx(vv):
        nz = [v for v in vv if v != 0]
        if additive:
            lev = sum(lev1[v - 1] for v in nz) if nz else 0
            return lev <= Bsel.maxlevel(nz, basis1p)
        return admissible(tup2b(vv))

    def filter_idx(vv):
        nz = [v for v in vv if v != 0]
        bb = [Aspec[v - 1] for v in nz]
        if not filterfun(bb):
            return False
        if not Bsel.filter(bb, basis1p):
            return False
        if property is not None and ls is not None:
            return property.filter([int(ls[v - 1]) for v in nz], [int(ms[v - 1]) for v in nz])
        return True

    nu = Bsel.maxorder
    AAspec = gensparse(NU=nu, maxvv=[len(Aspec)] * nu, admissible=admissible_idx,
                       filter=filter_idx, tup2b=lambda vv: vv, ordered=True)
    return PIBasisSpec.from_tuples(AAspec)


class PIBasis:
    """pibasis.jl:145-149.  ``real`` is True when AA is stored as its real part (Invariant)."""

    def __init__(self, basis1p: Product1pBasis, spec_or_symgrp, Bsel: Optional[DownsetBasisSelector] = None, *,
                 isreal: bool = False, **kwargs):
        self.basis1p = basis1p
        if isinstance(spec_or_symgrp, PIBasisSpec):
            self.spec = spec_or_symgrp
        else:
            symgrp = spec_or_symgrp
            if Bsel is None:  # PIBasis(basis1p, Bsel): default group O3 (pibasis.jl:162-163)
                symgrp, Bsel = O3(), spec_or_symgrp
            self.spec = build_pibasis_spec(basis1p, symgrp, Bsel, **kwargs)
        self.real = bool(isreal)

    def __len__(self):
        return len(self.spec)

    @property
    def maxcorrorder(self):
        return self.spec.maxcorrorder

    def get_spec(self, i=None):
        """AA index -> list of 1p basis functions (pibasis.jl:171-175)."""
        if i is None:
            return [self.get_spec(k) for k in range(1, len(self) + 1)]
        return [self.basis1p.get_spec(v) for v in self.spec.get_spec(i)]

    def sparsify(self, Ikeep):
        self._version = getattr(self, "_version", 0) + 1
        self.spec = self.spec.sparsify(Ikeep)
        return self

    def clean_1pbasis(self):
        """Drop unused 1p functions and renumber (pibasis.jl:117-127, 194-207)."""
        used = np.unique(self.spec.iAA2iA[self.spec.iAA2iA > 0])
        keep = [self.basis1p.spec[v - 1] for v in used]
        new_inds = self.basis1p.sparsify(keep)
        tab = self.spec.iAA2iA
        nz = tab > 0
        newtab = tab.copy()
        newtab[nz] = new_inds[tab[nz] - 1]
        assert np.all(newtab[nz] > 0)
        self._version = getattr(self, "_version", 0) + 1
        self.spec = PIBasisSpec(self.spec.orders, newtab)
        return self

"""Basis selectors and down-set enumeration (host-side integer logic).

Mirrors src/basisselectors.jl (``SimpleSparseBasis`` :106-124, ``SparseBasis`` :144-198,
``CategorySparseBasis`` :212-284, ``MaxBasis`` :79-91, ``NoConstant`` :290) and
src/sparsegrids.jl (``init1pspec!`` :9-31, ``gensparse`` :55-148).  The output of this file is
integer tables; they must be identical to the reference's.
"""
from __future__ import annotations

import itertools
import math
from typing import Callable, Dict, List, Sequence

import numpy as np

from .onepbasis import Product1pBasis


def _pnorm(x: Sequence[float], p: float) -> float:
    """LinearAlgebra.norm(x, p) for the p values a selector can carry."""
    if p == 1:
        return float(sum(abs(v) for v in x))
    if p == 2:
        return math.sqrt(sum(v * v for v in x))
    if math.isinf(p):
        return float(max(abs(v) for v in x))
    return float(sum(abs(v) ** p for v in x) ** (1.0 / p))


class DownsetBasisSelector:
    maxorder: int

    def level1(self, b, basis):  # basisselectors.jl:68-73
        return self.level(b, basis)

    def maxlevel1(self, basis):
        return self.maxlevel_all(basis)

    def filter(self, bb, basis) -> bool:  # basisselectors.jl:87
        return True


class MaxBasis(DownsetBasisSelector):
    def __init__(self, maxorder: int):
        self.maxorder = int(maxorder)

    def level(self, b, basis):
        return basis.degree(b)

    def level_bb(self, bb, basis):
        return 0 if len(bb) == 0 else sum(basis.degree(b) for b in bb)

    def maxlevel(self, bb, basis):
        return math.inf

    def maxlevel_all(self, basis):
        return math.inf


class SimpleSparseBasis(DownsetBasisSelector):
    """Total degree sum((n-1)+l) <= maxlevel (basisselectors.jl:106-124)."""

    def __init__(self, maxorder: int, maxlevel: float):
        self.maxorder = int(maxorder)
        self._maxlevel = float(maxlevel)

    def level(self, b, basis):
        return basis.degree(b)

    def level_bb(self, bb, basis):
        return 0 if len(bb) == 0 else sum(basis.degree(b) for b in bb)

    def maxlevel(self, bb, basis):
        return self._maxlevel

    def maxlevel_all(self, basis):
        return self._maxlevel


class SparseBasis(DownsetBasisSelector):
    """Weighted p-norm of weighted degrees with per-order max levels (basisselectors.jl:144-198)."""

    def __init__(self, *, maxorder: int, p=1, weight: Dict[str, float] = None,
                 default_maxdeg=None, maxlevels: Dict = None):
        if (default_maxdeg is None) == (maxlevels is None):
            raise ValueError("Either both or neither optional arguments `maxlevels` and "
                             "`default_maxdeg` were provided.")
        self.maxorder = int(maxorder)
        self.weight = dict(weight) if weight is not None else {"l": 1.0, "n": 1.0}
        self.maxlevels = {"default": float(default_maxdeg)} if maxlevels is None else dict(maxlevels)
        self.p = float(p)

    def level(self, b, basis):
        return basis.degree(b, self.weight)

    def level_bb(self, bb, basis):
        return 0.0 if len(bb) == 0 else _pnorm([self.level(b, basis) for b in bb], self.p)

    def maxlevel_ord(self, order: int):
        return self.maxlevels[order] if order in self.maxlevels else self.maxlevels["default"]

    def maxlevel(self, bb, basis):
        return self.maxlevel_ord(len(bb))

    def maxlevel_all(self, basis):
        return max(self.maxlevel_ord(o) for o in range(1, self.maxorder + 1))


class CategorySparseBasis(SparseBasis):
    """SparseBasis plus within-category order constraints and category weights
    (basisselectors.jl:212-284)."""

    def __init__(self, isym: str, categories, *, maxorder, p=1, weight=None, default_maxdeg=None,
                 maxlevels=None, minorder_dict=None, maxorder_dict=None, weight_cat=None):
        super().__init__(maxorder=maxorder, p=p, weight=weight or {}, default_maxdeg=default_maxdeg,
                         maxlevels=maxlevels)
        self.isym = isym
        self.minorder_dict = dict(minorder_dict or {})
        self.maxorder_dict = dict(maxorder_dict or {})
        self.weight_cat = dict(weight_cat) if weight_cat is not None else {c: 1.0 for c in categories}

    def level(self, b, basis):
        return basis.degree(b, self.weight) * self.weight_cat[b[basis.sym_index(self.isym)]]

    def filter(self, bb, basis) -> bool:
        if isinstance(bb, tuple):  # a single one-particle function: always kept (:253)
            return True
        k = basis.sym_index(self.isym)

        def num(s):
            return sum(1 for b in bb if b[k] == s)

        return (all(num(s) >= v for s, v in self.minorder_dict.items())
                and all(num(s) <= v for s, v in self.maxorder_dict.items()))


class NoConstant:
    """Filter removing the order-0 function (basisselectors.jl:290-293)."""

    def __call__(self, bb) -> bool:
        return len(bb) > 0


def init1pspec(B1p: Product1pBasis, Bsel: DownsetBasisSelector = None) -> Product1pBasis:
    """Enumerate, filter and (stably) sort the 1p basis by level (sparsegrids.jl:9-31).

    CartesianIndices runs the *first* symbol fastest; Python's itertools.product runs the last
    fastest, so the ranges are reversed going in and the tuples reversed coming out.
    """
    Bsel = MaxBasis(1) if Bsel is None else Bsel
    syms = B1p.symbols
    rgs = B1p.indexrange()
    maxlev = Bsel.maxlevel1(B1p)
    spec = []
    for Jrev in itertools.product(*[rgs[s] for s in reversed(syms)]):
        b = tuple(reversed(Jrev))
        if not B1p.isadmissible(b):
            continue
        if not Bsel.filter(b, B1p):
            continue
        if Bsel.level1(b, B1p) <= maxlev:
            spec.append(b)
    spec.sort(key=lambda b: Bsel.level(b, B1p))  # Python's sort is stable, like Julia's MergeSort
    return B1p.set_spec(spec)


def gensparse(*, NU: int, maxvv: Sequence[int], admissible: Callable, filter: Callable,
              tup2b: Callable = lambda vv: vv, ordered: bool = True, minvv=None) -> List[tuple]:
    """All admissible index tuples of a down-set, in the order ``gensparse`` of the reference visits them
    (src/sparsegrids.jl:55-148): lexicographic with the FIRST position most significant, ``vv[i] == 0`` = "no factor";
    with ``ordered`` only non-decreasing tuples.  The order matters: it is the order of the AA functions.

    Written as a depth-first walk over prefixes.  For a fixed prefix, position ``p`` runs upwards from its start value
    and stops at the first value whose *minimal completion* (the value repeated to the end if ``ordered``, zeros
    otherwise) is inadmissible -- in a down-set every larger tuple is then inadmissible too, which is the property the
    reference's increment / back-track loop relies on as well.
    """
    start = [0] * NU if minvv is None else [int(v) for v in minvv]
    limit = list(maxvv)
    kept: List[tuple] = []

    def ok(vv) -> bool:
        return all(v <= mx for v, mx in zip(vv, limit)) and bool(admissible(tup2b(vv)))

    if NU == 0:
        return [()] if ok([]) and filter(tup2b([])) else kept
    first = list(start)
    if not ok(first):
        raise ValueError("gensparse: the smallest index tuple is inadmissible, so the basis would be empty")

    def walk(p: int, vv: list) -> None:
        """vv[:p] is fixed and vv[p:] is the minimal completion of vv[p], already known to be admissible."""
        while True:
            if p == NU - 1:
                if filter(tup2b(vv)):
                    kept.append(tuple(vv))
            else:
                walk(p + 1, list(vv))
            vv[p] += 1
            for k in range(p + 1, NU):
                vv[k] = vv[p] if ordered else 0
            if not ok(vv):
                return

    walk(0, first)
    if ordered:
        assert all(all(t[i] <= t[i + 1] for i in range(NU - 1)) for t in kept) and len(set(kept)) == len(kept)
    return kept

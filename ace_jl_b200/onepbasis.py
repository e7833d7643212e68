"""One-particle basis: components and their product (host-side specification only).

Mirrors ``B1pComponent`` (src/b1pcomponent.jl:41-51), its wrappers ``Rn1pBasis``
(src/b1pcomponents/Rn.jl:17-29), ``Ylm1pBasis`` (src/b1pcomponents/Ylm.jl:18-26),
``Categorical1pBasis`` (src/discrete1pbasis.jl:58-131) and ``Product1pBasis``
(src/product_1pbasis.jl:5-8, 279-374).  Only the *specification* lives here: which component
functions exist, their degrees, and the ``indices`` table of the product.  Evaluation is done by
the CUDA library (``evaluate``/``evaluate_d``/``evaluate_ed`` in ``api.py`` call the C ABI).

A one-particle basis function ``b`` is a tuple of values in the order of ``Product1pBasis.symbols``,
e.g. ``(n, l, m)`` or ``(q, n, l, m)``; this plays the role of the reference's NamedTuple.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np

from .orthpolys import OrthPolyBasis
from .transforms import Lambda

COMP_RN, COMP_YLM, COMP_CAT = 0, 1, 2  # shared with include/aceb200.h


def idx2lm(i: int) -> Tuple[int, int]:
    """1-based flat index -> (l, m), sphericalharmonics.jl:104-108."""
    l = int(math.floor(math.sqrt(i - 1) + 1e-10))
    m = i - (l + l * l + 1)
    return l, m


def index_y(l: int, m: int) -> int:
    """sphericalharmonics.jl:102 (1-based)."""
    return m + l + l * l + 1


class _Component:
    kind: int
    symbols: Tuple[str, ...]
    spec: List[tuple]
    degrees: List[int]
    label: str

    def __len__(self):
        return len(self.spec)

    def _build_inv(self):
        self.invspec = {b: i + 1 for i, b in enumerate(self.spec)}

    def indexrange(self):
        """b1pcomponent.jl:114-119."""
        out = {}
        for k, s in enumerate(self.symbols):
            vals = [b[k] for b in self.spec]
            out[s] = list(range(min(vals), max(vals) + 1))
        return out

    def get_index(self, b: tuple) -> int:
        """1-based index of the component function used by b (b1pcomponent.jl:124-131)."""
        if b not in self.invspec:
            raise KeyError(f"B1pComponent ({self.label}): can't find {b} in spec")
        return self.invspec[b]

    def isadmissible(self, b: tuple) -> bool:
        return b in self.invspec

    def degree(self, b: tuple) -> int:
        return self.degrees[self.invspec[b] - 1]


class Rn1pBasis(_Component):
    """R_n(|rr|) = P_n(t(|rr|)); degrees 0..N-1 (Rn.jl:17-29)."""

    kind = COMP_RN

    def __init__(self, R: OrthPolyBasis, trans: Lambda, varsym="rr", nsym="n", label=None):
        self.R = R
        self.trans = trans
        self.varsym = varsym
        self.symbols = (nsym,)
        self.spec = [(i,) for i in range(1, len(R) + 1)]
        self.degrees = list(range(len(R)))
        self.label = label or f"R{nsym}"
        rl, rr = trans.inv(R.tl), trans.inv(R.tr)
        self.meta = {"rin": min(rl, rr), "rcut": max(rl, rr)}
        self._build_inv()


class Ylm1pBasis(_Component):
    """Complex spherical harmonics up to L in the order i -> idx2lm(i) (Ylm.jl:18-26)."""

    kind = COMP_YLM

    def __init__(self, L: int, varsym="rr", lsym="l", msym="m", label=None):
        self.L = int(L)
        self.varsym = varsym
        self.symbols = (lsym, msym)
        self.spec = [idx2lm(i) for i in range(1, (L + 1) ** 2 + 1)]
        self.degrees = [b[0] for b in self.spec]
        self.label = label or f"Y{lsym}{msym}"
        self._build_inv()


class Categorical1pBasis(_Component):
    """One-hot delta(u - U_q) over a list of categories; degree 0 (discrete1pbasis.jl:58-131)."""

    kind = COMP_CAT

    def __init__(self, categories: Sequence, varsym="mu", idxsym="q", label=None):
        cats = list(categories)
        if len(set(type(c) for c in cats)) != 1:
            raise TypeError("`SList` can only contain a single type")  # discrete1pbasis.jl:17-19
        self.categories = cats
        self.varsym = varsym
        self.symbols = (idxsym,)
        self.spec = [(c,) for c in cats]
        self.degrees = [0] * len(cats)
        self.label = label or f"C{idxsym}"
        self._build_inv()

    def indexrange(self):
        return {self.symbols[0]: list(self.categories)}  # discrete1pbasis.jl:117

    def val2i(self, val) -> int:
        """discrete1pbasis.jl:33-40; raises for an unknown category."""
        for j, c in enumerate(self.categories):
            if c == val:
                return j + 1
        raise ValueError(f"val = {val} not found in this list")


class Product1pBasis:
    """phi_v(X) = prod_i B_i[indices[v][i]](X) (product_1pbasis.jl:5-8)."""

    def __init__(self, bases: Sequence[_Component], indices=None, spec=None):
        self.bases = tuple(bases)
        # union of component symbols in component order (product_1pbasis.jl:279-281)
        syms: List[str] = []
        for B in self.bases:
            for s in B.symbols:
                if s not in syms:
                    syms.append(s)
        self.symbols = tuple(syms)
        self._proj = [tuple(self.symbols.index(s) for s in B.symbols) for B in self.bases]
        self.spec: List[tuple] = list(spec) if spec is not None else []
        self.indices = (np.zeros((0, len(self.bases)), dtype=np.int32) if indices is None
                        else np.asarray(indices, dtype=np.int32).reshape(-1, len(self.bases)))

    def __len__(self):
        return len(self.indices)

    def __mul__(self, other):
        ob = other.bases if isinstance(other, Product1pBasis) else (other,)
        return Product1pBasis(self.bases + tuple(ob))

    def component(self, kind):
        for B in self.bases:
            if B.kind == kind:
                return B
        return None

    def _sub(self, b: tuple, ib: int) -> tuple:
        return tuple(b[k] for k in self._proj[ib])

    def indexrange(self):
        """product_1pbasis.jl:283-304, including the m-range hack (-maxl..maxl)."""
        rg = {s: [] for s in self.symbols}
        for B in self.bases:
            for s, vals in B.indexrange().items():
                for v in vals:
                    if v not in rg[s]:
                        rg[s].append(v)
        ylm = self.component(COMP_YLM)
        if ylm is not None:
            lsym, msym = ylm.symbols
            maxl = max(rg[lsym])
            rg[msym] = list(range(-maxl, maxl + 1))
        return rg

    def isadmissible(self, b: tuple) -> bool:
        return all(B.isadmissible(self._sub(b, i)) for i, B in enumerate(self.bases))

    def degree(self, b: tuple, weight=None) -> float:
        """product_1pbasis.jl:328-331; b1pcomponent.jl:139-148; discrete1pbasis.jl:123."""
        tot = 0.0
        for i, B in enumerate(self.bases):
            if B.kind == COMP_CAT:
                continue
            d = B.degree(self._sub(b, i))
            tot += d if weight is None else weight[B.symbols[0]] * d
        return tot

    def set_spec(self, spec: Sequence[tuple]):
        """product_1pbasis.jl:308-315."""
        self._version = getattr(self, "_version", 0) + 1      # invalidates device handles built from the old tables (api._handle_of)
        self.spec = [tuple(b) for b in spec]
        self.indices = np.array(
            [[B.get_index(self._sub(b, i)) for i, B in enumerate(self.bases)] for b in self.spec],
            dtype=np.int32).reshape(-1, len(self.bases))
        return self

    def get_spec(self, i=None):
        return list(self.spec) if i is None else self.spec[i - 1]

    def sparsify(self, keep) -> np.ndarray:
        """Keep the functions in ``keep`` preserving order; returns old->new 1-based map, 0 = dropped
        (product_1pbasis.jl:345-374).  Components are *not* shrunk, as in the reference."""
        keep = set(keep)
        new_spec, new_inds = [], np.zeros(len(self.spec), dtype=np.int64)
        for ib, b in enumerate(self.spec):
            if b in keep:
                new_spec.append(b)
                new_inds[ib] = len(new_spec)
        self.set_spec(new_spec)
        return new_inds

    # -- decode helpers used to build the device descriptor ---------------------------------
    def sym_index(self, sym: str) -> int:
        return self.symbols.index(sym)

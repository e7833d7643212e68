"""Symmetric basis B = A2Bmap * AA: construction of the sparse coupling matrix.

Mirrors ``SymmetricBasis`` (src/symmbasis.jl:33-38), its constructor (:74-162) and the clean-up
``clean_pibasis!`` (:225-236).  The matrix is kept in the same CSC form as Julia's
``SparseMatrixCSC{PROP,Int}``: ``colptr`` (nAA+1), ``rowval`` (nnz), both 1-based, and ``nzval`` of
shape (nnz, ncomp) complex.  These arrays cross the C ABI unchanged.

The values are defined up to the SVD gauge: numpy/LAPACK here and
Julia/LAPACK there span the same space but need not produce identical numbers.  Everything downstream
(c~, energies, forces, B given this A2Bmap) is gauge-free.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from .onepbasis import Product1pBasis
from .pibasis import PIBasis, PIBasisSpec, _lm_of_spec
from .rotations3d import Rot3DCoeffs
from .selectors import DownsetBasisSelector
from .symmetrygroups import NoSym, O3


class SparseCSC:
    """Minimal CSC container with vector-valued entries (SparseMatrixCSC{PROP,Int})."""

    def __init__(self, m: int, n: int, colptr, rowval, nzval, ncomp: int = None):
        self.m, self.n = int(m), int(n)
        self.colptr = np.asarray(colptr, dtype=np.int32)
        self.rowval = np.asarray(rowval, dtype=np.int32)
        nz = np.asarray(nzval, dtype=np.complex128)
        if ncomp is None:
            ncomp = nz.shape[1] if nz.ndim == 2 else 1
        self.nzval = nz.reshape(len(self.rowval), ncomp)

    @property
    def shape(self):
        return (self.m, self.n)

    @property
    def nnz(self):
        return len(self.rowval)

    @property
    def ncomp(self):
        return self.nzval.shape[1]

    @classmethod
    def from_triplets(cls, I, J, V, m, n, ncomp):
        """sparse(I, J, V, m, n): duplicates are summed (symmbasis.jl:137-140, 156)."""
        I = np.asarray(I, dtype=np.int64)
        J = np.asarray(J, dtype=np.int64)
        V = np.asarray(V, dtype=np.complex128).reshape(len(I), ncomp)
        order = np.lexsort((I, J))
        I, J, V = I[order], J[order], V[order]
        rows: List[int] = []
        cols: List[int] = []
        vals: List[np.ndarray] = []
        for k in range(len(I)):
            if rows and rows[-1] == I[k] and cols[-1] == J[k]:
                vals[-1] = vals[-1] + V[k]
            else:
                rows.append(int(I[k]))
                cols.append(int(J[k]))
                vals.append(V[k].copy())
        colptr = np.ones(n + 1, dtype=np.int64)
        for c in cols:
            colptr[c] += 1
        colptr = np.concatenate(([1], 1 + np.cumsum(colptr[1:] - 1)))
        return cls(m, n, colptr, rows, np.array(vals).reshape(len(rows), ncomp), ncomp)

    def select_columns(self, keep0: np.ndarray):
        """A[:, keep] with keep 0-based sorted."""
        colptr = [1]
        rows, vals = [], []
        for c in keep0:
            a, b = self.colptr[c] - 1, self.colptr[c + 1] - 1
            rows.extend(self.rowval[a:b])
            vals.extend(self.nzval[a:b])
            colptr.append(len(rows) + 1)
        return SparseCSC(self.m, len(keep0), colptr, rows, np.array(vals).reshape(len(rows), self.ncomp), self.ncomp)

    def select_rows(self, keep0: np.ndarray):
        newrow = -np.ones(self.m, dtype=np.int64)
        newrow[np.asarray(keep0)] = np.arange(len(keep0))
        colptr = [1]
        rows, vals = [], []
        for c in range(self.n):
            for k in range(self.colptr[c] - 1, self.colptr[c + 1] - 1):
                r = newrow[self.rowval[k] - 1]
                if r >= 0:
                    rows.append(r + 1)
                    vals.append(self.nzval[k])
            colptr.append(len(rows) + 1)
        return SparseCSC(len(keep0), self.n, colptr, rows, np.array(vals).reshape(len(rows), self.ncomp), self.ncomp)

    def col_norms(self) -> np.ndarray:
        """sum(norm, A, dims=1)."""
        out = np.zeros(self.n)
        nrm = np.sqrt(np.sum(np.abs(self.nzval) ** 2, axis=1))
        for c in range(self.n):
            out[c] = nrm[self.colptr[c] - 1:self.colptr[c + 1] - 1].sum()
        return out

    def todense(self) -> np.ndarray:
        D = np.zeros((self.m, self.n, self.ncomp), dtype=np.complex128)
        for c in range(self.n):
            for k in range(self.colptr[c] - 1, self.colptr[c + 1] - 1):
                D[self.rowval[k] - 1, c, :] += self.nzval[k]
        return D


class SymmetricBasis:
    """symmbasis.jl:33-38.  ``real`` True means B = real(A2Bmap * AA)."""

    def __init__(self, phi, basis1p_or_pibasis, symgrp_or_Bsel=None, Bsel: Optional[DownsetBasisSelector] = None,
                 *, isreal: Optional[bool] = None, **kwargs):
        # accepted forms (symmbasis.jl:64-86):
        #   SymmetricBasis(phi, basis1p, Bsel); SymmetricBasis(phi, basis1p, symgrp, Bsel)
        #   SymmetricBasis(phi, pibasis);       SymmetricBasis(phi, symgrp, pibasis)
        a, b, c = basis1p_or_pibasis, symgrp_or_Bsel, Bsel
        if isinstance(a, (O3, NoSym)):          # (phi, symgrp, pibasis)
            symgrp, pibasis = a, b
            real = False if isreal is None else isreal
        elif isinstance(a, PIBasis):            # (phi, pibasis)
            symgrp, pibasis = O3(), a
            real = False if isreal is None else isreal
        else:                                   # 1p basis + selector
            symgrp, sel = (O3(), b) if c is None else (b, c)
            real = phi.isrealB if isreal is None else isreal
            pibasis = PIBasis(a, symgrp, sel, isreal=phi.isrealAA, property=phi, **kwargs)
        self.phi = phi
        self.symgrp = symgrp
        self.pibasis = pibasis
        self.real = bool(real)
        self.A2Bmap = self._build_A2B()
        self.clean_pibasis()

    @classmethod
    def from_parts(cls, phi, pibasis: PIBasis, A2Bmap: SparseCSC, symgrp, real: bool):
        """Assemble from finished tables (what `read_dict` does, symmbasis.jl:57-61)."""
        self = cls.__new__(cls)
        self.phi, self.pibasis, self.A2Bmap, self.symgrp, self.real = phi, pibasis, A2Bmap, symgrp, bool(real)
        return self

    def __len__(self):
        return self.A2Bmap.m

    # ---------------------------------------------------------------- construction
    def _build_A2B(self) -> SparseCSC:
        """symmbasis.jl:89-156."""
        pib, phi, grp = self.pibasis, self.phi, self.symgrp
        b1p: Product1pBasis = pib.basis1p
        spec = pib.spec
        nAA = len(spec)
        Aspec = b1p.get_spec()
        invA = {b: i + 1 for i, b in enumerate(Aspec)}
        invAA = {spec.get_spec(i): i for i in range(1, nAA + 1)}
        I, J, V = [], [], []
        idxB = 0
        if isinstance(grp, NoSym):
            for iAA in range(1, nAA + 1):
                idxB += 1
                I.append(idxB); J.append(iAA); V.append(np.ones(phi.ncomp))
            return SparseCSC.from_triplets(I, J, V, idxB, nAA, phi.ncomp)
        ls, ms = _lm_of_spec(b1p, grp)
        lk, mk = b1p.sym_index(grp.lsym), b1p.sym_index(grp.msym)
        rotc = Rot3DCoeffs(phi)
        for iAA in range(1, nAA + 1):
            vv = spec.get_spec(iAA)
            if not grp.is_refbasisfcn([ls[v - 1] for v in vv], [ms[v - 1] for v in vv]):
                continue
            if len(vv) == 0:                     # symmetrygroups.jl:100-102
                U = phi.coco_init0()
                cols = [()]
            else:
                bb = [Aspec[v - 1] for v in vv]
                ll = tuple(int(b[lk]) for b in bb)
                nn = tuple(tuple(x for k, x in enumerate(b) if k not in (lk, mk)) for b in bb)
                U, Ms = grp.rpe_basis(rotc, nn, ll)
                cols = []
                for mm in Ms:
                    bcol = []
                    for t, b in enumerate(bb):
                        bl = list(b)
                        bl[mk] = mm[t]
                        bcol.append(tuple(bl))
                    iAs = [invA.get(bt) for bt in bcol]
                    if any(v is None for v in iAs):
                        raise RuntimeError(f"bcol_ordered not in AA-spec: {bcol}")
                    cols.append(tuple(sorted(iAs, reverse=True)))   # _get_ordered (:174-177)
            for irow in range(U.shape[0]):
                idxB += 1
                for icol, key in enumerate(cols):
                    idxAA = invAA.get(key)
                    if idxAA is None:
                        raise RuntimeError(f"bcol_ordered not in AA-spec: {key}")
                    I.append(idxB); J.append(idxAA); V.append(U[irow, icol, :])
        return SparseCSC.from_triplets(I, J, V, idxB, nAA, phi.ncomp)

    def clean_pibasis(self, atol: float = 0.0):
        """symmbasis.jl:225-236."""
        self._version = getattr(self, "_version", 0) + 1
        nrm = self.A2Bmap.col_norms()
        Inz = np.nonzero(nrm > atol)[0]
        if len(Inz) < self.A2Bmap.n:
            self.pibasis.sparsify(Inz)
            self.A2Bmap = self.A2Bmap.select_columns(Inz)
        self.pibasis.clean_1pbasis()
        return self

    def sparsify(self, *, keep=None, delete=None):
        """symmbasis.jl:204-216 (1-based indices)."""
        if (keep is None) == (delete is None):
            raise ValueError("sparsify!: must provide either del or keep kwarg but not both")
        if delete is not None:
            dele = set(int(d) for d in delete)
            keep = [i for i in range(1, len(self) + 1) if i not in dele]
        keep0 = np.asarray(sorted(int(k) - 1 for k in keep))
        self._version = getattr(self, "_version", 0) + 1
        self.A2Bmap = self.A2Bmap.select_rows(keep0)
        return self.clean_pibasis(atol=0.0)

"""Flatten a SymmetricBasis (+ coefficients) into the ``aceb200_desc`` the C ABI takes.

This is the marshalling the Julia shim performs from the live ACE.jl objects:
``basis1p.bases`` / ``basis1p.indices`` (src/product_1pbasis.jl:5-8), ``pibasis.spec.orders`` /
``iAA2iA`` (src/pibasis.jl:10-13), ``A2Bmap`` (src/symmbasis.jl:33-38) and ``model.c``
(src/linearmodel.jl:36-40).  Index tables stay 1-based; ``iAA2iA`` is passed column-major like a
Julia ``Matrix``.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from ._lib import DescHolder
from .onepbasis import COMP_CAT, COMP_RN, COMP_YLM, Product1pBasis


def coeffs_array(c, nB: int):
    """Vector{T} or Vector{SVector{N,T}} -> ([nB][nprop] float64, nprop)."""
    if c is None:
        return None, 1
    c = np.asarray(c, dtype=np.float64)
    if c.ndim == 1:
        c = c.reshape(nB, 1)
    if c.shape[0] != nB:
        raise ValueError(f"coefficients: expected {nB} rows, got {c.shape}")
    return np.ascontiguousarray(c), c.shape[1]


def basis_descriptor(basis, c=None) -> DescHolder:
    pib = basis.pibasis
    b1p: Product1pBasis = pib.basis1p
    Rn = b1p.component(COMP_RN)
    Ylm = b1p.component(COMP_YLM)
    Cat = b1p.component(COMP_CAT)
    if Rn is None or Ylm is None:
        raise ValueError("the B200 path needs an Rn1pBasis and a Ylm1pBasis component")
    if len(b1p.bases) > 3 or sorted(B.kind for B in b1p.bases) not in ([COMP_RN, COMP_YLM], [COMP_RN, COMP_YLM, COMP_CAT]):
        raise ValueError("unsupported one-particle basis: supported are Rn*Ylm and Categorical*Rn*Ylm (any order)")
    A2B = basis.A2Bmap
    cc, nprop = coeffs_array(c, A2B.m)
    nz = np.ascontiguousarray(A2B.nzval.astype(np.complex128)).view(np.float64)
    return DescHolder(
        n_rad=len(Rn.R), pl=Rn.R.pl, pr=Rn.R.pr, tl=Rn.R.tl, tr=Rn.R.tr,
        rad_A=Rn.R.A, rad_B=Rn.R.B, rad_C=Rn.R.C,
        trans_kind=Rn.trans.kind, trans_par=Rn.trans.c_params(),
        maxL=Ylm.L, n_cat=(len(Cat) if Cat is not None else 0),
        n_comp=len(b1p.bases), comp_kind=[B.kind for B in b1p.bases],
        nA=len(b1p), indices=b1p.indices,
        nAA=len(pib), maxord=pib.maxcorrorder, orders=pib.spec.orders,
        iAA2iA=np.asfortranarray(pib.spec.iAA2iA).ravel(order="F"),
        pireal=int(pib.real), symreal=int(basis.real),
        nB=A2B.m, ncomp=A2B.ncomp, nnz=A2B.nnz,
        colptr=A2B.colptr, rowval=A2B.rowval, nzval=nz,
        nprop=nprop, c=cc,
    )

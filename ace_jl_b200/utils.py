"""Convenience constructors and the synthetic-input generators.

Mirrors ``ACE.Utils.Rn_basis`` / ``RnYlm_1pbasis`` (src/utils/utils.jl:28-64) and the random
generators ``rand_radial`` / ``rand_vec3`` (src/utils/random.jl:22-25), ``rand_sphere`` /
``rand_rot`` / ``rand_refl`` (src/utils/auxiliary.jl:6-25).  The random stream is numpy's Philox
(counter-based, reproducible anywhere); Julia's stream is not reproduced and no reference test
depends on it.
"""
from __future__ import annotations

import numpy as np

from .onepbasis import Product1pBasis, Rn1pBasis, Ylm1pBasis
from .orthpolys import discrete_jacobi
from .selectors import init1pspec
from .transforms import polytransform


def Rn_basis(*, r0=1.0, trans=None, maxdeg=6, rcut=2.5, rin=None, pcut=2, pin=0,
             varsym="rr", nsym="n", label=None) -> Rn1pBasis:
    """utils.jl:28-47."""
    trans = polytransform(2, r0) if trans is None else trans
    rin = 0.5 * r0 if rin is None else rin
    J = discrete_jacobi(maxdeg, pcut=pcut, xcut=rcut, pin=pin, xin=rin, trans=trans)
    return Rn1pBasis(J, trans, varsym=varsym, nsym=nsym, label=label)


def RnYlm_1pbasis(*, maxdeg=6, maxL=None, varsym="rr", idxsyms=("n", "l", "m"), Bsel=None,
                  **kwargs) -> Product1pBasis:
    """utils.jl:54-64."""
    maxL = maxdeg if maxL is None else maxL
    Rn = Rn_basis(maxdeg=maxdeg, varsym=varsym, nsym=idxsyms[0], **kwargs)
    Ylm = Ylm1pBasis(maxL, varsym=varsym, lsym=idxsyms[1], msym=idxsyms[2])
    B1p = Product1pBasis((Rn, Ylm))
    if Bsel is not None:
        init1pspec(B1p, Bsel)
    return B1p


def philox(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(seed))


def rand_sphere(rng: np.random.Generator, n: int = 1) -> np.ndarray:
    """auxiliary.jl:6-9: normalised Gaussian vectors, shape (n, 3)."""
    g = rng.standard_normal((n, 3))
    return g / np.linalg.norm(g, axis=1, keepdims=True)


def rand_radial(rng: np.random.Generator, Rn: Rn1pBasis, n: int = 1) -> np.ndarray:
    """random.jl:22-23: uniform in *radius* between rin and rcut."""
    return Rn.meta["rin"] + rng.random(n) * (Rn.meta["rcut"] - Rn.meta["rin"])


def rand_vec3(rng: np.random.Generator, Rn: Rn1pBasis, n: int = 1) -> np.ndarray:
    """random.jl:25, shape (n, 3)."""
    return rand_radial(rng, Rn, n)[:, None] * rand_sphere(rng, n)


def rand_rot(rng: np.random.Generator) -> np.ndarray:
    """auxiliary.jl:15: exp(K - K') of a random matrix."""
    from scipy.linalg import expm
    K = rng.random((3, 3)) - 0.5
    return expm(K - K.T)


def rand_refl(rng: np.random.Generator) -> int:
    return int(rng.choice([-1, 1]))


def rand_envs(rng: np.random.Generator, Rn: Rn1pBasis, nenv: int, J, nspecies: int = 0):
    """A ragged batch of random environments.

    ``J`` is an int (every environment has J neighbours) or an array of per-environment counts.
    Returns ``(R, offsets, species)`` with R of shape (sum J, 3), int64 offsets (nenv+1) and int32
    1-based species codes (or None).
    """
    counts = np.full(nenv, int(J), dtype=np.int64) if np.isscalar(J) else np.asarray(J, dtype=np.int64)
    offsets = np.concatenate(([0], np.cumsum(counts))).astype(np.int64)
    tot = int(offsets[-1])
    R = rand_vec3(rng, Rn, tot) if tot > 0 else np.zeros((0, 3))
    species = rng.integers(1, nspecies + 1, size=tot).astype(np.int32) if nspecies > 0 else None
    return np.ascontiguousarray(R), offsets, species


def fcc_structure(rng: np.random.Generator, ncell: int, d_nn: float = 1.3, jitter: float = 0.02, rcut: float = 2.5):
    """A jittered periodic FCC crystal with its full neighbour list, built analytically (vectorised, so that the
    10^6-atom benchmark structure takes seconds): 4 ncell^3 atoms, nearest-neighbour distance ``d_nn``; with the
    defaults every atom has the 42 neighbours of the first three shells inside ``rcut`` (shell radii 1.30, 1.84,
    2.25 | 2.60) and the jitter cannot move a pair across the cutoff.  Returns X, cell, first, nbr, image."""
    a = d_nn * np.sqrt(2.0)
    h = 0.5 * a                                          # work in half lattice constants: integer coordinates
    margin = 2.0 * jitter * np.sqrt(3.0)
    rng_i = int(np.ceil(rcut / h)) + 1
    g = np.arange(-rng_i, rng_i + 1)
    V = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    V = V[(V.sum(axis=1) % 2 == 0) & (np.abs(V).sum(axis=1) > 0)]
    r = np.linalg.norm(V, axis=1) * h
    if np.any(np.abs(r - rcut) < margin):
        raise ValueError("a neighbour shell is too close to the cutoff for this jitter")
    V = V[r < rcut]
    V = V[np.lexsort((V[:, 2], V[:, 1], V[:, 0]))]
    hb = np.array([[0, 0, 0], [0, 1, 1], [1, 0, 1], [1, 1, 0]])
    bmap = {tuple(b): k for k, b in enumerate(hb)}
    c = np.stack(np.meshgrid(*[np.arange(ncell)] * 3, indexing="ij"), -1).reshape(-1, 3)
    H = (2 * c[:, None, :] + hb[None, :, :]).reshape(-1, 3)              # [natoms, 3], atom index = cell * 4 + b
    natoms = len(H)
    Hn = H[:, None, :] + V[None, :, :]                                   # [natoms, nv, 3]
    par = Hn % 2
    bn = np.empty(par.shape[:2], dtype=np.int64)
    for b, k in bmap.items():
        bn[(par == np.array(b)).all(axis=2)] = k
    cn = (Hn - hb[bn]) // 2
    img = np.floor_divide(cn, ncell)
    cn = cn - img * ncell
    nbr = ((cn[..., 0] * ncell + cn[..., 1]) * ncell + cn[..., 2]) * 4 + bn
    X = H * h + (rng.random(H.shape) - 0.5) * 2.0 * jitter
    cell = np.eye(3) * (ncell * a)
    first = np.arange(natoms + 1, dtype=np.int64) * len(V)
    return X, cell, first, nbr.reshape(-1).astype(np.int32), img.reshape(-1, 3).astype(np.int8)

"""Caller side of the evaluation path: atomic structure + neighbour list.

In the reference ecosystem the loop over centres lives outside ACE.jl, in JuLIP / ACEatoms.jl
(``energy(V, at)``, ``forces(V, at)``, ``virial(V, at)``): for every centre i it builds the environment
``Rs = {x_j + S_ij - x_i}``, calls ``evaluate`` / ``evaluate_d`` on it and scatters ``frc[j] -= dV_j``,
``frc[i] += dV_j``, ``vir -= dV_j (x) R_j``.  ``B200Structure`` hands the whole structure to the library
(``aceb200_structure_energy_forces``, include/aceb200.h), which builds the environments and assembles the
forces on the device: the pair gradients (24 B per pair) never cross PCIe.

``neighbourlist`` is a plain cell-list builder (host, numpy) for tests, examples and the benchmark; a caller
that already has a neighbour list (JuLIP's ``neighbourlist(at, rcut)``: i, j, S) passes it as is after sorting
by centre.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib as L

try:  # torch only provides device tensors here
    import torch
except Exception:  # pragma: no cover
    torch = None


def _is_torch(x) -> bool:
    return torch is not None and isinstance(x, torch.Tensor)


class B200Structure:
    """X (natoms, 3); first (natoms+1) pair offsets by centre; nbr (npairs) 0-based neighbour indices;
    optional image (npairs, 3) int8 lattice shifts with cell (3, 3) (rows = lattice vectors, as JuLIP's
    neighbour list returns them: i, j, S), species (natoms) 1-based categories, rev (npairs) int32 reverse-pair
    table (-1: none).  numpy arrays = host, CUDA torch tensors = device (cell always numpy)."""

    def __init__(self, X, first, nbr, image=None, cell=None, species=None, rev=None, packed=False):
        """``packed=True``: ``nbr`` already holds packed words (``pack_neighbours``) and ``image`` is None."""
        self.packed = bool(packed)
        if self.packed and image is not None:
            raise ValueError("packed neighbour words carry the image shift: pass image=None")
        self.device = _is_torch(X)
        if self.device:
            cv = lambda a, dt: None if a is None else a.contiguous().to(dt)   # noqa: E731
            self.X = cv(X, torch.float64).reshape(-1, 3)
            self.first, self.nbr = cv(first, torch.int64), cv(nbr, torch.int32)
            self.image = None if image is None else cv(image, torch.int8).reshape(-1, 3)
            self.species, self.rev = cv(species, torch.int32), cv(rev, torch.int32)
        else:
            cv = lambda a, dt: None if a is None else np.ascontiguousarray(a, dtype=dt)   # noqa: E731
            self.X = cv(X, np.float64).reshape(-1, 3)
            self.first, self.nbr = cv(first, np.int64), cv(nbr, np.int32)
            self.image = None if image is None else cv(image, np.int8).reshape(-1, 3)
            self.species, self.rev = cv(species, np.int32), cv(rev, np.int32)
        if (self.image is not None or self.packed) and cell is None:
            raise ValueError("periodic images need a cell")
        self.cell = np.zeros((3, 3)) if cell is None else np.ascontiguousarray(cell, dtype=np.float64).reshape(3, 3)
        self.natoms, self.npairs = int(self.X.shape[0]), int(self.nbr.shape[0])
        if int(self.first.shape[0]) != self.natoms + 1:
            raise ValueError("first must have natoms + 1 entries")

    def _ptr(self, x):
        if x is None:
            return None
        return x.data_ptr() if self.device else x.ctypes.data

    def c_struct(self) -> L.Structure:
        s = L.Structure()
        s.natoms, s.npairs = self.natoms, self.npairs
        s.X = C.cast(C.c_void_p(self._ptr(self.X)), L.c_double_p)
        s.first = C.cast(C.c_void_p(self._ptr(self.first)), L.c_int64_p)
        s.nbr = C.cast(C.c_void_p(self._ptr(self.nbr)), L.c_int32_p)
        s.image = C.cast(C.c_void_p(self._ptr(self.image)), C.POINTER(C.c_int8))
        s.species = C.cast(C.c_void_p(self._ptr(self.species)), L.c_int32_p)
        s.rev = C.cast(C.c_void_p(self._ptr(self.rev)), L.c_int32_p)
        for i, v in enumerate(self.cell.ravel()):
            s.cell[i] = float(v)
        s.space = L.DEVICE if self.device else L.HOST
        s.flags = L.NBR_PACKED if self.packed else 0
        return s

    def empty(self, shape):
        if self.device:
            return torch.empty(shape, dtype=torch.float64, device=self.X.device)
        return np.empty(shape, dtype=np.float64)

    # ---- what the library does on the device, on the host: for callers that want the per-environment batch
    # (R, offsets, species) of this structure
    def environments(self):
        if self.device:
            raise ValueError("environments() is a host-side helper")
        centre = np.repeat(np.arange(self.natoms), np.diff(self.first))
        nbr, image = self.nbr, self.image
        if self.packed:
            w = self.nbr.view(np.uint32).astype(np.int64)
            nbr = (w & 0x03ffffff).astype(np.int32)
            image = np.stack([((w >> (26 + 2 * k)) & 3) - 1 for k in range(3)], axis=1).astype(np.int8)
        R = self.X[nbr] - self.X[centre]
        if image is not None:
            R = R + image.astype(np.float64) @ self.cell
        sp = None if self.species is None else self.species[nbr]
        return R, self.first.copy(), sp, centre


def pack_neighbours(nbr, image) -> np.ndarray:
    """int32 words ``j | (S0 + 1) << 26 | (S1 + 1) << 28 | (S2 + 1) << 30`` (ACEB200_NBR_PACKED): 4 instead of 7 bytes per
    pair cross PCIe.  Needs j < 2^26 and S in {-1, 0, 1} (any cell wider than the cutoff)."""
    nbr = np.asarray(nbr, dtype=np.int64)
    S = np.asarray(image, dtype=np.int64).reshape(-1, 3)
    if nbr.size and (nbr.max() >= (1 << 26) or nbr.min() < 0 or np.abs(S).max() > 1):
        raise ValueError("pack_neighbours: needs 0 <= j < 2^26 and image shifts in {-1, 0, 1}")
    w = nbr | ((S[:, 0] + 1) << 26) | ((S[:, 1] + 1) << 28) | ((S[:, 2] + 1) << 30)
    return w.astype(np.uint32).view(np.int32)


def reverse_pairs(first, nbr, image=None) -> np.ndarray:
    """rev[p] = index of the pair (centre nbr[p], neighbour centre(p), image -S_p), or -1."""
    first, nbr = np.asarray(first, dtype=np.int64), np.asarray(nbr, dtype=np.int64)
    natoms, npairs = len(first) - 1, len(nbr)
    centre = np.repeat(np.arange(natoms, dtype=np.int64), np.diff(first))
    S = np.zeros((npairs, 3), dtype=np.int64) if image is None else np.asarray(image, dtype=np.int64)
    key = {}
    for p in range(npairs):
        key.setdefault((int(centre[p]), int(nbr[p]), int(S[p, 0]), int(S[p, 1]), int(S[p, 2])), p)
    rev = np.full(npairs, -1, dtype=np.int32)
    for p in range(npairs):
        rev[p] = key.get((int(nbr[p]), int(centre[p]), -int(S[p, 0]), -int(S[p, 1]), -int(S[p, 2])), -1)
    return rev


def neighbourlist(X, rcut: float, cell: Optional[np.ndarray] = None, pbc=(False, False, False), with_rev: bool = True):
    """All pairs (i, j, S) with 0 < |x_j + S . cell - x_i| < rcut, sorted by centre i (then by j, image).

    ``cell`` rows are the lattice vectors; periodic images are enumerated for the directions in ``pbc``.  Brute
    force over images with a distance matrix per image: O(N^2), meant for the small structures of tests.
    Returns ``first, nbr, image, rev`` (image None when nothing is periodic).
    """
    X = np.asarray(X, dtype=np.float64).reshape(-1, 3)
    n = len(X)
    images = [(0, 0, 0)]
    periodic = cell is not None and any(pbc)
    if periodic:
        cell = np.asarray(cell, dtype=np.float64).reshape(3, 3)
        # number of images needed per direction: rcut over the height of the cell along that direction
        vol = abs(np.linalg.det(cell))
        reps = []
        for d in range(3):
            a, b = cell[(d + 1) % 3], cell[(d + 2) % 3]
            height = vol / np.linalg.norm(np.cross(a, b))
            reps.append(int(np.ceil(rcut / height)) if pbc[d] else 0)
        images = [(i, j, k) for i in range(-reps[0], reps[0] + 1) for j in range(-reps[1], reps[1] + 1)
                  for k in range(-reps[2], reps[2] + 1)]
    I, J, S = [], [], []
    for img in images:
        s = np.zeros(3) if not periodic else np.asarray(img, dtype=np.float64) @ cell
        D = X[None, :, :] + s[None, None, :] - X[:, None, :]           # D[i, j] = x_j + s - x_i
        r = np.linalg.norm(D, axis=2)
        ii, jj = np.nonzero((r < rcut) & (r > 0.0))
        I.append(ii); J.append(jj); S.append(np.broadcast_to(np.asarray(img, dtype=np.int64), (len(ii), 3)))
    I, J, S = np.concatenate(I), np.concatenate(J), np.concatenate(S, axis=0)
    order = np.lexsort((S[:, 2], S[:, 1], S[:, 0], J, I))
    I, J, S = I[order], J[order], np.ascontiguousarray(S[order])
    first = np.zeros(n + 1, dtype=np.int64)
    np.add.at(first, I + 1, 1)
    first = np.cumsum(first)
    image = S.astype(np.int8) if periodic else None
    rev = reverse_pairs(first, J, image) if with_rev else None
    return first, J.astype(np.int32), image, rev

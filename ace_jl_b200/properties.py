"""Output properties: selection rules, coupling-coefficient seeds, component layout.

Mirrors the parts of src/properties.jl that define *what the hot path outputs*: ``Invariant``
(:77-152), ``EuclideanVector`` (:165-248), ``EuclideanMatrix`` / ``SymmetricEuclideanMatrix``
(:252-343).  A coupling coefficient ("coco") is stored as a complex vector of ``ncomp`` components:
1 (Invariant), 3 (EuclideanVector), 9 (matrices, column-major like Julia's SMatrix).

The Euclidean seeds are *derived* here, not copied from the reference's tables
(src/eucl/*.jl): with U[:,m] the Cartesian components of the spherical unit vectors,
    crmatrices[(1,m,mu,i)][k]      = conj(U[k,m]) U[i,mu] / 3
    mrmatrices[(l,m,mu,i,j)][a,b]  = 1/(2l+1) sum_{al+be=m, al'+be'=mu} C(1al',1be'|l mu) C(1al,1be|l m)
                                      conj(U[a,al] U[b,be]) U[i,al'] U[j,be']
which are the Haar integrals of D^l_{mu m}(Q) (Q e_i) and D^l_{mu m}(Q) (Q E_ij Q^T).
tests/test_coupling.py checks them against the reference tables when /root/reference is present.
"""
from __future__ import annotations

import math

import numpy as np

_S = 1.0 / math.sqrt(2.0)
# U[k, m+1]: Cartesian components of e_{-1}, e_0, e_{+1}
_U = np.zeros((3, 3), dtype=np.complex128)
_U[:, 0] = [-_S, 1j * _S, 0.0]
_U[:, 1] = [0.0, 0.0, 1.0]
_U[:, 2] = [_S, 1j * _S, 0.0]


def crmatrix(m: int, mu: int, i: int) -> np.ndarray:
    """Derived equivalent of crmatrices[(l=1, m, mu, i)] (i is 1-based); a complex 3-vector."""
    return np.conj(_U[:, m + 1]) * _U[i - 1, mu + 1] / 3.0


def mrmatrix(l: int, m: int, mu: int, i: int, j: int) -> np.ndarray:
    """Derived equivalent of mrmatrices[(l, m, mu, i, j)]; complex 3x3, [a, b]."""
    from .rotations3d import clebschgordan
    out = np.zeros((3, 3), dtype=np.complex128)
    for al in (-1, 0, 1):
        be = m - al
        if abs(be) > 1:
            continue
        for alp in (-1, 0, 1):
            bep = mu - alp
            if abs(bep) > 1:
                continue
            c = clebschgordan(1, alp, 1, bep, l, mu) * clebschgordan(1, al, 1, be, l, m) / (2 * l + 1)
            if c != 0.0:
                out += c * np.conj(np.outer(_U[:, al + 1], _U[:, be + 1])) * _U[i - 1, alp + 1] * _U[j - 1, bep + 1]
    return out


class AbstractProperty:
    name: str
    ncomp: int        # complex components per coupling coefficient
    numcc: int        # number of coefficient vectors a seed returns (properties.jl: coco_init)
    isrealB: bool
    isrealAA: bool

    def coco_zeros(self):
        return np.zeros((self.numcc, self.ncomp), dtype=np.complex128)


class Invariant(AbstractProperty):
    """properties.jl:77-152."""
    name = "Invariant"
    ncomp, numcc = 1, 1
    isrealB, isrealAA = True, True

    def filter(self, ls, ms) -> bool:            # :115-125
        if len(ls) <= 1:
            return True
        return sum(ls) % 2 == 0 and sum(ms) == 0

    def coco_init0(self):                         # :139, order-0 function
        return np.ones((1, 1, 1), dtype=np.complex128)

    def coco_init(self, l, m, mu):                # :141-142
        out = self.coco_zeros()
        if l == 0 and m == 0 and mu == 0:
            out[0, 0] = 1.0
        return out

    def coco_filter(self, ll, mm, kk=None) -> bool:  # :146-150
        ok = sum(ll) % 2 == 0 and sum(mm) == 0
        return ok and (kk is None or sum(kk) == 0)

    @staticmethod
    def coco_dot(u1, u2):                         # :152, no conjugation
        return np.einsum("aic,bic->ab", u1, u2)


class EuclideanVector(AbstractProperty):
    """properties.jl:165-248."""
    name = "EuclideanVector"
    ncomp, numcc = 3, 3
    isrealB, isrealAA = True, False

    def filter(self, ls, ms) -> bool:            # :189-203
        if len(ls) == 0:
            return False
        if len(ls) == 1:
            return True
        return sum(ls) % 2 == 1 and abs(sum(ms)) <= 1

    def coco_init0(self):
        raise ValueError("EuclideanVector has no order-0 basis function")

    def coco_init(self, l, m, mu):                # :219-222
        out = self.coco_zeros()
        if l == 1 and abs(m) <= 1 and abs(mu) <= 1:
            for i in range(3):
                out[i, :] = np.conj(crmatrix(-m, -mu, i + 1))
        return out

    def coco_filter(self, ll, mm, kk=None) -> bool:  # :239-246
        ok = sum(ll) % 2 == 1 and abs(sum(mm)) <= 1
        return ok and (kk is None or abs(sum(kk)) <= 1)

    @staticmethod
    def coco_dot(u1, u2):                         # :248, LinearAlgebra.dot conjugates its first argument
        return np.einsum("aic,bic->ab", np.conj(u1), u2)


class EuclideanMatrix(AbstractProperty):
    """properties.jl:252-313."""
    name = "EuclideanMatrix"
    ncomp, numcc = 9, 9
    isrealB, isrealAA = True, False
    _ls = (0, 1, 2)

    def filter(self, ls, ms) -> bool:            # :264-277
        if len(ls) == 0:
            return False
        if len(ls) == 1:
            return True
        return sum(ls) % 2 == 0 and abs(sum(ms)) <= 2

    def coco_init0(self):
        raise ValueError("EuclideanMatrix has no order-0 basis function")

    def _seed_ok(self, l, m, mu):
        return l <= 2 and abs(m) <= l and abs(mu) <= l

    def coco_init(self, l, m, mu):                # :310-313, `for i=1:3 for j=1:3` (j fastest)
        out = self.coco_zeros()
        if self._seed_ok(l, m, mu):
            for i in range(3):
                for j in range(3):
                    M = np.conj(mrmatrix(l, -m, -mu, i + 1, j + 1))
                    out[3 * i + j, :] = M.reshape(-1, order="F")  # SMatrix memory order
        return out

    def coco_filter(self, ll, mm, kk=None) -> bool:  # :285-289
        ok = sum(ll) % 2 == 0 and abs(sum(mm)) <= 2
        return ok and (kk is None or abs(sum(kk)) <= 2)

    @staticmethod
    def coco_dot(u1, u2):
        """sum(transpose(conj.(u1.val)) * u2.val) (:290): the inner product of the row sums."""
        s1 = u1.reshape(u1.shape[0], u1.shape[1], 3, 3).sum(axis=2)  # flat = a + 3 b -> [b, a]; sum over b
        s2 = u2.reshape(u2.shape[0], u2.shape[1], 3, 3).sum(axis=2)
        return np.einsum("aik,bik->ab", np.conj(s1), s2)


class SymmetricEuclideanMatrix(EuclideanMatrix):
    """properties.jl:333-343: seeds only for l = 0 and l = 2."""
    name = "SymmetricEuclideanMatrix"

    def _seed_ok(self, l, m, mu):
        return (l == 2 and abs(m) <= 2 and abs(mu) <= 2) or (l == 0 and m == 0 and mu == 0)

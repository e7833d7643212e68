from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path() -> str:
    return os.path.join(_HERE, "libace_oracle.so")


def build(force: bool = False) -> str:
    """gcc -O3 -fopenmp the oracle; a no-op if the .so is newer than its sources."""
    so, src = lib_path(), os.path.join(_HERE, "ace_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "aceb200.h")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libace_oracle.so"])
    return so


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


class Oracle:
    """Evaluate a model descriptor on the CPU, following the reference line by line.

    ``holder`` is an ``ace_jl_b200._lib.DescHolder`` (the same descriptor the CUDA library gets), so
    both sides see identical tables.
    """

    def __init__(self, holder, threads: int | None = None):
        from ace_jl_b200 import _lib as L
        self.L = L
        self.holder = holder
        self.d = holder.desc
        if not os.path.exists(lib_path()):
            build()
        self.lib = C.CDLL(lib_path())
        if threads is not None:
            self.lib.oracle_set_threads(int(threads))
        self.lib.oracle_transform.restype = C.c_double
        self.lib.oracle_transform.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_double]
        self.lib.oracle_transform_d.restype = C.c_double
        self.lib.oracle_transform_d.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_double]
        self.cs = 1 if self.d.symreal else 2
        self.ca = 1 if self.d.pireal else 2

    # ---- helpers -----------------------------------------------------------------------
    def _batch(self, R, offsets, species=None):
        R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, 3)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        sp = None if species is None else np.ascontiguousarray(species, dtype=np.int32)
        b = self.L.make_batch(len(offsets) - 1, offsets.ctypes.data, R.ctypes.data,
                              sp.ctypes.data if sp is not None else 0, self.L.HOST)
        return b, (R, offsets, sp)

    def _call(self, name, *args):
        rc = getattr(self.lib, name)(*args)
        if rc != 0:
            raise self.L.AceB200Error(rc, f"oracle {name}")

    def _cplx(self, a, real):
        return a if real else a.view(np.complex128).reshape(a.shape[:-1])

    def num_threads(self) -> int:
        return int(self.lib.oracle_num_threads())

    def set_threads(self, n: int):
        """Explicit thread count (torchrun exports OMP_NUM_THREADS=1, which would hide the host's cores)."""
        self.lib.oracle_set_threads(int(n))

    # ---- components --------------------------------------------------------------------
    def transform(self, r):
        q = (C.c_double * 4)(*self.d.trans_par)
        return self.lib.oracle_transform(self.d.trans_kind, q, float(r))

    def transform_d(self, r):
        q = (C.c_double * 4)(*self.d.trans_par)
        return self.lib.oracle_transform_d(self.d.trans_kind, q, float(r))

    def rn(self, rr):
        rr = np.ascontiguousarray(rr, dtype=np.float64)
        P = np.zeros(self.d.n_rad)
        self.lib.oracle_rn(C.byref(self.d), _dp(rr), _dp(P))
        return P

    def rn_ed(self, rr):
        rr = np.ascontiguousarray(rr, dtype=np.float64)
        P, dP = np.zeros(self.d.n_rad), np.zeros((self.d.n_rad, 3))
        self.lib.oracle_rn_ed(C.byref(self.d), _dp(rr), _dp(P), _dp(dP))
        return P, dP

    def ylm(self, L, rr):
        rr = np.ascontiguousarray(rr, dtype=np.float64)
        Y = np.zeros((L + 1) ** 2, dtype=np.complex128)
        self.lib.oracle_ylm(int(L), _dp(rr), Y.ctypes.data_as(C.POINTER(C.c_double)))
        return Y

    def ylm_ed(self, L, rr):
        rr = np.ascontiguousarray(rr, dtype=np.float64)
        Y = np.zeros((L + 1) ** 2, dtype=np.complex128)
        dY = np.zeros(((L + 1) ** 2, 3), dtype=np.complex128)
        self.lib.oracle_ylm_ed(int(L), _dp(rr), Y.ctypes.data_as(C.POINTER(C.c_double)),
                               dY.ctypes.data_as(C.POINTER(C.c_double)))
        return Y, dY

    # ---- bases -------------------------------------------------------------------------
    def eval_A(self, R, offsets, species=None):
        b, keep = self._batch(R, offsets, species)
        out = np.zeros((b.nenv, self.d.nA), dtype=np.complex128)
        self._call("oracle_eval_A", C.byref(self.d), C.byref(b), out.ctypes.data_as(C.c_void_p))
        return out

    def eval_AA(self, R, offsets, species=None):
        b, keep = self._batch(R, offsets, species)
        out = np.zeros((b.nenv, self.d.nAA), dtype=np.float64 if self.d.pireal else np.complex128)
        self._call("oracle_eval_AA", C.byref(self.d), C.byref(b), out.ctypes.data_as(C.c_void_p))
        return out

    def eval_B(self, R, offsets, species=None):
        b, keep = self._batch(R, offsets, species)
        out = np.zeros((b.nenv, self.d.nB, self.d.ncomp), dtype=np.float64 if self.d.symreal else np.complex128)
        self._call("oracle_eval_B", C.byref(self.d), C.byref(b), out.ctypes.data_as(C.c_void_p))
        return out

    def eval_dA(self, R, offsets, species=None):
        b, keep = self._batch(R, offsets, species)
        nj = int(keep[1][-1])
        A = np.zeros((b.nenv, self.d.nA), dtype=np.complex128)
        dA = np.zeros((nj, self.d.nA, 3), dtype=np.complex128)
        self._call("oracle_eval_dA", C.byref(self.d), C.byref(b), A.ctypes.data_as(C.c_void_p), dA.ctypes.data_as(C.c_void_p))
        return A, dA

    def eval_dAA(self, R, offsets, species=None):
        b, keep = self._batch(R, offsets, species)
        nj = int(keep[1][-1])
        dt = np.float64 if self.d.pireal else np.complex128
        AA = np.zeros((b.nenv, self.d.nAA), dtype=dt)
        dAA = np.zeros((nj, self.d.nAA, 3), dtype=dt)
        self._call("oracle_eval_dAA", C.byref(self.d), C.byref(b), AA.ctypes.data_as(C.c_void_p), dAA.ctypes.data_as(C.c_void_p))
        return AA, dAA

    def eval_dB(self, R, offsets, species=None):
        b, keep = self._batch(R, offsets, species)
        nj = int(keep[1][-1])
        dt = np.float64 if self.d.symreal else np.complex128
        B = np.zeros((b.nenv, self.d.nB, self.d.ncomp), dtype=dt)
        dB = np.zeros((nj, self.d.nB, 3, self.d.ncomp), dtype=dt)
        self._call("oracle_eval_dB", C.byref(self.d), C.byref(b), B.ctypes.data_as(C.c_void_p), dB.ctypes.data_as(C.c_void_p))
        return B, dB

    # ---- model -------------------------------------------------------------------------
    def eff_coeffs(self, c=None):
        if c is None:
            cptr = self.d.c
        else:
            c = np.ascontiguousarray(c, dtype=np.float64)
            cptr = _dp(c)
        ct = np.zeros((self.d.nAA, self.d.nprop, self.d.ncomp), dtype=np.complex128)
        self._call("oracle_eff_coeffs", C.byref(self.d), cptr, ct.ctypes.data_as(C.c_void_p))
        return ct

    def energy(self, R, offsets, species=None, ctilde=None):
        b, keep = self._batch(R, offsets, species)
        ct = self.eff_coeffs() if ctilde is None else np.ascontiguousarray(ctilde, dtype=np.complex128)
        dt = np.float64 if self.d.symreal else np.complex128
        E = np.zeros((b.nenv, self.d.nprop, self.d.ncomp), dtype=dt)
        self._call("oracle_energy", C.byref(self.d), C.byref(b), ct.ctypes.data_as(C.c_void_p), E.ctypes.data_as(C.c_void_p))
        return E

    def energy_forces(self, R, offsets, species=None, ctilde=None):
        b, keep = self._batch(R, offsets, species)
        nj = int(keep[1][-1])
        ct = self.eff_coeffs() if ctilde is None else np.ascontiguousarray(ctilde, dtype=np.complex128)
        dt = np.float64 if self.d.symreal else np.complex128
        E = np.zeros((b.nenv, self.d.nprop, self.d.ncomp), dtype=dt)
        G = np.zeros((nj, self.d.nprop, 3, self.d.ncomp), dtype=dt)
        self._call("oracle_energy_forces", C.byref(self.d), C.byref(b), ct.ctypes.data_as(C.c_void_p),
                   E.ctypes.data_as(C.c_void_p), G.ctypes.data_as(C.c_void_p))
        return E, G

    def adjoint_eval_d(self, R, offsets, w, species=None):
        b, keep = self._batch(R, offsets, species)
        w = np.ascontiguousarray(w, dtype=np.float64)
        out = np.zeros((b.nenv, self.d.nB, self.d.ncomp), dtype=np.complex128)
        self._call("oracle_adjoint_eval_d", C.byref(self.d), C.byref(b), _dp(w), out.ctypes.data_as(C.c_void_p))
        return out

    def naive_energy_forces(self, R, offsets, species=None):
        b, keep = self._batch(R, offsets, species)
        nj = int(keep[1][-1])
        dt = np.float64 if self.d.symreal else np.complex128
        E = np.zeros((b.nenv, self.d.nprop, self.d.ncomp), dtype=dt)
        G = np.zeros((nj, self.d.nprop, 3, self.d.ncomp), dtype=dt)
        self._call("oracle_naive_energy_forces", C.byref(self.d), C.byref(b), self.d.c,
                   E.ctypes.data_as(C.c_void_p), G.ctypes.data_as(C.c_void_p))
        return E, G

    # ---- caller side: the host loop JuLIP / ACEatoms.jl run around the per-environment calls ----
    def structure_energy_forces(self, X, first, nbr, image=None, cell=None, species=None):
        """Site energies, forces and virial of a structure, restating JuLIP's assembly loop literally:
        for each centre i: Rs = x_j + S . cell - x_i; (E_i, dV) = energy_forces(Rs); frc[j] -= dV_j; frc[i] += dV_j;
        vir -= dV_j (x) R_j.  Plain Python loop over centres and pairs (small structures only)."""
        X = np.asarray(X, dtype=np.float64).reshape(-1, 3)
        first, nbr = np.asarray(first, dtype=np.int64), np.asarray(nbr, dtype=np.int64)
        n = len(X)
        centre = np.repeat(np.arange(n), np.diff(first))
        R = X[nbr] - X[centre]
        if image is not None:
            R = R + np.asarray(image, dtype=np.float64).reshape(-1, 3) @ np.asarray(cell, dtype=np.float64).reshape(3, 3)
        sp = None if species is None else np.asarray(species, dtype=np.int32)[nbr]
        E, G = self.energy_forces(R, first, sp)              # G [pair][nprop][3][ncomp]
        F = np.zeros((n,) + G.shape[1:], dtype=G.dtype)
        W = np.zeros((self.d.nprop, 3, 3))
        for i in range(n):
            for p in range(first[i], first[i + 1]):
                F[nbr[p]] -= G[p]
                F[i] += G[p]
                if self.d.ncomp == 1:
                    W -= np.einsum("qa,b->qab", G[p, :, :, 0].real, R[p])
        return E, F, W

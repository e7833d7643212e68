"""The reference's evaluation API on top of the CUDA library.

Mirrors, for the hot path only,
  * ``evaluate`` / ``evaluate_d`` / ``evaluate_ed`` on ``Product1pBasis`` (src/product_1pbasis.jl:120-250),
    ``PIBasis`` (src/pibasis.jl:258-332) and ``SymmetricBasis`` (src/symmbasis.jl:297-336),
  * ``LinearACEModel`` with ``evaluate`` / ``grad_config`` / ``grad_params`` / ``grad_params_config`` /
    ``set_params!`` (src/linearmodel.jl:36-128) through the ``evaluator`` seam (:107-111), here the
    ``B200Evaluator`` that owns a handle of the C library,
  * the ``Vector -> ACEConfig`` convenience dispatch (src/ACE.jl:157-170).

A configuration is an ``ACEConfig`` (one environment) or a ``B200Batch`` (many, ragged).  Arrays may be
numpy (host; the library copies them to the GPU and the results back) or torch CUDA tensors (device;
nothing is copied, results are returned as torch tensors on the same device).

There is no CPU implementation behind these functions: they raise if the CUDA library is not built
or no GPU is present.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib as L
from .descriptor import basis_descriptor, coeffs_array
from .onepbasis import COMP_CAT, Product1pBasis
from .pibasis import PIBasis, PIBasisSpec
from .symmbasis import SparseCSC, SymmetricBasis

try:  # torch is plumbing only (device memory, streams); the host path works without it
    import torch
except Exception:  # pragma: no cover
    torch = None


# ------------------------------------------------------------------------------------------------
# configurations
# ------------------------------------------------------------------------------------------------
class ACEConfig:
    """One atomic environment: positions (J, 3) [+ species] (src/states.jl:422-430)."""

    def __init__(self, rr, species=None):
        self.rr = np.ascontiguousarray(rr, dtype=np.float64).reshape(-1, 3)
        self.species = None if species is None else list(species)

    def __len__(self):
        return len(self.rr)


def _is_torch(x) -> bool:
    return torch is not None and isinstance(x, torch.Tensor)


class B200Batch:
    """A ragged batch of environments: R (sum J, 3), offsets (nenv+1), optional 1-based species codes.

    This is the batched configuration type the shim adds; R has the memory of
    ``Vector{PositionState{Float64}}`` (src/states.jl:396).
    """

    def __init__(self, R, offsets, species=None):
        if _is_torch(R):
            if not R.is_cuda:
                raise ValueError("torch tensors must live on a CUDA device (use numpy for host data)")
            self.device = True
            self.R = R.contiguous().to(torch.float64).reshape(-1, 3)
            self.offsets = offsets.contiguous().to(torch.int64)
            self.species = None if species is None else species.contiguous().to(torch.int32)
            self.nJ = int(self.R.shape[0])
        else:
            self.device = False
            self.R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, 3)
            self.offsets = np.ascontiguousarray(offsets, dtype=np.int64)
            self.species = None if species is None else np.ascontiguousarray(species, dtype=np.int32)
            self.nJ = int(self.R.shape[0])
            if self.offsets[0] != 0 or self.offsets[-1] != self.nJ:
                raise ValueError("offsets must start at 0 and end at the number of neighbours")
        self.nenv = int(self.offsets.shape[0]) - 1

    @classmethod
    def from_configs(cls, cfgs: Sequence[ACEConfig], val2i=None):
        counts = [len(c) for c in cfgs]
        offsets = np.concatenate(([0], np.cumsum(counts))).astype(np.int64)
        R = np.concatenate([c.rr for c in cfgs], axis=0) if cfgs else np.zeros((0, 3))
        species = None
        if val2i is not None:
            species = np.array([val2i(s) for c in cfgs for s in c.species], dtype=np.int32)
        return cls(R, offsets, species)

    def _ptr(self, x):
        if x is None:
            return 0
        return x.data_ptr() if self.device else x.ctypes.data

    def c_batch(self) -> L.Batch:
        # nJ = len(R): known on the host even for a device-resident batch, so that the library need not read offsets back
        return L.make_batch(self.nenv, self._ptr(self.offsets), self._ptr(self.R), self._ptr(self.species),
                            L.DEVICE if self.device else L.HOST, self.nJ)

    def empty(self, shape, complex_=False):
        if self.device:
            return torch.empty(shape, dtype=torch.complex128 if complex_ else torch.float64, device=self.R.device)
        return np.empty(shape, dtype=np.complex128 if complex_ else np.float64)


def _out_ptr(x) -> int:
    if x is None:
        return 0
    return x.data_ptr() if _is_torch(x) else x.ctypes.data


# ------------------------------------------------------------------------------------------------
# handle
# ------------------------------------------------------------------------------------------------
class Handle:
    """Owns an ``aceb200_model*``."""

    def __init__(self, holder: L.DescHolder):
        self.lib = L.load()
        self.holder = holder
        self.ptr = C.c_void_p()
        L.check(self.lib.aceb200_model_create(C.byref(holder.desc), C.byref(self.ptr)))
        s = L.Sizes()
        L.check(self.lib.aceb200_model_sizes(self.ptr, C.byref(s)))
        self.s = s

    def __del__(self):
        try:
            if getattr(self, "ptr", None) and self.ptr.value:
                self.lib.aceb200_model_destroy(self.ptr)
                self.ptr = C.c_void_p()
        except Exception:
            pass

    def set_devices(self, devices):
        """Evaluate HOST batches on these CUDA devices (must include the handle's own): aceb200_set_devices."""
        d = np.ascontiguousarray(devices, dtype=np.int32)
        L.check(self.lib.aceb200_set_devices(self.ptr, len(d), d.ctypes.data_as(L.c_int32_p)))

    def set_stream(self, stream_ptr: int):
        L.check(self.lib.aceb200_set_stream(self.ptr, C.c_void_p(stream_ptr)))

    def use_current_torch_stream(self):
        if torch is not None and torch.cuda.is_available():
            self.set_stream(torch.cuda.current_stream().cuda_stream)

    def launch_count(self) -> int:
        return int(self.lib.aceb200_launch_count(self.ptr))

    def last_kernel_ms(self) -> float:
        v = C.c_double()
        L.check(self.lib.aceb200_last_kernel_ms(self.ptr, C.byref(v)))
        return v.value

    def last_stage_ms(self):
        v = (C.c_double * 3)()
        L.check(self.lib.aceb200_last_stage_ms(self.ptr, v))
        return {"pool": v[0], "adjoint": v[1], "forces": v[2]}

    def set_params(self, c: np.ndarray):
        c = np.ascontiguousarray(c, dtype=np.float64)
        L.check(self.lib.aceb200_set_params(self.ptr, c.ctypes.data_as(L.c_double_p), c.size))

    def eff_coeffs(self) -> np.ndarray:
        ct = np.zeros((self.s.nAA, self.s.nprop, self.s.ncomp), dtype=np.complex128)
        L.check(self.lib.aceb200_get_eff_coeffs(self.ptr, ct.ctypes.data_as(L.c_double_p)))
        return ct

    # ---- raw batched calls; each returns arrays shaped like the reference's results, env-major
    def _call(self, name, batch: B200Batch, *outs):
        if batch.device:
            self.use_current_torch_stream()
        cb = batch.c_batch()
        L.check(getattr(self.lib, name)(self.ptr, C.byref(cb), *[C.c_void_p(_out_ptr(o)) for o in outs]))

    def eval_A(self, b: B200Batch):
        out = b.empty((b.nenv, self.s.nA), True)
        self._call("aceb200_eval_A", b, out)
        return out

    def eval_AA(self, b: B200Batch):
        out = b.empty((b.nenv, self.s.nAA), not self.s.pireal)
        self._call("aceb200_eval_AA", b, out)
        return out

    def eval_B(self, b: B200Batch, out=None):
        if out is None:
            out = b.empty((b.nenv, self.s.nB, self.s.ncomp), not self.s.symreal)
        self._call("aceb200_eval_B", b, out)
        return out

    def eval_dA(self, b: B200Batch):
        A = b.empty((b.nenv, self.s.nA), True)
        dA = b.empty((b.nJ, self.s.nA, 3), True)
        self._call("aceb200_eval_dA", b, A, dA)
        return A, dA

    def eval_dAA(self, b: B200Batch):
        AA = b.empty((b.nenv, self.s.nAA), not self.s.pireal)
        dAA = b.empty((b.nJ, self.s.nAA, 3), not self.s.pireal)
        self._call("aceb200_eval_dAA", b, AA, dAA)
        return AA, dAA

    def eval_dB(self, b: B200Batch, B=None, dB=None):
        if B is None:
            B = b.empty((b.nenv, self.s.nB, self.s.ncomp), not self.s.symreal)
        if dB is None:
            dB = b.empty((b.nJ, self.s.nB, 3, self.s.ncomp), not self.s.symreal)
        self._call("aceb200_eval_dB", b, B, dB)
        return B, dB

    def adjoint_eval_d(self, b: B200Batch, w):
        """sum_j w_j . dB_k/dr_j, complex (nenv, nB, ncomp)."""
        if b.device:
            w = w.contiguous().to(torch.float64)
        else:
            w = np.ascontiguousarray(w, dtype=np.float64)
        if tuple(w.shape) != (b.nJ, 3):
            raise ValueError("w must have shape (sum J, 3)")
        out = b.empty((b.nenv, self.s.nB, self.s.ncomp), True)
        self._call("aceb200_adjoint_eval_d", b, w, out)
        return out

    def energy(self, b: B200Batch, E=None):
        if E is None:
            E = b.empty((b.nenv, self.s.nprop, self.s.ncomp), not self.s.symreal)
        self._call("aceb200_energy", b, E)
        return E

    def energy_forces(self, b: B200Batch, E=None, G=None):
        if E is None:
            E = b.empty((b.nenv, self.s.nprop, self.s.ncomp), not self.s.symreal)
        if G is None:
            G = b.empty((b.nJ, self.s.nprop, 3, self.s.ncomp), not self.s.symreal)
        self._call("aceb200_energy_forces", b, E, G)
        return E, G


    def energy_forces_dp(self, b: B200Batch, dp):
        """_rrule_evaluate(dp::SVector, model, cfg) (src/evaluator.jl:161-200): (E, sum_p dp[p] * grad of property p);
        the gradient has shape (sum J, 3, ncomp)."""
        dp = np.ascontiguousarray(dp, dtype=np.float64)
        if dp.shape != (self.s.nprop,):
            raise ValueError("dp must have one entry per property")
        E = b.empty((b.nenv, self.s.nprop, self.s.ncomp), not self.s.symreal)
        G = b.empty((b.nJ, 3, self.s.ncomp), not self.s.symreal)
        if b.device:
            self.use_current_torch_stream()
        cb = b.c_batch()
        L.check(self.lib.aceb200_energy_forces_dp(self.ptr, C.byref(cb), dp.ctypes.data_as(L.c_double_p),
                                                  C.c_void_p(_out_ptr(E)), C.c_void_p(_out_ptr(G))))
        return E, G

    def structure_energy_forces(self, st, virial: bool = True, E=None, F=None, W=None):
        """aceb200_structure_energy_forces: (Esite [natoms][nprop][ncomp], F [natoms][nprop][3][ncomp], W [nprop][3][3] or None).
        Pre-allocated (e.g. pinned) outputs may be passed in."""
        E = st.empty((st.natoms, self.s.nprop, self.s.ncomp)) if E is None else E
        F = st.empty((st.natoms, self.s.nprop, 3, self.s.ncomp)) if F is None else F
        W = (st.empty((self.s.nprop, 3, 3)) if virial else None) if W is None else W
        if st.device:
            self.use_current_torch_stream()      # inputs produced on a non-default torch stream stay stream-ordered
        cs = st.c_struct()
        L.check(self.lib.aceb200_structure_energy_forces(self.ptr, C.byref(cs), _out_ptr(E), _out_ptr(F), _out_ptr(W)))
        return E, F, W


def measure_fp64_tflops() -> float:
    v = C.c_double()
    L.check(L.load().aceb200_measure_fp64(C.byref(v)))
    return v.value


def measure_dmma_tflops() -> float:
    """FP64 tensor-core (mma.sync.m8n8k4.f64) throughput of the current device, TFLOP/s."""
    v = C.c_double()
    L.check(L.load().aceb200_measure_dmma(C.byref(v)))
    return v.value


# ------------------------------------------------------------------------------------------------
# handles for the three basis types
# ------------------------------------------------------------------------------------------------
def _trivial_symm(basis) -> SymmetricBasis:
    """Wrap a bare Product1pBasis / PIBasis so that it can be described to the library."""
    if isinstance(basis, Product1pBasis):
        nA = len(basis)
        spec = PIBasisSpec(np.ones(nA, dtype=np.int32), np.arange(1, nA + 1, dtype=np.int32).reshape(nA, 1))
        pib = PIBasis(basis, spec, isreal=False)
    else:
        pib = basis
    A2B = SparseCSC(0, len(pib), np.ones(len(pib) + 1, dtype=np.int32), [], np.zeros((0, 1)), 1)
    from .properties import Invariant
    from .symmetrygroups import NoSym
    return SymmetricBasis.from_parts(Invariant(), pib, A2B, NoSym(), False)


def _table_stamp(obj) -> tuple:
    """Version counters of every table a handle of `obj` is built from.  sparsify / clean_* / set_spec mutate the
    tables in place (as the reference's `sparsify!`, `clean_pibasis!`, `set_spec!` do) and bump these counters."""
    chain = [obj]
    if isinstance(obj, SymmetricBasis):
        chain += [obj.pibasis, obj.pibasis.basis1p, id(obj.A2Bmap)]
    elif isinstance(obj, PIBasis):
        chain += [obj.basis1p, id(obj.spec)]
    return tuple(getattr(o, "_version", 0) if not isinstance(o, int) else o for o in chain)


def _handle_of(obj) -> Handle:
    cached = getattr(obj, "_b200_handle", None)
    stamp = _table_stamp(obj)
    if cached is None or cached[1] != stamp:        # never evaluate with device tables of a basis that has changed since
        symm = obj if isinstance(obj, SymmetricBasis) else _trivial_symm(obj)
        cached = (Handle(basis_descriptor(symm, None)), stamp)
        obj._b200_handle = cached
    return cached[0]


def _species_map(basis1p: Product1pBasis):
    cat = basis1p.component(COMP_CAT)
    return None if cat is None else cat.val2i


def _basis1p_of(obj) -> Product1pBasis:
    if isinstance(obj, Product1pBasis):
        return obj
    if isinstance(obj, PIBasis):
        return obj.basis1p
    if isinstance(obj, SymmetricBasis):
        return obj.pibasis.basis1p
    if isinstance(obj, LinearACEModel):
        return obj.basis.pibasis.basis1p
    raise TypeError(type(obj))


def _as_batch(obj, cfg):
    """Returns (batch, single?)."""
    if isinstance(cfg, B200Batch):
        return cfg, False
    if isinstance(cfg, ACEConfig):
        single = cfg
    else:  # a plain list/array of positions: the Vector -> ACEConfig dispatch (src/ACE.jl:157-170)
        single = ACEConfig(cfg)
    if len(single) == 0:
        # @assert length(cfg) > 0 (src/product_1pbasis.jl:124)
        raise L.AceB200Error(-5, "Product1pBasis can only be evaluated with non-empty configurations")
    return B200Batch.from_configs([single], _species_map(_basis1p_of(obj))), True


def _squeeze_prop(x, ncomp):
    """Invariant values are scalars: drop the component axis of length 1."""
    return x[..., 0] if ncomp == 1 else x


# ------------------------------------------------------------------------------------------------
# evaluate / evaluate_d / evaluate_ed
# ------------------------------------------------------------------------------------------------
def evaluate(obj, cfg):
    """evaluate(basis, cfg) for the three bases; evaluate(model, cfg) for a LinearACEModel."""
    if isinstance(obj, LinearACEModel):
        return obj.evaluate(cfg)
    b, single = _as_batch(obj, cfg)
    h = _handle_of(obj)
    if isinstance(obj, Product1pBasis):
        out = h.eval_A(b)
    elif isinstance(obj, PIBasis):
        out = h.eval_AA(b)
    elif isinstance(obj, SymmetricBasis):
        out = _squeeze_prop(h.eval_B(b), h.s.ncomp)
    else:
        raise TypeError(type(obj))
    return out[0] if single else out


def evaluate_ed(obj, cfg):
    """(values, Jacobian).  The Jacobian of one configuration is (J, nbasis, 3[, ncomp]): the memory of
    the reference's column-major Matrix{DState}(nbasis, J)."""
    b, single = _as_batch(obj, cfg)
    h = _handle_of(obj)
    if isinstance(obj, Product1pBasis):
        v, d = h.eval_dA(b)
    elif isinstance(obj, PIBasis):
        v, d = h.eval_dAA(b)
    elif isinstance(obj, SymmetricBasis):
        v, d = h.eval_dB(b)
        v, d = _squeeze_prop(v, h.s.ncomp), _squeeze_prop(d, h.s.ncomp)
    else:
        raise TypeError(type(obj))
    return (v[0], d) if single else (v, d)


def evaluate_d(obj, cfg):
    return evaluate_ed(obj, cfg)[1]


# ------------------------------------------------------------------------------------------------
# LinearACEModel
# ------------------------------------------------------------------------------------------------
class B200Evaluator:
    """The evaluator that replaces ``ProductEvaluator`` (src/evaluator.jl:9-13): it owns the device
    copy of the tables and of c~ = A2Bmap' c."""

    def __init__(self, basis: SymmetricBasis, c):
        cc, self.nprop = coeffs_array(c, len(basis))
        self.handle = Handle(basis_descriptor(basis, cc))

    @property
    def coeffs(self) -> np.ndarray:
        """c~ as the reference stores it (ProductEvaluator.coeffs)."""
        return self.handle.eff_coeffs()


class LinearACEModel:
    """src/linearmodel.jl:36-56.  ``c`` is (nB,) or (nB, nprop)."""

    def __init__(self, basis: SymmetricBasis, c=None, evaluator: str = "b200"):
        if evaluator in ("recursive",):
            raise ValueError("Recursive evaluator not yet implemented")   # src/linearmodel.jl:50-51
        if evaluator not in ("b200", "standard"):
            raise ValueError("unknown evaluator")
        self.basis = basis
        c = np.zeros(len(basis)) if c is None else np.array(c, dtype=np.float64)
        self.c = c
        self.multi = c.ndim == 2
        self.evaluator = B200Evaluator(basis, c)

    # parameter wrangling (src/linearmodel.jl:63-71)
    def nparams(self):
        return len(self.c)

    def params(self):
        return self.c.copy()

    def set_params(self, c):
        c = np.array(c, dtype=np.float64)
        if c.shape != self.c.shape:
            raise ValueError("set_params!: shape of c must not change")
        self.c[...] = c
        self.evaluator.handle.set_params(c.reshape(len(self.basis), -1))
        return self

    def _shape_val(self, E, single):
        h = self.evaluator.handle
        E = _squeeze_prop(E, h.s.ncomp)              # (nenv, nprop[, ncomp])
        if not self.multi:
            E = E[:, 0]
        return E[0] if single else E

    def evaluate(self, cfg):
        b, single = _as_batch(self, cfg)
        return self._shape_val(self.evaluator.handle.energy(b), single)

    def grad_config(self, cfg):
        """(J, [nprop,] 3[, ncomp]) per configuration (src/evaluator.jl:150-200)."""
        return self.evaluate_and_grad_config(cfg)[1]

    def evaluate_and_grad_config(self, cfg):
        b, single = _as_batch(self, cfg)
        h = self.evaluator.handle
        E, G = h.energy_forces(b)
        G = _squeeze_prop(G, h.s.ncomp)
        if not self.multi:
            G = G[:, 0]
        return self._shape_val(E, single), G

    def rrule_evaluate(self, dp, cfg):
        """``_rrule_evaluate(dp, model, cfg)`` (src/evaluator.jl:161-200): the pullback of `evaluate` with respect to the
        configuration for the output cotangent ``dp`` (one number per property): (J, 3[, ncomp])."""
        b, _ = _as_batch(self, cfg)
        h = self.evaluator.handle
        dp = np.atleast_1d(np.asarray(dp, dtype=np.float64))
        _, G = h.energy_forces_dp(b, dp)
        return G[..., 0] if h.s.ncomp == 1 else G

    def energy_forces_virial(self, st, virial: bool = True):
        """JuLIP's energy / forces / virial of a whole structure (``B200Structure``) in one call: site energies
        (natoms[, nprop]), forces (natoms[, nprop], 3), virial ([nprop, ]3, 3)."""
        E, F, W = self.evaluator.handle.structure_energy_forces(st, virial)
        if self.evaluator.handle.s.ncomp == 1:
            E, F = E[..., 0], F[..., 0]
        if not self.multi:
            E, F, W = E[:, 0], F[:, 0], (None if W is None else W[0])
        return E, F, W

    def adjoint_EVAL_D(self, cfg, w):
        """src/linearmodel.jl:133-134 -> src/evaluator.jl:204-244: dB_k = sum_j w_j . dB_k/dr_j."""
        b, single = _as_batch(self, cfg)
        h = self.evaluator.handle
        out = _squeeze_prop(h.adjoint_eval_d(b, w), h.s.ncomp)
        if h.s.ncomp == 1:
            out = out.real
        return out[0] if single else out

    def grad_params(self, cfg):
        """src/linearmodel.jl:114-123: the basis values."""
        return evaluate(self.basis, cfg)

    def grad_params_config(self, cfg):
        """src/linearmodel.jl:127: the basis Jacobian."""
        return evaluate_d(self.basis, cfg)


def grad_config(model: LinearACEModel, cfg):
    return model.grad_config(cfg)


def rrule_evaluate(dp, model: LinearACEModel, cfg):
    return model.rrule_evaluate(dp, cfg)


def grad_params(model: LinearACEModel, cfg):
    return model.grad_params(cfg)


def grad_params_config(model: LinearACEModel, cfg):
    return model.grad_params_config(cfg)


def adjoint_EVAL_D(model: LinearACEModel, cfg, w):
    return model.adjoint_EVAL_D(cfg, w)


def set_params(model: LinearACEModel, c):
    return model.set_params(c)

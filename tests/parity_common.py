"""Shared body of the parity tests: every C-ABI entry point against the CPU oracle on the same inputs.

Used twice: tests/test_gpu_parity.py runs it on the B200 through csrc/libaceb200.so, and
tests/test_emu_parity.py runs it on the CPU against the thread-emulated build of the same kernel
sources (tests/emu) so that index logic is exercised in a GPU-less container.

Tolerance: FP64, <= 1e-12 relative to the largest magnitude of the compared array (BASELINE.json).
"""
import numpy as np

import ace_jl_b200 as ace
from ace_jl_b200.descriptor import basis_descriptor
from ace_jl_b200.utils import philox, rand_envs
from conftest import nspecies_of, relerr, rn_of
from oracle import Oracle

TOL = 1e-12


def compare_all(basis, nprop, Js, seed=7, jacobians=True, tol=TOL):
    rng = philox(seed)
    c = rng.random((len(basis), nprop)) - 0.5
    R, off, sp = rand_envs(rng, rn_of(basis), len(Js), Js, nspecies_of(basis))
    model = ace.LinearACEModel(basis, c if nprop > 1 else c[:, 0])
    h = model.evaluator.handle
    o = Oracle(basis_descriptor(basis, c))
    b = ace.B200Batch(R, off, sp)
    errs = {}
    errs["ctilde"] = relerr(h.eff_coeffs(), o.eff_coeffs())
    errs["A"] = relerr(h.eval_A(b), o.eval_A(R, off, sp))
    errs["AA"] = relerr(h.eval_AA(b), o.eval_AA(R, off, sp))
    errs["B"] = relerr(h.eval_B(b), o.eval_B(R, off, sp))
    if basis.real:
        E, G = h.energy_forces(b)
        Eo, Go = o.energy_forces(R, off, sp)
        errs["E"], errs["G"] = relerr(E, Eo), relerr(G, Go)
        errs["E_only"] = relerr(h.energy(b), Eo)
        # the property-contracted pullback, _rrule_evaluate(dp::SVector, ...) (src/evaluator.jl:183)
        dp = rng.standard_normal(nprop)
        Edp, Gdp = h.energy_forces_dp(b, dp)
        errs["E_dp"], errs["G_dp"] = relerr(Edp, Eo), relerr(Gdp, np.einsum("p,jpxc->jxc", dp, Go))
    else:   # a complex symmetric basis: values and Jacobians only; the model calls must refuse loudly
        import pytest
        from ace_jl_b200._lib import AceB200Error
        with pytest.raises(AceB200Error) as ei:
            h.energy_forces(b)
        assert ei.value.code == -2
    w = rng.standard_normal((len(R), 3))
    errs["adjoint_EVAL_D"] = relerr(h.adjoint_eval_d(b, w), o.adjoint_eval_d(R, off, w, sp))
    if jacobians:
        A, dA = h.eval_dA(b)
        Ao, dAo = o.eval_dA(R, off, sp)
        errs["dA"], errs["A_ed"] = relerr(dA, dAo), relerr(A, Ao)
        AA, dAA = h.eval_dAA(b)
        AAo, dAAo = o.eval_dAA(R, off, sp)
        errs["dAA"], errs["AA_ed"] = relerr(dAA, dAAo), relerr(AA, AAo)
        B, dB = h.eval_dB(b)
        Bo, dBo = o.eval_dB(R, off, sp)
        errs["dB"], errs["B_ed"] = relerr(dB, dBo), relerr(B, Bo)
    bad = {k: v for k, v in errs.items() if not (v <= tol)}
    assert not bad, f"parity failures (rel err > {tol}): {bad}; all: {errs}"
    return errs

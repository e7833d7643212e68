"""Full-size BASELINE configurations on the B200 (the sizes BASELINE.json names, not their small cousins):
sampled comparison with the CPU oracle through the C ABI, tolerance 1e-12 relative.

  config 1  evaluate / evaluate_d on the invariant SymmetricBasis, ord 3, deg 10, 30 neighbours  (bm_basis.jl:56-71)
  config 3  energy + forces, ord 4, deg 14, 60 neighbours, 2 x 10^4 environments, 500 sampled    (profile_linearmodel.jl:13-25)
  config 4  B and dB of the EuclideanVector and EuclideanMatrix bases, ord 3, deg 10              (test_euclvec.jl, test_EuclideanMatrix.jl)
  config 5  16 properties x 4 species, 99 883 AA functions: energies, and energies + 16 force fields (bm_linear.jl:68-96)

The tables come from ace_jl_b200/workloads.py, the same builders bench.py --config N uses.
"""
import numpy as np
import pytest

import ace_jl_b200 as ace
from ace_jl_b200.descriptor import basis_descriptor
from ace_jl_b200.utils import philox, rand_envs
from ace_jl_b200.workloads import WORKLOADS, build_basis, coefficients
from conftest import relerr
from oracle import Oracle

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _setup(key, nenv):
    import torch
    w = WORKLOADS[key]
    basis = build_basis(w)
    c = coefficients(w, basis)
    model = ace.LinearACEModel(basis, c if w.nprop > 1 else c[:, 0])
    rng = philox(w.seed + 5)
    R, off, sp = rand_envs(rng, basis.pibasis.basis1p.component(0), nenv, w.J, w.nspecies)
    t = lambda a: None if a is None else torch.from_numpy(a).cuda()   # noqa: E731
    return w, basis, c, model.evaluator.handle, (R, off, sp), ace.B200Batch(t(R), t(off), t(sp)), rng


def _sample(R, off, sp, J, sel):
    Rs = np.concatenate([R[off[e]:off[e + 1]] for e in sel])
    sps = None if sp is None else np.concatenate([sp[off[e]:off[e + 1]] for e in sel])
    return Rs, np.arange(len(sel) + 1, dtype=np.int64) * J, sps


def test_config1_full_size_values_and_jacobian():
    w, basis, c, h, (R, off, sp), b, rng = _setup("1", 100_000)
    assert (len(basis.pibasis.basis1p), len(basis.pibasis), len(basis)) == (73, 762, 266)
    B = h.eval_B(b).cpu().numpy()
    sel = np.sort(rng.choice(100_000, size=500, replace=False))
    Rs, offs, _ = _sample(R, off, sp, w.J, sel)
    o = Oracle(basis_descriptor(basis, c))
    assert relerr(B[sel], o.eval_B(Rs, offs)) < TOL
    # evaluate_d at config-1 size: the full Jacobian dB (266 x 30 x 3 per environment) of 2000 environments
    import torch
    n2 = 2000
    b2 = ace.B200Batch(torch.from_numpy(R[: n2 * w.J]).cuda(), torch.from_numpy(off[: n2 + 1]).cuda())
    B2, dB = h.eval_dB(b2)
    sel2 = np.sort(rng.choice(n2, size=200, replace=False))
    Rs2, offs2, _ = _sample(R, off, sp, w.J, sel2)
    Bo, dBo = o.eval_dB(Rs2, offs2)
    dBg = dB.cpu().numpy().reshape(n2, w.J, *dB.shape[1:])[sel2].reshape(len(sel2) * w.J, *dB.shape[1:])
    assert relerr(B2.cpu().numpy()[sel2], Bo) < TOL and relerr(dBg, dBo) < TOL
    assert relerr(B2.cpu().numpy(), B[:n2]) == 0.0          # the same kernel produced both


def test_config3_full_size_energy_forces():
    w, basis, c, h, (R, off, sp), b, rng = _setup("3", 20_000)
    assert len(basis.pibasis.basis1p) == 164 and len(basis) == 2566 and abs(len(basis.pibasis) - 15945) <= 20   # round-off dependent cleaning
    E, G = h.energy_forces(b)
    E, G = E.cpu().numpy(), G.cpu().numpy()
    sel = np.sort(rng.choice(20_000, size=500, replace=False))
    Rs, offs, _ = _sample(R, off, sp, w.J, sel)
    Eo, Go = Oracle(basis_descriptor(basis, c)).energy_forces(Rs, offs)
    Gs = G.reshape(20_000, w.J, *G.shape[1:])[sel].reshape(len(sel) * w.J, *G.shape[1:])
    assert relerr(E[sel], Eo) < TOL and relerr(Gs, Go) < TOL
    assert relerr(h.energy(b).cpu().numpy(), E) < TOL        # evaluate(model, cfg) alone: the energy-only stream


@pytest.mark.parametrize("key,nB,ncomp", [("4a", 300, 3), ("4", 769, 9)])
def test_config4_full_size_equivariant_values_and_jacobian(key, nB, ncomp):
    w, basis, c, h, (R, off, sp), b, rng = _setup(key, 20_000)
    assert len(basis) == nB and basis.A2Bmap.ncomp == ncomp
    B = h.eval_B(b).cpu().numpy()
    sel = np.sort(rng.choice(20_000, size=200, replace=False))
    Rs, offs, _ = _sample(R, off, sp, w.J, sel)
    o = Oracle(basis_descriptor(basis, None))
    assert relerr(B[sel], o.eval_B(Rs, offs)) < TOL
    # dB of a few environments (30 x nB x 3 x ncomp each)
    import torch
    n2 = 24
    b2 = ace.B200Batch(torch.from_numpy(R[: n2 * w.J]).cuda(), torch.from_numpy(off[: n2 + 1]).cuda())
    B2, dB = h.eval_dB(b2)
    Bo, dBo = o.eval_dB(R[: n2 * w.J], off[: n2 + 1])
    assert relerr(B2.cpu().numpy(), Bo) < TOL and relerr(dB.cpu().numpy(), dBo) < TOL


def test_config5_full_size_multiproperty_species():
    w, basis, c, h, (R, off, sp), b, rng = _setup("5", 4_000)
    assert (len(basis.pibasis.basis1p), len(basis.pibasis), len(basis)) == (456, 99883, 22091)
    E = h.energy(b).cpu().numpy()
    Ef, G = h.energy_forces(b)
    Ef, G = Ef.cpu().numpy(), G.cpu().numpy()
    assert E.shape == (4_000, 16, 1) and G.shape == (4_000 * w.J, 16, 3, 1)
    sel = np.sort(rng.choice(4_000, size=60, replace=False))
    Rs, offs, sps = _sample(R, off, sp, w.J, sel)
    Eo, Go = Oracle(basis_descriptor(basis, c)).energy_forces(Rs, offs, sps)
    Gs = G.reshape(4_000, w.J, *G.shape[1:])[sel].reshape(len(sel) * w.J, *G.shape[1:])
    assert relerr(E[sel], Eo) < TOL and relerr(Ef[sel], Eo) < TOL and relerr(Gs, Go) < TOL

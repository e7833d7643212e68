"""Parity of the CUDA path with the CPU oracle, through the C ABI, on the B200.

Small cases: every entry point vs the oracle on identical seeded inputs (all model families of
BASELINE.json's configs, ragged batches, host- and device-resident buffers, forced multi-chunk).
Full size (config 2's 40 neighbours, 10^5..10^6 environments): size-independent properties --
rotation/reflection/permutation invariance of E and equivariance of the forces, linearity in the
parameters, a sampled oracle comparison, finite differences.
Edge cases the reference tests or guards: empty configuration, unknown category, neighbours beyond the
cutoff, a neighbour on the pole, a single neighbour.
Tolerance everywhere: 1e-12 relative (FP64), as BASELINE.json states.
"""
import os

import numpy as np
import pytest

import ace_jl_b200 as ace
from ace_jl_b200 import _lib
from ace_jl_b200.descriptor import basis_descriptor
from ace_jl_b200.utils import philox, rand_envs, rand_rot
from conftest import make_basis, nspecies_of, relerr, rn_of
from oracle import Oracle
from parity_common import TOL, compare_all

pytestmark = pytest.mark.gpu

CASES = [
    ("inv_simple_3_6", 1, [3, 10, 1, 35, 7, 2, 140, 64, 33]),
    ("inv_sparse_3_10", 1, [30] * 8),                       # config 1
    ("inv_sparse_3_12", 1, [40] * 8 + [1, 77]),             # config 2
    ("inv_sparse_4_8", 4, [5, 60, 33, 60]),                 # order 4 (config 3's family), multi-property
    ("euclvec_3_5", 1, [3, 30, 1, 35]),                     # config 4
    ("euclmat_2_5", 2, [3, 30]),                            # config 4
    ("species_3_5", 16, [5, 40, 33, 130]),                  # config 5's family: 16 properties, 4 species
    ("inv_morse_2_6", 1, [6, 20]),
    ("inv_agnesi_2_6", 1, [6, 20]),
    ("inv_highL_2_12", 1, [5, 17]),
    ("inv_sparse_4_8", 1, [7, 21]),                       # order 4 through the grouped single-channel stream
    ("inv_complexB_2_5", 1, [6, 20, 33]),                   # complex B / dB outputs (symreal = false)
]


def test_native_library_is_the_one_loaded():
    lib = _lib.load()
    assert lib._name == _lib.LIB_PATH and lib.aceb200_device_count() >= 1


@pytest.mark.parametrize("kind,nprop,Js", CASES)
def test_all_entry_points_match_oracle(kind, nprop, Js):
    compare_all(make_basis(kind), nprop, Js)


def test_multichunk_path_matches_oracle(monkeypatch):
    monkeypatch.setenv("ACEB200_CHUNK_ENVS", "32")
    compare_all(make_basis("inv_simple_3_6"), 2, [7] * 70 + [3, 129, 1], seed=21)
    compare_all(make_basis("species_3_5"), 1, [9] * 45, seed=22)
    # many chunks over all four pipeline lanes, ragged small environments
    monkeypatch.setenv("ACEB200_CHUNK_ENVS", "64")
    rng = np.random.default_rng(5)
    compare_all(make_basis("inv_simple_3_6"), 1, [int(j) for j in rng.integers(1, 7, 700)], seed=23, jacobians=False)


def test_device_resident_batch_matches_host_batch():
    import torch
    basis = make_basis("inv_sparse_3_10")
    rng = philox(31)
    c = rng.random(len(basis)) - 0.5
    model = ace.LinearACEModel(basis, c)
    R, off, _ = rand_envs(rng, rn_of(basis), 300, rng.integers(1, 50, size=300))
    E, G = model.evaluator.handle.energy_forces(ace.B200Batch(R, off))
    bd = ace.B200Batch(torch.from_numpy(R).cuda(), torch.from_numpy(off).cuda())
    Ed, Gd = model.evaluator.handle.energy_forces(bd)
    assert Ed.is_cuda and np.array_equal(Ed.cpu().numpy(), E) and np.array_equal(Gd.cpu().numpy(), G)
    Bd = ace.evaluate(basis, bd)
    assert relerr(Bd.cpu().numpy(), ace.evaluate(basis, ace.B200Batch(R, off))) == 0.0
    # a non-default torch stream is honoured
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        E2, _ = model.evaluator.handle.energy_forces(bd)
    s.synchronize()
    assert np.array_equal(E2.cpu().numpy(), E)


def test_reference_api_surface():
    """evaluate / evaluate_d / evaluate_ed / grad_config / set_params! with the reference's call shapes
    (test_linearmodel.jl:47-90, ACE.jl:157-170)."""
    basis = make_basis("inv_simple_3_6")
    rng = philox(32)
    c = rng.random(len(basis)) - 0.5
    model = ace.LinearACEModel(basis, c)
    o = Oracle(basis_descriptor(basis, c.reshape(-1, 1)))
    R, off, _ = rand_envs(rng, rn_of(basis), 1, 10)
    cfg = ace.ACEConfig(R)
    B = ace.evaluate(basis, cfg)
    assert B.shape == (len(basis),) and relerr(B, o.eval_B(R, off)[0, :, 0]) < TOL
    assert relerr(ace.evaluate(basis, R), B) == 0.0                      # Vector -> ACEConfig dispatch
    assert relerr(ace.evaluate(basis.pibasis, cfg), o.eval_AA(R, off)[0]) < TOL
    assert relerr(ace.evaluate(basis.pibasis.basis1p, cfg), o.eval_A(R, off)[0]) < TOL
    dB = ace.evaluate_d(basis, cfg)
    assert dB.shape == (10, len(basis), 3) and relerr(dB, o.eval_dB(R, off)[1][..., 0]) < TOL
    A, dA = ace.evaluate_ed(basis.pibasis.basis1p, cfg)
    assert relerr(dA, o.eval_dA(R, off)[1]) < TOL
    val = ace.evaluate(model, cfg)
    g = ace.grad_config(model, cfg)
    Eo, Go = o.energy_forces(R, off)
    assert abs(val - Eo[0, 0, 0]) < TOL * abs(Eo[0, 0, 0]) and g.shape == (10, 3) and relerr(g, Go[:, 0, :, 0]) < TOL
    # naive formulas of the reference's own test: val = sum c_i B_i, g = sum c_i dB_i
    assert abs(val - np.dot(c, B)) < 1e-11 * max(1, abs(val))
    assert relerr(g, np.einsum("i,jik->jk", c, dB)) < 1e-10
    assert relerr(ace.grad_params(model, cfg), B) == 0.0
    # set_params!
    c2 = rng.random(len(basis)) - 0.5
    ace.set_params(model, c2)
    assert abs(ace.evaluate(model, cfg) - np.dot(c2, B)) < 1e-11 * max(1, abs(np.dot(c2, B)))
    assert relerr(model.evaluator.coeffs, Oracle(basis_descriptor(basis, c2.reshape(-1, 1))).eff_coeffs()) < 1e-14


def test_adjoint_eval_d_api():
    """adjoint_EVAL_D(model, cfg, w) (src/linearmodel.jl:133-134): equals sum_j w_j . dB_k/dr_j for an invariant model."""
    basis = make_basis("inv_simple_3_6")
    rng = philox(36)
    model = ace.LinearACEModel(basis, rng.random(len(basis)) - 0.5)
    R, off, _ = rand_envs(rng, rn_of(basis), 1, 14)
    w = rng.standard_normal((14, 3))
    got = ace.adjoint_EVAL_D(model, ace.ACEConfig(R), w)
    dB = ace.evaluate_d(basis, ace.ACEConfig(R))                       # (J, nB, 3)
    assert got.shape == (len(basis),) and relerr(got, np.einsum("jx,jkx->k", w, dB)) < 1e-11


def test_multiproperty_matches_single_property_models():
    """property i of an N-property model == the single-property model with c[:, i] (test_multiprop.jl:24-92)."""
    for kind in ("inv_simple_3_6", "euclvec_3_5"):
        basis = make_basis(kind)
        rng = philox(33)
        c = rng.random((len(basis), 3)) - 0.5
        R, off, _ = rand_envs(rng, rn_of(basis), 5, [4, 9, 17, 30, 2])
        b = ace.B200Batch(R, off)
        E, G = ace.LinearACEModel(basis, c).evaluator.handle.energy_forces(b)
        for i in range(3):
            Ei, Gi = ace.LinearACEModel(basis, c[:, i].copy()).evaluator.handle.energy_forces(b)
            assert relerr(E[:, i], Ei[:, 0]) < TOL and relerr(G[:, i], Gi[:, 0]) < TOL


def test_edge_cases():
    basis = make_basis("inv_simple_3_6")
    rng = philox(34)
    c = rng.random(len(basis)) - 0.5
    model = ace.LinearACEModel(basis, c)
    h = model.evaluator.handle
    o = Oracle(basis_descriptor(basis, c.reshape(-1, 1)))
    rcut = rn_of(basis).meta["rcut"]
    R, off, _ = rand_envs(rng, rn_of(basis), 1, 6)
    # neighbours beyond the cutoff contribute exactly nothing (orthpolys.jl:41-46)
    far = np.array([[rcut + 0.3, 0.0, 0.1], [0.0, 0.0, 2 * rcut]])
    R2 = np.concatenate([R, far])
    E1, G1 = h.energy_forces(ace.B200Batch(R, off))
    E2, G2 = h.energy_forces(ace.B200Batch(R2, np.array([0, 8])))
    assert np.array_equal(E1, E2) and np.array_equal(G2[:6], G1) and np.all(G2[6:] == 0.0)
    # a neighbour exactly on the pole, and a single neighbour (test_ylm.jl:52-67)
    Rp = np.array([[0.0, 0.0, 1.3], [0.0, 0.0, -0.9], [0.4, 0.1, 0.8]])
    for offp in ([0, 3], [0, 1, 2, 3]):
        Ep, Gp = h.energy_forces(ace.B200Batch(Rp, np.array(offp)))
        Eo, Go = o.energy_forces(Rp, np.array(offp))
        assert np.all(np.isfinite(Gp)) and relerr(Ep, Eo) < TOL and relerr(Gp, Go) < 1e-11
    # empty configuration: the reference asserts (product_1pbasis.jl:124)
    with pytest.raises(_lib.AceB200Error) as ei:
        h.energy(ace.B200Batch(R, np.array([0, 6, 6])))
    assert ei.value.code == -5
    with pytest.raises(_lib.AceB200Error):
        ace.evaluate(basis, [])
    # an empty batch is fine
    E0 = h.energy(ace.B200Batch(np.zeros((0, 3)), np.array([0])))
    assert E0.shape[0] == 0
    # unknown category (discrete1pbasis.jl:39)
    sb = make_basis("species_3_5")
    hs = ace.LinearACEModel(sb, np.zeros(len(sb))).evaluator.handle
    with pytest.raises(_lib.AceB200Error) as ei:
        hs.energy(ace.B200Batch(R, off, np.array([1, 2, 3, 4, 5, 1], dtype=np.int32)))
    assert ei.value.code == -6
    assert h.launch_count() > 0 and h.last_kernel_ms() >= 0.0


def test_concurrent_calls_on_one_handle():
    """Several host threads may evaluate on one handle (the reference's per-thread pools, src/utils/pools.jl:44-75)."""
    import threading
    basis = make_basis("inv_sparse_3_10")
    rng = philox(37)
    h = ace.LinearACEModel(basis, rng.random(len(basis)) - 0.5).evaluator.handle
    batches = [rand_envs(rng, rn_of(basis), 500 + 37 * k, 20 + k)[:2] for k in range(6)]
    ref = [h.energy_forces(ace.B200Batch(R, off)) for R, off in batches]
    out = [None] * len(batches)

    def work(k):
        for _ in range(5):
            out[k] = h.energy_forces(ace.B200Batch(*batches[k]))

    th = [threading.Thread(target=work, args=(k,)) for k in range(len(batches))]
    [t.start() for t in th]
    [t.join() for t in th]
    for k in range(len(batches)):
        assert np.array_equal(out[k][0], ref[k][0]) and np.array_equal(out[k][1], ref[k][1])


def test_finite_difference_forces():
    for kind, nprop in (("inv_sparse_3_12", 1), ("euclvec_3_5", 2), ("species_3_5", 2)):
        basis = make_basis(kind)
        rng = philox(35)
        c = rng.random((len(basis), nprop)) - 0.5
        h = ace.LinearACEModel(basis, c).evaluator.handle
        R, off, sp = rand_envs(rng, rn_of(basis), 2, [12, 25], nspecies_of(basis))
        _, G = h.energy_forces(ace.B200Batch(R, off, sp))
        for j, k in ((3, 0), (20, 2), (36, 1)):
            hh = 1e-6
            Rp, Rm = R.copy(), R.copy()
            Rp[j, k] += hh
            Rm[j, k] -= hh
            e = 0 if j < 12 else 1
            fd = (h.energy(ace.B200Batch(Rp, off, sp))[e] - h.energy(ace.B200Batch(Rm, off, sp))[e]) / (2 * hh)
            assert np.abs(fd - G[j, :, k, :]).max() < 2e-6 * max(1.0, np.abs(G).max())


def _config2_model(seed=20242):
    basis = make_basis("inv_sparse_3_12")
    c = philox(seed + 1000).random(len(basis)) - 0.5
    return basis, c, ace.LinearACEModel(basis, c)


def test_config2_sampled_oracle_parity_and_properties():
    """BASELINE config 2 at scale: 2*10^5 environments x 40 neighbours on the device."""
    import torch
    basis, c, model = _config2_model()
    h = model.evaluator.handle
    rng = philox(20242)
    nenv, J = 200_000, 40
    R, off, _ = rand_envs(rng, rn_of(basis), nenv, J)
    Rd, offd = torch.from_numpy(R).cuda(), torch.from_numpy(off).cuda()
    E, G = h.energy_forces(ace.B200Batch(Rd, offd))
    E, G = E.cpu().numpy(), G.cpu().numpy()
    assert np.all(np.isfinite(E)) and np.all(np.isfinite(G))
    # (a) a sample of environments against the oracle
    sel = rng.choice(nenv, size=1500, replace=False)
    Rs = np.concatenate([R[off[e]:off[e + 1]] for e in sel])
    offs = np.arange(len(sel) + 1) * J
    Eo, Go = Oracle(basis_descriptor(basis, c.reshape(-1, 1))).energy_forces(Rs, offs)
    Gs = np.concatenate([G[off[e]:off[e + 1]] for e in sel])
    assert relerr(E[sel], Eo) < TOL and relerr(Gs, Go) < TOL
    # (b) O(3) x permutation: E invariant, forces co-rotate
    Q = rand_rot(rng) * -1
    perm = (np.arange(nenv)[:, None] * J + np.argsort(rng.random((nenv, J)), axis=1)).ravel()
    Rq = np.ascontiguousarray((R @ Q.T)[perm])
    Eq, Gq = h.energy_forces(ace.B200Batch(torch.from_numpy(Rq).cuda(), offd))
    Eq, Gq = Eq.cpu().numpy(), Gq.cpu().numpy()
    assert relerr(Eq, E) < 1e-11
    assert relerr(Gq[:, 0, :, 0], (G[:, 0, :, 0] @ Q.T)[perm]) < 1e-11
    # (c) linearity in the parameters
    c2 = rng.random(len(basis)) - 0.5
    model.set_params(c2)
    E2, G2 = h.energy_forces(ace.B200Batch(Rd, offd))
    model.set_params(c + c2)
    E3, G3 = h.energy_forces(ace.B200Batch(Rd, offd))
    assert relerr(E3.cpu().numpy(), E + E2.cpu().numpy()) < 1e-11
    assert relerr(G3.cpu().numpy(), G + G2.cpu().numpy()) < 1e-11
    # (d) host-resident call (chunked copies) gives the same numbers as the device-resident one
    model.set_params(c)
    Eh, Gh = h.energy_forces(ace.B200Batch(R[: 40 * 50_000], off[:50_001]))
    assert np.array_equal(Eh, E[:50_000]) and np.array_equal(Gh, G[: 40 * 50_000])
    # (e) evaluate(model, cfg) alone walks the energy-only stream (every AA function once): same energies
    Eonly = h.energy(ace.B200Batch(Rd, offd)).cpu().numpy()
    assert relerr(Eonly, E) < TOL


def test_malformed_offsets_are_rejected_not_dereferenced():
    """A malformed interior offset (e.g. [0, 100, 5, 10]) must come back as an EDESC error from both the HOST path
    (validated on the host) and the DEVICE path (validated by a kernel; every offset-consuming kernel is gated on it),
    and the handle must stay usable."""
    import torch
    basis = make_basis("inv_simple_3_6")
    rng = philox(61)
    h = ace.LinearACEModel(basis, rng.random(len(basis)) - 0.5).evaluator.handle
    R, off, _ = rand_envs(rng, rn_of(basis), 3, [4, 3, 3])
    good = h.energy_forces(ace.B200Batch(R, off))
    bad = np.array([0, 100, 5, 10], dtype=np.int64)
    hb = ace.B200Batch(R, off)
    hb.offsets = bad                                       # bypass the Python-side check
    with pytest.raises(_lib.AceB200Error) as ei:
        h.energy_forces(hb)
    assert ei.value.code == -1
    db = ace.B200Batch(torch.from_numpy(R).cuda(), torch.from_numpy(bad).cuda())
    for call in (h.energy_forces, h.eval_B, h.eval_dB):
        with pytest.raises(_lib.AceB200Error) as ei:
            call(db)
        assert ei.value.code == -1
    again = h.energy_forces(ace.B200Batch(torch.from_numpy(R).cuda(), torch.from_numpy(off).cuda()))
    assert np.array_equal(again[0].cpu().numpy(), good[0]) and np.array_equal(again[1].cpu().numpy(), good[1])

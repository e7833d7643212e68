"""Caller side of the path: structure + neighbour list -> site energies, forces, virial.

CPU tests pin the host-side helpers and the oracle's assembly loop (finite differences of the total energy with
respect to atom positions and to a homogeneous strain: the sign conventions of JuLIP's forces / virial); GPU
tests compare aceb200_structure_energy_forces with that oracle through the C ABI (tolerance 1e-12 relative).
"""
import numpy as np
import pytest

import ace_jl_b200 as ace
from ace_jl_b200.descriptor import basis_descriptor
from ace_jl_b200.structure import B200Structure, neighbourlist, reverse_pairs
from ace_jl_b200.utils import philox
from conftest import make_basis, relerr
from oracle import Oracle

TOL = 1e-12
RCUT = 2.5


def cluster(rng, n, box):
    """Random points in a box with a minimum separation above the inner cutoff (rin = 0.5)."""
    X = []
    while len(X) < n:
        x = rng.random(3) * box
        if all(np.linalg.norm(x - y) > 0.8 for y in X):
            X.append(x)
    return np.array(X)


def periodic_crystal(rng, reps=(2, 2, 2), a=1.5, jitter=0.1):
    """A jittered simple-cubic crystal in a (slightly sheared) periodic cell smaller than 2 rcut: every pair list
    contains several images of the same atom."""
    g = np.stack(np.meshgrid(*[np.arange(r) for r in reps], indexing="ij"), -1).reshape(-1, 3).astype(float)
    cell = np.diag(np.array(reps, dtype=float) * a)
    cell[1, 0] = 0.2
    X = g * a + (rng.random(g.shape) - 0.5) * 2 * jitter
    return X, cell


def test_neighbourlist_and_reverse_pairs():
    rng = philox(5)
    X, cell = periodic_crystal(rng)
    first, nbr, image, rev = neighbourlist(X, RCUT, cell, (True, True, True))
    shift = image.astype(float) @ cell
    centre = np.repeat(np.arange(len(X)), np.diff(first))
    R = X[nbr] + shift - X[centre]
    r = np.linalg.norm(R, axis=1)
    assert r.max() < RCUT and r.min() > 0 and np.all(np.diff(first) > 10)
    # brute-force count over a 5 x 5 x 5 block of images
    cnt = 0
    for i in range(-2, 3):
        for j in range(-2, 3):
            for k in range(-2, 3):
                s = i * cell[0] + j * cell[1] + k * cell[2]
                d = np.linalg.norm(X[None] + s - X[:, None], axis=2)
                cnt += int(((d < RCUT) & (d > 0)).sum())
    assert cnt == len(nbr)
    # the reverse table is an involution that flips (centre, neighbour, shift)
    assert np.all(rev >= 0) and np.array_equal(rev[rev], np.arange(len(nbr)))
    assert np.array_equal(centre[rev], nbr) and np.array_equal(nbr[rev], centre) and np.array_equal(image[rev], -image)
    assert np.array_equal(reverse_pairs(first, nbr, image), rev)
    st = B200Structure(X, first, nbr, image, cell)
    assert np.allclose(st.environments()[0], R)
    # open boundaries: no shifts
    f2, n2, s2, r2 = neighbourlist(X, RCUT)
    assert s2 is None and np.array_equal(r2[r2], np.arange(len(n2)))


def total_energy(o, X, cell, pbc, species=None):
    first, nbr, image, _ = neighbourlist(X, RCUT, cell, pbc, with_rev=False)
    E, _, _ = o.structure_energy_forces(X, first, nbr, image, cell, species)
    return E.sum(axis=0)[:, 0]


def test_oracle_assembly_is_minus_the_gradient_of_the_total_energy():
    """forces = -dE/dx (JuLIP forces(V, at)) and virial = -dE/d(strain) (JuLIP virial): central differences."""
    basis = make_basis("inv_simple_3_6")
    rng = philox(6)
    c = rng.random((len(basis), 1)) - 0.5
    o = Oracle(basis_descriptor(basis, c))
    X, cell = periodic_crystal(rng)
    pbc = (True, True, True)
    first, nbr, image, _ = neighbourlist(X, RCUT, cell, pbc)
    E, F, W = o.structure_energy_forces(X, first, nbr, image, cell)
    h = 1e-5
    for (i, a) in [(0, 0), (3, 1), (7, 2)]:
        Xp, Xm = X.copy(), X.copy()
        Xp[i, a] += h; Xm[i, a] -= h
        dE = (total_energy(o, Xp, cell, pbc) - total_energy(o, Xm, cell, pbc)) / (2 * h)
        assert abs(-dE[0] - F[i, 0, a, 0]) < 1e-7 * max(1.0, np.abs(F).max())
    for (a, b) in [(0, 0), (1, 2), (2, 1)]:
        eps = np.zeros((3, 3)); eps[a, b] = h
        Ep = total_energy(o, X @ (np.eye(3) + eps).T, cell @ (np.eye(3) + eps).T, pbc)
        Em = total_energy(o, X @ (np.eye(3) - eps).T, cell @ (np.eye(3) - eps).T, pbc)
        dE = (Ep - Em) / (2 * h)
        assert abs(-dE[0] - W[0, a, b]) < 1e-7 * max(1.0, np.abs(W).max())
    # Newton's third law: the forces of a closed system sum to zero
    assert np.abs(F.sum(axis=0)).max() < 1e-12 * np.abs(F).max() * len(X)


def tiny_cell(rng):
    """Two atoms in a cell smaller than the cutoff: every atom sees several images of itself (pairs with i == j)."""
    cell = np.array([[1.9, 0.0, 0.0], [0.3, 2.1, 0.0], [0.0, 0.2, 2.0]])
    X = np.array([[0.1, 0.2, 0.1], [1.0, 1.1, 0.9]]) + 0.05 * rng.random((2, 3))
    return X, cell


GPU_CASES = [("inv_simple_3_6", 1, True, False), ("inv_sparse_3_12", 1, True, True), ("inv_simple_3_6", 3, False, True),
             ("species_3_5", 4, True, True), ("inv_sparse_4_8", 1, False, False)]


@pytest.mark.gpu
@pytest.mark.parametrize("kind,nprop,periodic,use_rev", GPU_CASES)
def test_structure_matches_oracle(kind, nprop, periodic, use_rev, monkeypatch):
    basis = make_basis(kind)
    rng = philox(17)
    c = rng.random((len(basis), nprop)) - 0.5
    model = ace.LinearACEModel(basis, c if nprop > 1 else c[:, 0])
    o = Oracle(basis_descriptor(basis, c))
    if periodic:
        X, cell = periodic_crystal(rng, reps=(3, 2, 2))
        first, nbr, image, rev = neighbourlist(X, RCUT, cell, (True, True, True))
    else:
        X, cell = cluster(rng, 40, 4.0), None
        first, nbr, image, rev = neighbourlist(X, RCUT)
    assert np.all(np.diff(first) > 0)
    species = None
    if kind.startswith("species"):
        species = rng.integers(1, 5, len(X)).astype(np.int32)
    Eo, Fo, Wo = o.structure_energy_forces(X, first, nbr, image, cell, species)
    for chunk_mb in (None, "0.002"):                      # one chunk, then many chunks of centres
        if chunk_mb:
            monkeypatch.setenv("ACEB200_STRUCT_MB", chunk_mb)
        st = B200Structure(X, first, nbr, image, cell, species, rev if use_rev else None)
        E, F, W = model.evaluator.handle.structure_energy_forces(st)
        assert relerr(E, Eo) < TOL and relerr(F, Fo) < TOL and relerr(W, Wo) < TOL
    # a list that is not sorted within a centre (the device-side reverse search falls back to a scan)
    perm = np.concatenate([first[i] + rng.permutation(first[i + 1] - first[i]) for i in range(len(X))])
    Ep, Fp, Wp = model.evaluator.handle.structure_energy_forces(
        B200Structure(X, first, nbr[perm], None if image is None else image[perm], cell, species))
    assert relerr(Ep, Eo) < TOL and relerr(Fp, Fo) < TOL and relerr(Wp, Wo) < TOL
    # packed neighbour words (ACEB200_NBR_PACKED: image shift in the upper six bits, 4 B per pair): same bits out, with the
    # caller's reverse table and with the device-side search, sorted and unsorted
    if periodic and np.abs(image).max() <= 1:
        from ace_jl_b200.structure import pack_neighbours
        wp = pack_neighbours(nbr, image)
        for r, w_ in ((rev if use_rev else None, wp), (None, wp[perm])):
            Ek, Fk, Wk = model.evaluator.handle.structure_energy_forces(B200Structure(X, first, w_, None, cell, species, r, packed=True))
            assert relerr(Ek, Eo) < TOL and relerr(Fk, Fo) < TOL and relerr(Wk, Wo) < TOL
    # the model-level wrapper squeezes like evaluate / grad_config
    E1, F1, W1 = model.energy_forces_virial(st)
    assert E1.shape == ((len(X),) if nprop == 1 else (len(X), nprop)) and F1.shape[-1] == 3
    # the assembly is a gather in a fixed order: a device-resident structure, with the caller's reverse table or
    # the one found on the device, gives the same bits
    import torch
    t = lambda a: None if a is None else torch.from_numpy(a).cuda()   # noqa: E731
    sd = B200Structure(t(X), t(first), t(nbr), t(image), cell, t(species), t(rev) if use_rev else None)
    Ed, Fd, Wd = model.evaluator.handle.structure_energy_forces(sd)
    assert np.array_equal(Fd.cpu().numpy(), F) and np.array_equal(Ed.cpu().numpy(), E)


@pytest.mark.gpu
def test_structure_with_self_images():
    basis = make_basis("inv_simple_3_6")
    rng = philox(23)
    c = rng.random((len(basis), 1)) - 0.5
    model = ace.LinearACEModel(basis, c[:, 0])
    X, cell = tiny_cell(rng)
    first, nbr, image, rev = neighbourlist(X, RCUT, cell, (True, True, True))
    centre = np.repeat(np.arange(2), np.diff(first))
    assert np.any(nbr == centre) and np.all(rev >= 0)
    Eo, Fo, Wo = Oracle(basis_descriptor(basis, c)).structure_energy_forces(X, first, nbr, image, cell)
    for r in (rev, None):
        E, F, W = model.evaluator.handle.structure_energy_forces(B200Structure(X, first, nbr, image, cell, None, r))
        assert relerr(E, Eo) < TOL and relerr(F, Fo) < 1e-10 and relerr(W, Wo) < TOL   # forces nearly cancel: looser


@pytest.mark.gpu
def test_structure_errors():
    from ace_jl_b200._lib import AceB200Error
    basis = make_basis("inv_simple_3_6")
    model = ace.LinearACEModel(basis, philox(1).random(len(basis)))
    X = cluster(philox(2), 12, 2.5)
    first, nbr, image, rev = neighbourlist(X, RCUT)
    bad = nbr.copy(); bad[3] = 99
    with pytest.raises(AceB200Error) as ei:
        model.evaluator.handle.structure_energy_forces(B200Structure(X, first, bad))
    assert ei.value.code == -1
    # an atom without neighbours is an empty configuration (src/product_1pbasis.jl:124)
    X2 = np.vstack([X, [50.0, 50.0, 50.0]])
    f2, n2, s2, r2 = neighbourlist(X2, RCUT)
    with pytest.raises(AceB200Error) as ei:
        model.evaluator.handle.structure_energy_forces(B200Structure(X2, f2, n2))
    assert ei.value.code == -5


def test_fcc_structure_generator_matches_the_generic_neighbour_list():
    from ace_jl_b200.utils import fcc_structure
    X, cell, first, nbr, image = fcc_structure(philox(9), 2)
    assert len(X) == 32 and np.all(np.diff(first) == 42)
    f2, n2, i2, _ = neighbourlist(X, RCUT, cell, (True, True, True), with_rev=False)
    assert np.array_equal(first, f2)
    key = lambda f, n, im: sorted(zip(np.repeat(np.arange(len(f) - 1), np.diff(f)).tolist(), n.tolist(), map(tuple, im.tolist())))  # noqa: E731
    assert key(first, nbr, image) == key(f2, n2, i2)
    rev = reverse_pairs(first, nbr, image)
    assert np.all(rev >= 0) and np.array_equal(rev[rev], np.arange(len(nbr)))


@pytest.mark.gpu
def test_large_periodic_structure_properties():
    """BASELINE config 2's model on a 10^5-atom periodic crystal (42 neighbours): size-independent properties of the
    caller-side path, plus a sampled comparison with the oracle and with the per-environment entry point."""
    import math
    from ace_jl_b200.utils import RnYlm_1pbasis, fcc_structure
    Bsel = ace.SparseBasis(maxorder=3, p=1, default_maxdeg=12, weight={"n": 1.0, "l": 1.5})
    basis = ace.SymmetricBasis(ace.Invariant(), RnYlm_1pbasis(maxdeg=12, maxL=math.ceil(12 / 1.5), Bsel=Bsel), Bsel)
    rng = philox(31)
    c = rng.random(len(basis)) - 0.5
    model = ace.LinearACEModel(basis, c)
    h = model.evaluator.handle
    X, cell, first, nbr, image = fcc_structure(rng, 29)            # 97 556 atoms, 4.1 * 10^6 pairs
    st = B200Structure(X, first, nbr, image, cell)
    E, F, W = h.structure_energy_forces(st)
    F3, W3 = F[:, 0, :, 0], W[0]
    assert np.all(np.isfinite(E)) and np.all(np.isfinite(F))
    # Newton's third law over the periodic cell, and a symmetric virial (rotation invariance of the energy)
    assert np.abs(F3.sum(axis=0)).max() < 1e-10 * np.abs(F3).max() * math.sqrt(len(X))
    assert np.abs(W3 - W3.T).max() < 1e-10 * np.abs(W3).max()
    # a rigid translation changes nothing; the result does not depend on the chunking of the centres
    E2, F2, W2 = h.structure_energy_forces(B200Structure(X + np.array([0.3, -1.1, 2.7]), first, nbr, image, cell))
    assert relerr(E2, E) < 1e-11 and relerr(F2, F) < 1e-9
    import os
    os.environ["ACEB200_STRUCT_MB"] = "1"
    try:
        E4, F4, W4 = h.structure_energy_forces(st)
    finally:
        del os.environ["ACEB200_STRUCT_MB"]
    assert np.array_equal(E4, E) and np.array_equal(F4, F)
    # the per-environment entry point on the same environments gives the same site energies, and its pair gradients
    # assemble to the same forces
    R, off, _, centre = st.environments()
    Eb, G = h.energy_forces(ace.B200Batch(R, off))
    assert np.array_equal(Eb, E)
    Fa = np.zeros_like(F3)
    np.add.at(Fa, centre, G[:, 0, :, 0])
    np.add.at(Fa, nbr, -G[:, 0, :, 0])
    assert relerr(Fa, F3) < 1e-11
    assert relerr(-np.einsum("pa,pb->ab", G[:, 0, :, 0], R), W3) < 1e-11
    # a sample of centres against the oracle
    o = Oracle(basis_descriptor(basis, c.reshape(-1, 1)))
    sel = rng.choice(len(X), size=300, replace=False)
    Rs = np.concatenate([R[off[i]:off[i + 1]] for i in sel])
    offs = np.concatenate(([0], np.cumsum([off[i + 1] - off[i] for i in sel])))
    Eo, Go = o.energy_forces(Rs, offs)
    assert relerr(E[sel], Eo) < TOL
    assert relerr(np.concatenate([G[off[i]:off[i + 1]] for i in sel]), Go) < TOL

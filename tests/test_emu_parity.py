"""Kernel index logic on the CPU: the CUDA sources compiled against tests/emu/cuda_emu.h (one std::thread
per CUDA thread) must reproduce the oracle.  This is NOT a product path: the emulated library lives in
tests/emu/, is built by tests/emu/build_emu.sh and is loaded only here."""
import ctypes
import os
import subprocess

import pytest

from ace_jl_b200 import _lib
from conftest import ROOT, make_basis
from parity_common import compare_all

EMU_SO = os.path.join(ROOT, "tests", "emu", "libaceb200_emu.so")


@pytest.fixture(scope="module")
def emu():
    srcs = [os.path.join(ROOT, "ace_jl_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "ace_jl_b200", "csrc"))
            if f.endswith((".cu", ".cuh", ".h"))] + [os.path.join(ROOT, "tests", "emu", "cuda_emu.h")]
    if not os.path.exists(EMU_SO) or os.path.getmtime(EMU_SO) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call([os.path.join(ROOT, "tests", "emu", "build_emu.sh")])
    lib = ctypes.CDLL(EMU_SO)
    for name, res, args in _lib.SYMBOLS:
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    saved = _lib._lib
    _lib._lib = lib
    yield lib
    _lib._lib = saved


@pytest.mark.parametrize("kind,nprop,Js", [
    ("inv_simple_3_6", 1, [3, 10, 1, 35, 7, 2, 140]),     # ragged, one tile boundary crossing (J > 128)
    ("inv_sparse_4_8", 4, [5, 12, 33]),                   # order 4, multi-property
    ("euclvec_3_5", 1, [3, 10, 1, 35]),                   # complex effective coefficients, 3 components
    ("euclmat_2_5", 2, [3, 10]),                          # 9 components x 2 properties
    ("species_3_5", 2, [5, 12, 33, 130]),                 # categorical component
    ("inv_morse_2_6", 1, [6, 20]),
    ("inv_agnesi_2_6", 1, [6, 20]),
    ("inv_highL_2_12", 1, [5, 17]),
    ("inv_sparse_4_8", 1, [7, 21]),                       # order 4 through the grouped single-channel stream
    ("inv_complexB_2_5", 1, [6, 20]),                     # complex B / dB outputs (symreal = false)
])
def test_emulated_kernels_match_oracle(emu, kind, nprop, Js):
    compare_all(make_basis(kind), nprop, Js)


def test_emulated_errors(emu):
    import numpy as np
    import ace_jl_b200 as ace
    basis = make_basis("inv_simple_3_6")
    model = ace.LinearACEModel(basis, np.zeros(len(basis)))
    h = model.evaluator.handle
    R = np.array([[0.5, 0.5, 0.5], [1.0, 0.2, 0.1]])
    with pytest.raises(_lib.AceB200Error) as ei:            # an empty environment (product_1pbasis.jl:124)
        h.energy(ace.B200Batch(R, [0, 2, 2]))
    assert ei.value.code == -5
    sp_basis = make_basis("species_3_5")
    hs = ace.LinearACEModel(sp_basis, np.zeros(len(sp_basis))).evaluator.handle
    with pytest.raises(_lib.AceB200Error) as ei:            # unknown category (discrete1pbasis.jl:39)
        hs.energy(ace.B200Batch(R, [0, 2], np.array([1, 9], dtype=np.int32)))
    assert ei.value.code == -6
    with pytest.raises(_lib.AceB200Error):
        model.evaluate([])


def test_emulated_structure_path(emu, monkeypatch):
    """aceb200_structure_energy_forces (pair gather, chunking, force assembly with the caller's and the device-found reverse table, virial) vs the oracle's
    JuLIP-style loop, on the emulated kernels."""
    import numpy as np
    import ace_jl_b200 as ace
    from ace_jl_b200.descriptor import basis_descriptor
    from ace_jl_b200.structure import B200Structure, neighbourlist
    from ace_jl_b200.utils import philox
    from conftest import relerr
    from oracle import Oracle
    from test_structure import RCUT, periodic_crystal
    for kind, nprop in (("inv_simple_3_6", 2), ("species_3_5", 1)):
        basis = make_basis(kind)
        rng = philox(4)
        c = rng.random((len(basis), nprop)) - 0.5
        model = ace.LinearACEModel(basis, c if nprop > 1 else c[:, 0])
        X, cell = periodic_crystal(rng)
        first, nbr, image, rev = neighbourlist(X, RCUT, cell, (True, True, True))
        species = rng.integers(1, 5, len(X)).astype(np.int32) if kind.startswith("species") else None
        Eo, Fo, Wo = Oracle(basis_descriptor(basis, c)).structure_energy_forces(X, first, nbr, image, cell, species)
        for mb, r in ((None, rev), ("0.001", rev), ("0.001", None)):
            if mb:
                monkeypatch.setenv("ACEB200_STRUCT_MB", mb)
            E, F, W = model.evaluator.handle.structure_energy_forces(B200Structure(X, first, nbr, image, cell, species, r))
            assert relerr(E, Eo) < 1e-12 and relerr(F, Fo) < 1e-12 and relerr(W, Wo) < 1e-12
        # a neighbour list that is NOT sorted within a centre: the device-side reverse search falls back to a scan
        perm = np.concatenate([first[i] + rng.permutation(first[i + 1] - first[i]) for i in range(len(X))])
        E, F, W = model.evaluator.handle.structure_energy_forces(B200Structure(X, first, nbr[perm], image[perm], cell, species))
        if np.abs(image).max() <= 1:        # packed neighbour words: the image shift rides in the upper bits of nbr
            from ace_jl_b200.structure import pack_neighbours
            Ek, Fk, Wk = model.evaluator.handle.structure_energy_forces(
                B200Structure(X, first, pack_neighbours(nbr, image)[perm], None, cell, species, None, packed=True))
            assert relerr(Ek, Eo) < 1e-12 and relerr(Fk, Fo) < 1e-12 and relerr(Wk, Wo) < 1e-12
        assert relerr(E, Eo) < 1e-12 and relerr(F, Fo) < 1e-12 and relerr(W, Wo) < 1e-12


def test_emulated_structure_self_images(emu):
    """A cell smaller than the cutoff: pairs with i == j (an atom and its own periodic image) in the reverse search."""
    import numpy as np
    import ace_jl_b200 as ace
    from ace_jl_b200.descriptor import basis_descriptor
    from ace_jl_b200.structure import B200Structure, neighbourlist
    from ace_jl_b200.utils import philox
    from conftest import relerr
    from oracle import Oracle
    from test_structure import RCUT, tiny_cell
    basis = make_basis("inv_simple_3_6")
    rng = philox(23)
    c = rng.random((len(basis), 1)) - 0.5
    model = ace.LinearACEModel(basis, c[:, 0])
    X, cell = tiny_cell(rng)
    first, nbr, image, rev = neighbourlist(X, RCUT, cell, (True, True, True))
    Eo, Fo, Wo = Oracle(basis_descriptor(basis, c)).structure_energy_forces(X, first, nbr, image, cell)
    for r in (rev, None):
        E, F, W = model.evaluator.handle.structure_energy_forces(B200Structure(X, first, nbr, image, cell, None, r))
        assert relerr(E, Eo) < 1e-12 and relerr(F, Fo) < 1e-10 and relerr(W, Wo) < 1e-12


def test_emulated_many_chunks(emu, monkeypatch):
    """A host batch of 700 small ragged environments in chunks of 64: every pipeline lane is reused several times."""
    import numpy as np
    monkeypatch.setenv("ACEB200_CHUNK_ENVS", "64")
    rng = np.random.default_rng(5)
    compare_all(make_basis("inv_simple_3_6"), 1, [int(j) for j in rng.integers(1, 7, 700)], seed=23, jacobians=False)


def test_emulated_multi_device_sharding(emu):
    """aceb200_set_devices: a HOST batch cut into per-device shards (two pretend devices in the emulation) gives bit for
    bit the single-device results, for environment-indexed (E, B) and neighbour-indexed (G, dB) outputs, ragged batches,
    and errors raised inside a shard reach the caller."""
    import numpy as np
    import ace_jl_b200 as ace
    from ace_jl_b200.utils import philox, rand_envs
    from conftest import rn_of
    basis = make_basis("inv_simple_3_6")
    rng = philox(41)
    c = rng.random((len(basis), 2)) - 0.5
    h = ace.LinearACEModel(basis, c).evaluator.handle
    Js = [3, 9, 1, 14, 7, 2, 5, 11, 6]
    R, off, _ = rand_envs(rng, rn_of(basis), len(Js), Js)
    b = ace.B200Batch(R, off)
    E1, G1 = h.energy_forces(b)
    B1, dB1 = h.eval_dB(b)
    h.set_devices([0, 1])
    E2, G2 = h.energy_forces(b)
    B2, dB2 = h.eval_dB(b)
    assert np.array_equal(E1, E2) and np.array_equal(G1, G2) and np.array_equal(B1, B2) and np.array_equal(dB1, dB2)
    c2 = rng.random((len(basis), 2)) - 0.5
    h.set_params(c2)                                    # reaches the replica too
    h1 = ace.LinearACEModel(basis, c2).evaluator.handle
    assert np.array_equal(h.energy(b), h1.energy(b))
    off_bad = off.copy()
    off_bad[-2] = off_bad[-1]                           # the last environment (second shard) is empty
    with pytest.raises(_lib.AceB200Error) as ei:
        h.energy(ace.B200Batch(R, off_bad))
    assert ei.value.code == -5
    with pytest.raises(_lib.AceB200Error):
        h.set_devices([1])                              # must contain the model's own device
    h.set_devices([0])
    assert np.array_equal(h.energy(b), h1.energy(b))

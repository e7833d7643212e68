"""Model (de)serialisation, tested the way the reference tests it: ACEbase.Testing.test_fio
is `read_dict(write_dict(x)) == x` through a dictionary and through a JSON file (test_pibasis.jl:72-73,
test_1pbasis.jl:69, test_euclvec.jl:49, test_EuclideanMatrix.jl:47, test_transforms.jl:79).  Equality follows the
reference's `==` definitions (symmbasis.jl:45-48, product_1pbasis.jl:53-55, b1pcomponent.jl:152-156,
orthpolys.jl:112): tables, flags and component parameters.  On the GPU the re-loaded model must evaluate
bit-identically to the original.
"""
import json

import numpy as np
import pytest

import ace_jl_b200 as ace
from ace_jl_b200 import fio
from ace_jl_b200.descriptor import basis_descriptor
from ace_jl_b200.utils import philox, rand_envs
from conftest import make_basis, nspecies_of, rn_of

KINDS = ["inv_simple_3_6", "inv_sparse_4_8", "euclvec_3_5", "euclmat_2_5", "species_3_5", "inv_complexB_2_5",
         "inv_morse_2_6", "inv_agnesi_2_6"]


def same_1p(a, b):
    assert len(a.bases) == len(b.bases) and np.array_equal(a.indices, b.indices) and a.spec == b.spec
    assert a.symbols == b.symbols
    for x, y in zip(a.bases, b.bases):
        assert type(x) is type(y) and x.spec == y.spec and list(x.degrees) == list(y.degrees) and x.label == y.label
        assert x.symbols == y.symbols and x.varsym == y.varsym
        if isinstance(x, ace.Rn1pBasis):
            assert x.trans == y.trans and x.meta == y.meta
            for k in ("pl", "tl", "pr", "tr"):
                assert getattr(x.R, k) == getattr(y.R, k)
            for k in ("A", "B", "C", "tdf", "ww"):
                assert np.array_equal(getattr(x.R, k), getattr(y.R, k))
        if isinstance(x, ace.Categorical1pBasis):
            assert x.categories == y.categories


def same_pibasis(a, b):
    same_1p(a.basis1p, b.basis1p)
    assert a.real == b.real and np.array_equal(a.spec.orders, b.spec.orders) and np.array_equal(a.spec.iAA2iA, b.spec.iAA2iA)


def same_basis(a, b):
    same_pibasis(a.pibasis, b.pibasis)
    assert type(a.phi) is type(b.phi) and a.real == b.real and type(a.symgrp) is type(b.symgrp)
    A, B = a.A2Bmap, b.A2Bmap
    assert A.shape == B.shape and np.array_equal(A.colptr, B.colptr) and np.array_equal(A.rowval, B.rowval)
    assert np.array_equal(A.nzval, B.nzval)          # bit for bit: JSON round-trips float64 exactly (repr)


def through_json(D):
    return json.loads(json.dumps(D))


@pytest.mark.parametrize("kind", KINDS)
def test_fio_basis_round_trip(kind, zoo, tmp_path):
    basis = zoo(kind)
    same_basis(basis, fio.read_dict(fio.write_dict(basis)))
    same_basis(basis, fio.read_dict(through_json(fio.write_dict(basis))))
    f = tmp_path / "basis.json"
    fio.save_json(str(f), fio.write_dict(basis))
    same_basis(basis, fio.read_dict(fio.load_json(str(f))))
    # the component layers on their own, as the reference tests them
    same_pibasis(basis.pibasis, fio.read_dict(through_json(fio.write_dict(basis.pibasis))))
    same_1p(basis.pibasis.basis1p, fio.read_dict(through_json(fio.write_dict(basis.pibasis.basis1p))))


def test_fio_device_descriptor_is_unchanged(zoo):
    """What reaches the GPU (include/aceb200.h: aceb200_desc) is identical for the re-loaded basis."""
    for kind in ("inv_simple_3_6", "species_3_5", "euclvec_3_5"):
        basis = zoo(kind)
        c = philox(3).random((len(basis), 2)) - 0.5
        d0 = basis_descriptor(basis, c)
        d1 = basis_descriptor(fio.read_dict(through_json(fio.write_dict(basis))), c)
        assert d0.kw.keys() == d1.kw.keys()
        for k in d0.kw:
            assert np.array_equal(np.asarray(d0.kw[k]), np.asarray(d1.kw[k])), k


def test_fio_transforms_and_tags():
    for t in (ace.polytransform(2, 1.0), ace.idtransform(), ace.morsetransform(1.3, 1.1), ace.agnesitransform(1.0, 3)):
        D = through_json(fio.write_dict(t))
        assert D["__id__"] == "ACE_Lambda" and fio.read_dict(D) == t          # test_transforms.jl:79
    D = fio.write_dict(make_basis("inv_simple_3_6"))
    assert D["__id__"] == "ACE_SymmetricBasis" and set(D) == {"__id__", "pibasis", "A2Bmap", "symgrp", "isreal"}
    assert set(D["pibasis"]) == {"__id__", "basis1p", "spec", "real"}
    assert set(D["pibasis"]["spec"]) == {"__id__", "orders", "iAA2iA"}
    assert set(D["pibasis"]["basis1p"]) == {"__id__", "bases", "indices"}
    Rn = D["pibasis"]["basis1p"]["bases"][0]
    assert set(Rn) == {"__id__", "syms", "basis", "fval", "spec", "degrees", "label"}
    assert [F["__id__"] for F in Rn["basis"]["F"]] == ["ACE_Lambda", "ACE_Lambda", "ACE_OrthPolyBasis"]
    assert set(Rn["basis"]["F"][2]) == {"__id__", "T", "pr", "tr", "pl", "tl", "A", "B", "C", "tdf", "ww"}


def test_fio_reader_accepts_triplet_sparse_and_column_lists(zoo):
    """ACEbase spellings the reader tolerates: (I, J, V) triplets and iAA2iA as a JSON list of columns."""
    basis = zoo("inv_simple_3_6")
    D = through_json(fio.write_dict(basis))
    M = D["A2Bmap"]
    colptr = np.asarray(M.pop("colptr"))
    J = np.repeat(np.arange(1, len(colptr)), np.diff(colptr))
    M["I"], M["J"], M["V"] = M.pop("rowval"), J.tolist(), M.pop("nzval")
    D["pibasis"]["spec"]["iAA2iA"] = basis.pibasis.spec.iAA2iA.T.tolist()
    same_basis(basis, fio.read_dict(D))


def test_fio_rejects_what_the_gpu_path_cannot_run(zoo):
    D = through_json(fio.write_dict(zoo("inv_simple_3_6")))
    D["pibasis"]["basis1p"]["bases"][0]["basis"]["F"][1]["exstr"] = "r -> exp(-r^2)"
    with pytest.raises(ValueError):
        fio.read_dict(D)
    with pytest.raises(ValueError):
        fio.read_dict({"__id__": "ACE_Trig1pBasis"})


@pytest.mark.gpu
@pytest.mark.parametrize("kind,nprop", [("inv_simple_3_6", 1), ("species_3_5", 4), ("euclvec_3_5", 1)])
def test_fio_model_evaluates_identically(kind, nprop, tmp_path):
    basis = make_basis(kind)
    rng = philox(11)
    c = rng.random((len(basis), nprop)) - 0.5
    model = ace.LinearACEModel(basis, c if nprop > 1 else c[:, 0])
    f = tmp_path / "model.json"
    fio.save_model(str(f), model)
    model2 = fio.load_model(str(f))
    assert np.array_equal(model.c, model2.c)
    R, off, sp = rand_envs(rng, rn_of(basis), 9, [3, 20, 1, 33, 7, 40, 2, 5, 11], nspecies_of(basis))
    b = ace.B200Batch(R, off, sp)
    E1, G1 = model.evaluator.handle.energy_forces(b)
    E2, G2 = model2.evaluator.handle.energy_forces(b)
    assert np.array_equal(E1, E2) and np.array_equal(G1, G2)
    assert np.array_equal(model.evaluator.handle.eval_B(b), model2.evaluator.handle.eval_B(b))

"""Model (de)serialisation in the layout of ACE.jl's ``write_dict`` / ``read_dict``.

A model fitted in Julia is exported with ``ACE.save_json(fname, write_dict(model))`` and loaded here with
``load_json`` -- no Julia at run time.  Every dictionary carries the reference's ``"__id__"`` tag and the
reference's keys:

    ACE_LinearACEModel   src/linearmodel.jl:80-93          basis, c, evaluator
    ACE_SymmetricBasis   src/symmbasis.jl:50-61            pibasis, A2Bmap, symgrp, isreal
    ACE_PIBasis          src/pibasis.jl:235-243            basis1p, spec, real
    ACE_PIBasisSpec      src/pibasis.jl:245-251            orders, iAA2iA
    ACE_Product1pBasis   src/product_1pbasis.jl:57-66      bases, indices
    ACE_B1pComponent     src/b1pcomponent.jl:159-177       syms, basis, fval, spec, degrees, label
    ACE_SChain           src/chain.jl:75-81                F   (Rn = chain(norm, trans, OrthPolyBasis), Rn.jl:22)
    ACE_OrthPolyBasis    src/polynomials/orthpolys.jl:119-140
    ACE_SHBasis          src/polynomials/sphericalharmonics.jl:318-324
    ACE_Lambda           src/transforms/lambdas.jl:46-51   exstr
    ACE_StaticGet        src/transforms/statetransforms.jl:85-88
    ACE_Categorical1pBasis / ACE_SList   src/discrete1pbasis.jl:42-52, 134-144
    ACE_O3               src/symmetrygroups.jl:82-89
    ACE_Invariant / ACE_EuclideanVector / ACE_EuclideanMatrix / ACE_SymmetricEuclideanMatrix
                         src/properties.jl:106-112, 207-213, 316-327, 345-357
    ACE_ProductEvaluator src/evaluator.jl:22-24

The encodings of *leaves* (number types, dense arrays, ``SparseMatrixCSC``) live in ACEbase 0.2.4, which is not
vendored under /root/reference; they are restated here from that package's FIO conventions and the reader
accepts the spellings it has used (``colptr/rowval/nzval`` and ``I/J/V``; plain lists and tagged arrays;
``real``/``imag`` splits for complex data).  Like the reference's own ``test_fio`` (ACEbase.Testing), the
tests pin this module by round trips: ``read_dict(write_dict(x)) == x`` through a dictionary and through a
JSON file, and by bit-identical evaluation of the re-loaded model on the GPU.
"""
from __future__ import annotations

import json
import re
from typing import Any, Dict

import numpy as np

from .onepbasis import Categorical1pBasis, Product1pBasis, Rn1pBasis, Ylm1pBasis
from .orthpolys import OrthPolyBasis
from .pibasis import PIBasis, PIBasisSpec
from .properties import EuclideanMatrix, EuclideanVector, Invariant, SymmetricEuclideanMatrix
from .symmbasis import SparseCSC, SymmetricBasis
from .symmetrygroups import NoSym, O3
from .transforms import Lambda, parse_exstr

_PROPS = {"ACE_Invariant": Invariant, "ACE_EuclideanVector": EuclideanVector,
          "ACE_EuclideanMatrix": EuclideanMatrix, "ACE_SymmetricEuclideanMatrix": SymmetricEuclideanMatrix}
_PROP_ID = {v: k for k, v in _PROPS.items()}


# ------------------------------------------------------------------------------------------------ leaves
def _write_type(name: str) -> Dict[str, Any]:
    return {"__id__": "Type", "T": name}


def _read_type(D) -> str:
    return D if isinstance(D, str) else D["T"]


def _write_matrix(A: np.ndarray) -> Dict[str, Any]:
    """Dense real matrix: column-major ``vals`` like Julia's ``A[:]``."""
    A = np.asarray(A)
    T = "Int64" if np.issubdtype(A.dtype, np.integer) else "Float64"
    return {"__id__": "ACE_ArrayOfNumber", "T": _write_type(T), "size": list(A.shape),
            "vals": A.reshape(-1, order="F").tolist()}


def _read_array(D) -> np.ndarray:
    """Tagged array (column-major ``vals`` + ``size`` / ``nrows, ncols``), list of columns, or plain list."""
    if isinstance(D, dict):
        if "real" in D and "imag" in D:
            return _read_array(D["real"]) + 1j * _read_array(D["imag"])
        vals = np.asarray(D.get("vals", D.get("data")))
        if "size" in D:
            return vals.reshape(tuple(D["size"]), order="F")
        if "nrows" in D:
            return vals.reshape((D["nrows"], D["ncols"]), order="F")
        return vals
    return np.asarray(D)


def _prop_value(phi_cls, D: Dict[str, Any]) -> np.ndarray:
    """Complex components of one serialised property value, in the device order of ``SparseCSC.nzval``:
    Invariant 1; EuclideanVector 3; matrices 9, column-major (properties.jl:310-313 fills `for i for j` into an
    SMatrix, whose memory is column-major)."""
    if phi_cls is Invariant:
        v = D["val"]
        return np.array([complex(v["re"], v["im"]) if isinstance(v, dict) else v], dtype=np.complex128)
    if phi_cls is EuclideanVector:
        return _read_array(D["val"]).astype(np.complex128).reshape(3)
    M = _read_array(D["valr"]).astype(np.float64).reshape(3, 3) + 1j * _read_array(D["vali"]).astype(np.float64).reshape(3, 3)
    return M.reshape(-1, order="F")


def _write_prop_value(phi, v: np.ndarray) -> Dict[str, Any]:
    cls = type(phi)
    if cls is Invariant:
        val = float(v[0].real) if v[0].imag == 0.0 else {"re": float(v[0].real), "im": float(v[0].imag)}
        return {"__id__": "ACE_Invariant", "val": val, "T": _write_type("Float64")}
    if cls is EuclideanVector:
        return {"__id__": "ACE_EuclideanVector",
                "val": {"__id__": "ACE_ArrayOfNumber", "T": _write_type("ComplexF64"), "size": [3],
                        "real": v.real.tolist(), "imag": v.imag.tolist()}}
    M = v.reshape(3, 3, order="F")
    return {"__id__": _PROP_ID[cls], "valr": _write_matrix(M.real), "vali": _write_matrix(M.imag),
            "T": _write_type("Float64")}


# ------------------------------------------------------------------------------------------------ writers
def write_dict(obj) -> Dict[str, Any]:
    """``write_dict`` for every type on the evaluation path (see the module docstring for file:line)."""
    from .api import LinearACEModel
    if isinstance(obj, LinearACEModel):
        c = np.asarray(obj.c)
        cD = ({"__id__": "ACE_ArrayOfNumber", "T": _write_type("Float64"), "size": [len(c)], "vals": c.tolist()}
              if c.ndim == 1 else
              # Vector{SVector{N,Float64}}: one inner list per basis function
              {"__id__": "ACE_VectorOfSVector", "T": _write_type("Float64"), "N": int(c.shape[1]), "vals": c.tolist()})
        return {"__id__": "ACE_LinearACEModel", "basis": write_dict(obj.basis), "c": cD,
                "evaluator": {"__id__": "ACE_ProductEvaluator"}}
    if isinstance(obj, SymmetricBasis):
        M = obj.A2Bmap
        return {"__id__": "ACE_SymmetricBasis", "pibasis": write_dict(obj.pibasis),
                "A2Bmap": {"__id__": "SparseMatrixCSC", "TF": _write_type("ACE." + obj.phi.name), "TI": _write_type("Int64"),
                           "m": M.m, "n": M.n, "colptr": M.colptr.tolist(), "rowval": M.rowval.tolist(),
                           "nzval": [_write_prop_value(obj.phi, v) for v in M.nzval]},
                "symgrp": write_dict(obj.symgrp), "isreal": bool(obj.real)}
    if isinstance(obj, PIBasis):
        return {"__id__": "ACE_PIBasis", "basis1p": write_dict(obj.basis1p), "spec": write_dict(obj.spec),
                "real": bool(obj.real)}
    if isinstance(obj, PIBasisSpec):
        return {"__id__": "ACE_PIBasisSpec", "orders": obj.orders.tolist(), "iAA2iA": _write_matrix(obj.iAA2iA)}
    if isinstance(obj, Product1pBasis):
        return {"__id__": "ACE_Product1pBasis", "bases": [write_dict(B) for B in obj.bases],
                "indices": np.asarray(obj.indices).tolist()}
    if isinstance(obj, Rn1pBasis):
        chain = {"__id__": "ACE_SChain", "F": [{"__id__": "ACE_Lambda", "exstr": "rr -> norm(rr)"},
                                               write_dict(obj.trans), write_dict(obj.R)]}
        return _write_component(obj, chain)
    if isinstance(obj, Ylm1pBasis):
        return _write_component(obj, {"__id__": "ACE_SHBasis", "T": _write_type("Float64"), "maxL": obj.L})
    if isinstance(obj, Categorical1pBasis):
        T = "Symbol" if isinstance(obj.categories[0], str) else "Int64"
        return {"__id__": "ACE_Categorical1pBasis",
                "categories": {"__id__": "ACE_SList", "T": _write_type(T), "list": list(obj.categories)},
                "VSYM": obj.varsym, "ISYM": obj.symbols[0], "label": obj.label}
    if isinstance(obj, OrthPolyBasis):
        return {"__id__": "ACE_OrthPolyBasis", "T": _write_type("Float64"), "pr": int(obj.pr), "tr": float(obj.tr),
                "pl": int(obj.pl), "tl": float(obj.tl), "A": np.asarray(obj.A).tolist(), "B": np.asarray(obj.B).tolist(),
                "C": np.asarray(obj.C).tolist(), "tdf": np.asarray(obj.tdf).tolist(), "ww": np.asarray(obj.ww).tolist()}
    if isinstance(obj, Lambda):
        return {"__id__": "ACE_Lambda", "exstr": obj.exstr}
    if isinstance(obj, O3):
        return {"__id__": "ACE_O3", "lsym": obj.lsym, "msym": obj.msym}
    if isinstance(obj, NoSym):
        return {"__id__": "ACE_NoSym"}
    raise TypeError(f"write_dict: no serialisation for {type(obj).__name__}")


def _write_component(B, inner) -> Dict[str, Any]:
    """b1pcomponent.jl:159-169."""
    return {"__id__": "ACE_B1pComponent", "syms": list(B.symbols), "basis": inner,
            "fval": {"__id__": "ACE_StaticGet", "expr": f"ACE.Transforms.GetVal{{:{B.varsym}}}"},
            "spec": [dict(zip(B.symbols, (int(x) for x in b))) for b in B.spec],
            "degrees": [int(d) for d in B.degrees], "label": B.label}


# ------------------------------------------------------------------------------------------------ readers
def read_dict(D: Dict[str, Any]):
    """Dispatch on ``D["__id__"]`` like ``read_dict(::Val{:ACE_...}, D)``."""
    tag = D["__id__"]
    if tag == "ACE_LinearACEModel":
        from .api import LinearACEModel
        basis = read_dict(D["basis"])
        ev = D.get("evaluator", {}).get("__id__", "ACE_ProductEvaluator")
        if ev not in ("ACE_ProductEvaluator", "ACE_NaiveEvaluator", "ACE_B200Evaluator"):
            raise ValueError(f"read_dict: unknown evaluator {ev!r}")
        cD = D["c"]
        c = np.asarray(cD["vals"] if isinstance(cD, dict) else cD, dtype=np.float64)
        if c.ndim == 2 and isinstance(cD, dict) and "size" in cD:       # dense matrix written column-major
            c = c.reshape(tuple(cD["size"]), order="F")
        return LinearACEModel(basis, c)
    if tag == "ACE_SymmetricBasis":
        pib = read_dict(D["pibasis"])
        M = D["A2Bmap"]
        nz = M["nzval"] if "nzval" in M else M["V"]
        if not nz:
            raise ValueError("read_dict: empty A2Bmap")
        phi_cls = _PROPS[nz[0]["__id__"]]
        vals = np.stack([_prop_value(phi_cls, v) for v in nz])
        if "colptr" in M:
            A2B = SparseCSC(M["m"], M["n"], M["colptr"], M["rowval"], vals)
        else:
            A2B = SparseCSC.from_triplets(np.asarray(M["I"]), np.asarray(M["J"]), vals, M["m"], M["n"], vals.shape[1])
        return SymmetricBasis.from_parts(phi_cls(), pib, A2B, read_dict(D["symgrp"]), bool(D["isreal"]))
    if tag == "ACE_PIBasis":
        return PIBasis(read_dict(D["basis1p"]), read_dict(D["spec"]), isreal=bool(D["real"]))
    if tag == "ACE_PIBasisSpec":
        tab = _read_array(D["iAA2iA"])
        if tab.ndim == 2 and tab.shape[0] != len(D["orders"]):           # JSON list of columns
            tab = tab.T
        return PIBasisSpec(np.asarray(D["orders"]), tab)
    if tag == "ACE_Product1pBasis":
        bases = [read_dict(b) for b in D["bases"]]
        B = Product1pBasis(bases, indices=np.asarray(D["indices"], dtype=np.int32).reshape(-1, len(bases)))
        # the spec is implied by the indices (product_1pbasis.jl:283-304)
        B.spec = [_spec_of(B, row) for row in B.indices]
        return B
    if tag == "ACE_B1pComponent":
        return _read_component(D)
    if tag == "ACE_Categorical1pBasis":
        L = D["categories"]
        cats = L["list"] if isinstance(L, dict) else L
        return Categorical1pBasis(list(cats), varsym=D["VSYM"], idxsym=D["ISYM"], label=D["label"])
    if tag == "ACE_OrthPolyBasis":
        f = lambda k: np.asarray(D[k], dtype=np.float64)   # noqa: E731
        return OrthPolyBasis(int(D["pl"]), float(D["tl"]), int(D["pr"]), float(D["tr"]), f("A"), f("B"), f("C"),
                             f("tdf"), f("ww"))
    if tag == "ACE_Lambda":
        return parse_exstr(D["exstr"])
    if tag == "ACE_O3":
        return O3(D["lsym"], D["msym"])
    if tag == "ACE_NoSym":
        return NoSym()
    raise ValueError(f"read_dict: unknown __id__ {tag!r}")


def _spec_of(B: Product1pBasis, row) -> tuple:
    b = [None] * len(B.symbols)
    for ib, comp in enumerate(B.bases):
        sub = comp.spec[int(row[ib]) - 1]
        for k, s in enumerate(comp.symbols):
            b[B.symbols.index(s)] = sub[k]
    return tuple(b)


_GET = re.compile(r"Get(?:Val|Norm)\{:(\w+)\}")


def _read_component(D):
    """b1pcomponent.jl:172-178.  The inner basis decides which of the supported components this is; anything
    else (Scal1pBasis, Trig1pBasis, multipliers) is not on the B200 path and is refused loudly."""
    inner, syms = D["basis"], list(D["syms"])
    m = _GET.search(D["fval"].get("expr", ""))
    varsym = m.group(1) if m else "rr"
    spec = [tuple(int(b[s]) for s in syms) for b in D["spec"]]
    if inner["__id__"] == "ACE_SHBasis":
        B = Ylm1pBasis(int(inner["maxL"]), varsym=varsym, lsym=syms[0], msym=syms[1], label=D["label"])
    elif inner["__id__"] == "ACE_SChain":
        F = inner["F"]
        if len(F) != 3 or F[0].get("exstr", "").replace(" ", "") != "rr->norm(rr)" or F[2]["__id__"] != "ACE_OrthPolyBasis":
            raise ValueError("read_dict: only chain(norm, transform, OrthPolyBasis) radial components are supported")
        B = Rn1pBasis(read_dict(F[2]), read_dict(F[1]), varsym=varsym, nsym=syms[0], label=D["label"])
    else:
        raise ValueError(f"read_dict: unsupported B1pComponent basis {inner['__id__']!r}")
    if spec != B.spec or [int(d) for d in D["degrees"]] != list(B.degrees):
        B.spec, B.degrees = spec, [int(d) for d in D["degrees"]]
        B._build_inv()
    return B


# ------------------------------------------------------------------------------------------------ files
def save_json(fname: str, D: Dict[str, Any]) -> None:
    """ACEbase.FIO.save_json / save_dict."""
    with open(fname, "w") as f:
        json.dump(D, f)


def load_json(fname: str) -> Dict[str, Any]:
    with open(fname) as f:
        return json.load(f)


def save_model(fname: str, model) -> None:
    save_json(fname, write_dict(model))


def load_model(fname: str):
    """``read_dict(load_json(fname))``: a LinearACEModel (or basis) ready to evaluate on the GPU."""
    return read_dict(load_json(fname))

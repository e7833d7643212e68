"""Reader / writer of the "aceb200-golden-v1" JSON schema that tests/golden/export_golden.jl emits from a real
ACE.jl installation.

`load(path)` turns one file into the descriptor the C ABI takes -- straight from the dumped tables, exactly what
the Julia shim of INTEGRATION.md does from the live objects; nothing is re-derived in Python, so the comparison
does not depend on the SVD gauge or on this repository's reading of the construction code.

`dump_from_mirror(...)` writes the SAME schema from the Python mirror + the CPU oracle.  Such a file pins
nothing about the reference (its "generator" field says so); it exists so that the loader and the comparison
code are exercised in a container without Julia.
"""
import json
import os

import numpy as np

from ace_jl_b200._lib import DescHolder
from ace_jl_b200.transforms import parse_exstr

SCHEMA = "aceb200-golden-v1"
KIND = {"Rn": 0, "Ylm": 1, "Cat": 2}
HERE = os.path.dirname(os.path.abspath(__file__))
JULIA_DIR = os.path.join(HERE, "golden", "julia")


def _c(a, shape):
    """interleaved (re, im) list -> complex array of `shape`."""
    a = np.asarray(a, dtype=np.float64).reshape(-1, 2)
    return (a[:, 0] + 1j * a[:, 1]).reshape(shape)


def load(path):
    D = json.load(open(path))
    assert D["schema"] == SCHEMA, D.get("schema")
    rn = D["rn"]
    trans = parse_exstr(rn["trans_exstr"])
    orders = np.asarray(D["orders"], dtype=np.int32)
    iAA2iA = np.asarray(D["iAA2iA"], dtype=np.int32).reshape(len(orders), -1)
    A2B = D["A2B"]
    I, J = np.asarray(A2B["I"], dtype=np.int64), np.asarray(A2B["J"], dtype=np.int64)
    V = np.asarray(A2B["V"], dtype=np.float64).reshape(len(I), -1)          # [nnz][ncomp * 2]
    ncomp = V.shape[1] // 2 if len(I) else 1
    # findnz of a SparseMatrixCSC is column-major sorted: rebuild colptr
    order = np.lexsort((I, J))
    I, J, V = I[order], J[order], V[order]
    colptr = np.concatenate([[1], 1 + np.cumsum(np.bincount(J - 1, minlength=A2B["n"]))]).astype(np.int32)
    c = np.asarray(D["c"], dtype=np.float64).reshape(A2B["m"], -1)
    cats = D.get("categories")
    holder = DescHolder(
        n_rad=len(rn["A"]), pl=rn["pl"], pr=rn["pr"], tl=rn["tl"], tr=rn["tr"],
        rad_A=rn["A"], rad_B=rn["B"], rad_C=rn["C"], trans_kind=trans.kind, trans_par=trans.c_params(),
        maxL=D["maxL"], n_cat=len(cats) if cats else 0,
        n_comp=len(D["comp_kinds"]), comp_kind=[KIND[k] for k in D["comp_kinds"]],
        nA=len(D["indices"]), indices=np.asarray(D["indices"], dtype=np.int32),
        nAA=len(orders), maxord=iAA2iA.shape[1], orders=orders,
        iAA2iA=np.asfortranarray(iAA2iA).ravel(order="F"),
        pireal=int(D["pireal"]), symreal=int(D["symreal"]),
        nB=A2B["m"], ncomp=ncomp, nnz=len(I), colptr=colptr, rowval=I.astype(np.int32), nzval=V,
        nprop=c.shape[1], c=c)
    nA, nAA, nB, nprop = len(D["indices"]), len(orders), A2B["m"], c.shape[1]
    envs = []
    for e in D["envs"]:
        R = np.asarray(e["R"], dtype=np.float64).reshape(-1, 3)
        Jn = len(R)
        envs.append(dict(
            R=R, species=None if e.get("species") is None else np.asarray(e["species"], dtype=np.int32),
            A=_c(e["A"], (nA,)), AA=_c(e["AA"], (nAA,)), B=_c(e["B"], (nB, ncomp)),
            dA=_c(e["dA"], (Jn, nA, 3)), dAA=_c(e["dAA"], (Jn, nAA, 3)), dB=_c(e["dB"], (Jn, nB, 3, ncomp)),
            E=_c(e["E"], (nprop, ncomp)), G=_c(e["G"], (Jn, nprop, 3, ncomp))))
    ctilde = _c(np.asarray(D["ctilde"], dtype=np.float64).ravel(), (nAA, nprop, ncomp))
    return D, holder, ctilde, envs


def _il(a):
    a = np.asarray(a, dtype=np.complex128).ravel()
    return np.stack([a.real, a.imag], axis=1).ravel().tolist()


def dump_from_mirror(path, basis, c, R_list, species_list=None, name="mirror"):
    """Write the schema from the Python mirror of the construction + the CPU oracle (NOT a reference pin)."""
    from ace_jl_b200.descriptor import basis_descriptor
    from oracle import Oracle
    from conftest import rn_of
    c = np.asarray(c, dtype=np.float64).reshape(len(basis), -1)
    holder = basis_descriptor(basis, c)
    o = Oracle(holder)
    b1p = basis.pibasis.basis1p
    Rn = rn_of(basis)
    A2B = basis.A2Bmap
    Jcol = np.repeat(np.arange(1, A2B.n + 1), np.diff(A2B.colptr))
    kinds = {0: "Rn", 1: "Ylm", 2: "Cat"}
    cat = b1p.component(2)
    D = {
        "schema": SCHEMA, "generator": "python-mirror + CPU oracle (exercises the loader; pins nothing)", "config": name,
        "rn": {"pl": int(Rn.R.pl), "tl": float(Rn.R.tl), "pr": int(Rn.R.pr), "tr": float(Rn.R.tr),
               "A": list(map(float, Rn.R.A)), "B": list(map(float, Rn.R.B)), "C": list(map(float, Rn.R.C)),
               "trans_exstr": Rn.trans.exstr},
        "maxL": int(b1p.component(1).L), "comp_kinds": [kinds[B.kind] for B in b1p.bases],
        "categories": None if cat is None else [str(x) for x in cat.categories],
        "indices": np.asarray(b1p.indices).tolist(),
        "spec1p": [dict(zip(b1p.symbols, map(str, b))) for b in b1p.spec],
        "orders": np.asarray(basis.pibasis.spec.orders).tolist(),
        "iAA2iA": np.asarray(basis.pibasis.spec.iAA2iA).tolist(),
        "pireal": bool(basis.pibasis.real), "symreal": bool(basis.real), "property": type(basis.phi).__name__,
        "A2B": {"m": int(A2B.m), "n": int(A2B.n), "I": np.asarray(A2B.rowval).tolist(), "J": Jcol.tolist(),
                "V": [_il(v) for v in np.asarray(A2B.nzval)]},
        "nprop": int(c.shape[1]), "c": c.tolist(), "ctilde": [_il(v) for v in o.eff_coeffs()], "envs": [],
    }
    for k, R in enumerate(R_list):
        R = np.asarray(R, dtype=np.float64).reshape(-1, 3)
        off = np.array([0, len(R)], dtype=np.int64)
        sp = None if species_list is None else np.asarray(species_list[k], dtype=np.int32)
        A, dA = o.eval_dA(R, off, sp)
        AA, dAA = o.eval_dAA(R, off, sp)
        B, dB = o.eval_dB(R, off, sp)
        env = {"R": R.tolist(), "species": None if sp is None else sp.tolist(),
               "A": _il(A[0]), "AA": _il(AA[0]), "B": _il(B[0]), "dA": _il(dA), "dAA": _il(dAA), "dB": _il(dB)}
        if basis.real:
            E, G = o.energy_forces(R, off, sp)
            env["E"], env["G"] = _il(E[0]), _il(G)
        else:
            env["E"], env["G"] = _il(np.zeros((c.shape[1], A2B.ncomp))), _il(np.zeros((len(R), c.shape[1], 3, A2B.ncomp)))
        D["envs"].append(env)
    with open(path, "w") as f:
        json.dump(D, f)
    return path

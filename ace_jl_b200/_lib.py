"""ctypes binding of the C ABI declared in include/aceb200.h.

This is the Python stand-in for the Julia `ccall` shim shown in INTEGRATION.md: the same structs, the
same entry points.  Loading fails loudly if ``libaceb200.so`` has not been built; there is no CPU
fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libaceb200.so")

ABI_VERSION = 2
HOST, DEVICE = 0, 1
MAX_COMP = 4

ERRORS = {0: "OK", -1: "EDESC", -2: "EUNSUPPORTED", -3: "ECUDA", -4: "ENOMEM", -5: "EEMPTY", -6: "ECATEGORY"}

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)


class Desc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("struct_bytes", C.c_int32),
        ("n_rad", C.c_int32), ("pl", C.c_int32), ("pr", C.c_int32),
        ("tl", C.c_double), ("tr", C.c_double),
        ("rad_A", c_double_p), ("rad_B", c_double_p), ("rad_C", c_double_p),
        ("trans_kind", C.c_int32), ("_pad0", C.c_int32), ("trans_par", C.c_double * 4),
        ("maxL", C.c_int32), ("n_cat", C.c_int32),
        ("n_comp", C.c_int32), ("comp_kind", C.c_int32 * MAX_COMP), ("nA", C.c_int32),
        ("indices", c_int32_p),
        ("nAA", C.c_int32), ("maxord", C.c_int32), ("orders", c_int32_p), ("iAA2iA", c_int32_p),
        ("pireal", C.c_int32), ("symreal", C.c_int32),
        ("nB", C.c_int32), ("ncomp", C.c_int32), ("nnz", C.c_int64),
        ("colptr", c_int32_p), ("rowval", c_int32_p), ("nzval", c_double_p),
        ("nprop", C.c_int32), ("_pad1", C.c_int32), ("c", c_double_p),
    ]


class Batch(C.Structure):
    _fields_ = [
        ("nenv", C.c_int64), ("offsets", c_int64_p), ("R", c_double_p), ("species", c_int32_p),
        ("space", C.c_int32), ("_pad", C.c_int32), ("nJ", C.c_int64),
    ]


class Structure(C.Structure):
    _fields_ = [
        ("natoms", C.c_int64), ("npairs", C.c_int64), ("X", c_double_p), ("first", c_int64_p), ("nbr", c_int32_p),
        ("image", C.POINTER(C.c_int8)), ("species", c_int32_p), ("rev", c_int32_p), ("cell", C.c_double * 9),
        ("space", C.c_int32), ("flags", C.c_int32),
    ]


NBR_PACKED = 1  # aceb200_structure.flags: ACEB200_NBR_PACKED


class Sizes(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("nA", "nAA", "nB", "ncomp", "nprop", "maxord", "pireal", "symreal")]


class AceB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"aceb200 error {ERRORS.get(code, code)}: {msg}")
        self.code = code


_lib: Optional[C.CDLL] = None

# every symbol include/aceb200.h declares: (name, restype, argtypes)
_VOIDPP = C.POINTER(C.c_void_p)
SYMBOLS = [
    ("aceb200_device_count", C.c_int, []),
    ("aceb200_set_device", C.c_int, [C.c_int]),
    ("aceb200_set_devices", C.c_int, [C.c_void_p, C.c_int, c_int32_p]),
    ("aceb200_last_error", C.c_int, [C.c_char_p, C.c_int]),
    ("aceb200_model_create", C.c_int, [C.POINTER(Desc), _VOIDPP]),
    ("aceb200_model_destroy", C.c_int, [C.c_void_p]),
    ("aceb200_set_params", C.c_int, [C.c_void_p, c_double_p, C.c_int64]),
    ("aceb200_get_eff_coeffs", C.c_int, [C.c_void_p, c_double_p]),
    ("aceb200_set_stream", C.c_int, [C.c_void_p, C.c_void_p]),
    ("aceb200_launch_count", C.c_int64, [C.c_void_p]),
    ("aceb200_eval_A", C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_void_p]),
    ("aceb200_eval_AA", C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_void_p]),
    ("aceb200_eval_B", C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_void_p]),
    ("aceb200_eval_dA", C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.c_void_p]),
    ("aceb200_eval_dAA", C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.c_void_p]),
    ("aceb200_eval_dB", C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.c_void_p]),
    ("aceb200_energy", C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_void_p]),
    ("aceb200_energy_forces", C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.c_void_p]),
    ("aceb200_energy_forces_dp", C.c_int, [C.c_void_p, C.POINTER(Batch), c_double_p, C.c_void_p, C.c_void_p]),
    ("aceb200_adjoint_eval_d", C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.c_void_p]),
    ("aceb200_structure_energy_forces", C.c_int, [C.c_void_p, C.POINTER(Structure), C.c_void_p, C.c_void_p, C.c_void_p]),
    ("aceb200_model_sizes", C.c_int, [C.c_void_p, C.POINTER(Sizes)]),
    ("aceb200_last_kernel_ms", C.c_int, [C.c_void_p, c_double_p]),
    ("aceb200_last_stage_ms", C.c_int, [C.c_void_p, c_double_p]),
    ("aceb200_measure_fp64", C.c_int, [c_double_p]),
    ("aceb200_measure_dmma", C.c_int, [c_double_p]),
]


def load() -> C.CDLL:
    """Load libaceb200.so; raise if it is missing (the product has no other path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a).  ace_jl_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    buf = C.create_string_buffer(1024)
    load().aceb200_last_error(buf, 1024)
    return buf.value.decode(errors="replace")


def check(rc: int):
    if rc != 0:
        raise AceB200Error(rc, last_error())


def _p(arr: Optional[np.ndarray], ctype):
    if arr is None:
        return C.cast(None, C.POINTER(ctype))
    return arr.ctypes.data_as(C.POINTER(ctype))


class DescHolder:
    """An ``aceb200_desc`` plus the numpy arrays that back its pointers (kept alive together)."""

    def __init__(self, **kw):
        self.arrays = {}
        d = Desc()
        d.abi_version = ABI_VERSION
        d.struct_bytes = C.sizeof(Desc)

        def arr(name, a, dtype, ctype):
            if a is None:
                setattr(d, name, C.cast(None, C.POINTER(ctype)))
                return
            a = np.ascontiguousarray(a, dtype=dtype)
            self.arrays[name] = a
            setattr(d, name, _p(a, ctype))

        for k in ("n_rad", "pl", "pr", "trans_kind", "maxL", "n_cat", "n_comp", "nA", "nAA", "maxord",
                  "pireal", "symreal", "nB", "ncomp", "nnz", "nprop"):
            setattr(d, k, int(kw[k]))
        d.tl, d.tr = float(kw["tl"]), float(kw["tr"])
        for i, v in enumerate(kw["trans_par"]):
            d.trans_par[i] = float(v)
        for i, v in enumerate(kw["comp_kind"]):
            d.comp_kind[i] = int(v)
        arr("rad_A", kw["rad_A"], np.float64, C.c_double)
        arr("rad_B", kw["rad_B"], np.float64, C.c_double)
        arr("rad_C", kw["rad_C"], np.float64, C.c_double)
        arr("indices", kw["indices"], np.int32, C.c_int32)
        arr("orders", kw["orders"], np.int32, C.c_int32)
        arr("iAA2iA", kw["iAA2iA"], np.int32, C.c_int32)
        arr("colptr", kw["colptr"], np.int32, C.c_int32)
        arr("rowval", kw["rowval"], np.int32, C.c_int32)
        arr("nzval", kw["nzval"], np.float64, C.c_double)
        arr("c", kw.get("c"), np.float64, C.c_double)
        self.desc = d
        self.kw = kw

    def with_c(self, c: Optional[np.ndarray], nprop: int) -> "DescHolder":
        kw = dict(self.kw)
        kw["c"] = c
        kw["nprop"] = nprop
        return DescHolder(**kw)


def make_batch(nenv: int, offsets_ptr: int, R_ptr: int, species_ptr: int, space: int, nJ: int = 0) -> Batch:
    b = Batch()
    b.nenv = int(nenv)
    b.nJ = int(nJ)
    b.offsets = C.cast(C.c_void_p(offsets_ptr), c_int64_p)
    b.R = C.cast(C.c_void_p(R_ptr), c_double_p)
    b.species = C.cast(C.c_void_p(species_ptr or None), c_int32_p)
    b.space = int(space)
    return b

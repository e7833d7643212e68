"""The C-ABI library loads here (no GPU) and exports every symbol include/aceb200.h declares."""
import ctypes
import os
import re

from conftest import ROOT
from ace_jl_b200 import _lib


def _declared():
    txt = open(os.path.join(ROOT, "include", "aceb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(aceb200_[A-Za-z0-9_]+)\s*\(", txt)))


def test_header_symbols_are_exported():
    names = _declared()
    assert len(names) >= 20
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in aceb200.h but not exported"


def test_binding_covers_header():
    assert sorted(n for n, _, _ in _lib.SYMBOLS) == _declared()


def test_struct_sizes_match_header():
    # compile a tiny C program against the header and compare sizeof / offsetof with the ctypes mirror
    import subprocess, tempfile, textwrap
    src = textwrap.dedent("""
        #include <stdio.h>
        #include <stddef.h>
        #include "aceb200.h"
        int main(void) {
            printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(aceb200_desc), offsetof(aceb200_desc, trans_par),
                   offsetof(aceb200_desc, indices), offsetof(aceb200_desc, nzval), sizeof(aceb200_batch), sizeof(aceb200_sizes));
            return 0; }""")
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        out = subprocess.check_output([os.path.join(d, "t")]).split()
    got = [int(x) for x in out]
    D = _lib.Desc
    assert got == [ctypes.sizeof(D), D.trans_par.offset, D.indices.offset, D.nzval.offset,
                   ctypes.sizeof(_lib.Batch), ctypes.sizeof(_lib.Sizes)]


def test_no_gpu_means_loud_failure():
    """Without a device model_create must fail with ECUDA, never fall back to a CPU path."""
    import numpy as np
    import pytest
    import ace_jl_b200 as ace
    lib = _lib.load()
    if lib.aceb200_device_count() > 0:
        pytest.skip("a GPU is present")
    basis = ace.SymmetricBasis(ace.Invariant(), ace.utils.RnYlm_1pbasis(maxdeg=3), ace.SimpleSparseBasis(2, 3))
    with pytest.raises(_lib.AceB200Error) as ei:
        ace.LinearACEModel(basis, np.zeros(len(basis)))
    assert ei.value.code == -3


def test_unsupported_transform_is_rejected():
    import pytest
    from ace_jl_b200.transforms import parse_exstr
    assert parse_exstr("r -> ((1+1.0)/(1+r))^2").kind == 1
    assert parse_exstr("r -> exp(- 1.3 * (r / 1.1 - 1))").params == (1.3, 1.1)
    assert parse_exstr("r -> 1 / (1 + 0.5 * (r / 1.0)^3)").params == (1.0, 3.0, 0.5)
    with pytest.raises(ValueError):
        parse_exstr("r -> sin(r)")

"""Load a tests/golden/*.npz fixture into the descriptor both the oracle and the CUDA library take."""
import os

import numpy as np

from ace_jl_b200._lib import DescHolder

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = ["config1_inv_sparse_3_10", "config2_inv_sparse_3_12", "config4_euclvec_3_5", "config5_species_3_5"]


def load(name):
    z = np.load(os.path.join(HERE, "golden", name + ".npz"))
    kw = {}
    for k in z.files:
        if k.startswith("tab_"):
            kw[k[4:]] = z[k]
        elif k.startswith("par_"):
            v = z[k]
            kw[k[4:]] = v.tolist() if v.ndim else v.item()
    kw.setdefault("c", z["c"])
    return z, DescHolder(**kw)

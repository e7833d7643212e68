"""Regenerate tests/golden/*.npz.

No golden vectors for this path exist in the reference and the reference itself (Julia)
cannot run in the build container, so these fixtures are ORACLE-generated: seeded inputs, the tables the
model was built from, and the oracle's outputs.  They pin (i) the oracle against regressions and (ii) the
CUDA path on the GPU box, where neither /root/reference nor a rebuild of the tables is needed.
When a Julia + ACE.jl install becomes available, SURVEY.md Appendix C.7 describes the dump that should
replace them.

Usage:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

from ace_jl_b200.descriptor import basis_descriptor  # noqa: E402
from ace_jl_b200.utils import philox, rand_envs  # noqa: E402
from conftest import make_basis, nspecies_of, rn_of  # noqa: E402
from oracle import Oracle  # noqa: E402

CASES = {
    # name: (zoo kind, nprop, neighbour counts, seed)
    "config1_inv_sparse_3_10": ("inv_sparse_3_10", 1, [30, 30, 30], 20241),
    "config2_inv_sparse_3_12": ("inv_sparse_3_12", 1, [40, 40, 40], 20242),
    "config4_euclvec_3_5": ("euclvec_3_5", 1, [30, 30], 20244),
    "config5_species_3_5": ("species_3_5", 4, [40, 40], 20245),
}


def main():
    for name, (kind, nprop, Js, seed) in CASES.items():
        basis = make_basis(kind)
        rng = philox(seed)
        R, off, sp = rand_envs(rng, rn_of(basis), len(Js), Js, nspecies_of(basis))
        c = philox(seed + 1000).random((len(basis), nprop)) - 0.5
        holder = basis_descriptor(basis, c)
        o = Oracle(holder)
        E, G = o.energy_forces(R, off, sp)
        B, dB = o.eval_dB(R, off, sp)
        out = dict(R=R, offsets=off, c=c, E=E, G=G, A=o.eval_A(R, off, sp), AA=o.eval_AA(R, off, sp), B=B,
                   dB_checksum=np.array([np.abs(dB).sum(), np.sqrt((dB ** 2).sum()), np.abs(dB).max()]),
                   ctilde=o.eff_coeffs())
        if sp is not None:
            out["species"] = sp
        # the tables, so that the GPU test does not depend on re-deriving them (SVD gauge!)
        for k, v in holder.arrays.items():
            out["tab_" + k] = v
        for k in ("n_rad", "pl", "pr", "tl", "tr", "trans_kind", "maxL", "n_cat", "n_comp", "nA", "nAA", "maxord",
                  "pireal", "symreal", "nB", "ncomp", "nnz", "nprop"):
            out["par_" + k] = np.array(holder.kw[k])
        out["par_trans_par"] = np.array(holder.kw["trans_par"], dtype=np.float64)
        out["par_comp_kind"] = np.array(holder.kw["comp_kind"], dtype=np.int32)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()

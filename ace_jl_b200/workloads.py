"""The five BASELINE.json configurations as (basis builder, call, sizes, algorithmic work) records.

Shared by bench.py (`--config N`), the full-size `-m gpu` parity tests and benchmarks/run_configs.py, so that
"config 3" means the same tables everywhere.  Sources of the workload definitions:
benchmark/bm_basis.jl:56-71 (config 1), benchmark/bm_linear.jl:68-96,149-155 (configs 2, 5),
profile/profile_linearmodel.jl:13-25 (config 3), test/test_euclvec.jl / test/test_EuclideanMatrix.jl (config 4).
Synthetic inputs follow src/utils/random.jl:22-25 (radius uniform in [rin, rcut], direction uniform).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Dict

import numpy as np

from .onepbasis import Categorical1pBasis, Product1pBasis
from .properties import EuclideanMatrix, EuclideanVector, Invariant
from .selectors import SparseBasis
from .symmbasis import SymmetricBasis
from .utils import RnYlm_1pbasis


def sparse_basis(phi, order: int, deg: int, wL: float = 1.5, species: int = 0) -> SymmetricBasis:
    """SparseBasis(maxorder, p = 1, weight n = 1, l = wL) over Rn * Ylm [* Categorical(species)]."""
    Bsel = SparseBasis(maxorder=order, p=1, default_maxdeg=deg, weight={"n": 1.0, "l": wL})
    b1p = RnYlm_1pbasis(maxdeg=deg, maxL=math.ceil(deg / wL), Bsel=None if species else Bsel)
    if species:
        b1p = Product1pBasis((Categorical1pBasis(list(range(species)), varsym="mu", idxsym="q"),) + b1p.bases)
    return SymmetricBasis(phi, b1p, Bsel)


@dataclass
class Workload:
    key: str
    title: str                      # goes into bench.py's config.workload
    build: Callable[[], SymmetricBasis]
    call: str                       # "B" (evaluate basis), "E" (evaluate model), "EF" (evaluate + grad_config)
    nprop: int
    J: int
    nenv: int                       # environments per GPU in bench.py
    nspecies: int = 0
    seed: int = 20240
    cache_key: str = ""


WORKLOADS: Dict[str, Workload] = {
    "1": Workload("1", "SymmetricBasis evaluate (B values), Invariant, ord=3, maxdeg=10, wL=1.5 SparseBasis, 30 neighbours (BASELINE config 1)",
                  lambda: sparse_basis(Invariant(), 3, 10), "B", 1, 30, 2_000_000, seed=20241, cache_key="inv_3_10"),
    "2": Workload("2", "LinearACEModel energy+forces, Invariant, ord=3, maxdeg=12, wL=1.5 SparseBasis, 40 neighbours (BASELINE config 2)",
                  lambda: sparse_basis(Invariant(), 3, 12), "EF", 1, 40, 1_000_000, seed=20242, cache_key="inv_3_12"),
    "3": Workload("3", "LinearACEModel energy+forces, Invariant, ord=4, maxdeg=14, wL=1.5 SparseBasis, 60 neighbours (BASELINE config 3)",
                  lambda: sparse_basis(Invariant(), 4, 14), "EF", 1, 60, 400_000, seed=20243, cache_key="inv_4_14"),
    "4a": Workload("4a", "SymmetricBasis evaluate (B values), EuclideanVector, ord=3, maxdeg=10, wL=1.5 SparseBasis, 30 neighbours (BASELINE config 4, L=1)",
                   lambda: sparse_basis(EuclideanVector(), 3, 10), "B", 1, 30, 1_000_000, seed=20244, cache_key="vec_3_10"),
    "4": Workload("4", "SymmetricBasis evaluate (B values), EuclideanMatrix, ord=3, maxdeg=10, wL=1.5 SparseBasis, 30 neighbours (BASELINE config 4, L=2)",
                  lambda: sparse_basis(EuclideanMatrix(), 3, 10), "B", 1, 30, 400_000, seed=20245, cache_key="mat_3_10"),
    "5": Workload("5", "LinearACEModel evaluate, 16 properties, 4-species Categorical1pBasis x Rn x Ylm, Invariant, ord=3, maxdeg=12, 40 neighbours (BASELINE config 5)",
                  lambda: sparse_basis(Invariant(), 3, 12, species=4), "E", 16, 40, 200_000, nspecies=4, seed=20246, cache_key="sp_3_12"),
    "5f": Workload("5f", "LinearACEModel energy + 16 force fields, 16 properties, 4-species Categorical1pBasis x Rn x Ylm, Invariant, ord=3, maxdeg=12, 40 neighbours (BASELINE config 5)",
                   lambda: sparse_basis(Invariant(), 3, 12, species=4), "EF", 16, 40, 100_000, nspecies=4, seed=20247, cache_key="sp_3_12"),
}
WORKLOADS["4b"] = WORKLOADS["4"]
# the Jacobian of config 1: evaluate_d(basis, cfg) = dB (266 x 30 x 3 per environment, 191 KB), the training-side call
# (benchmark/bm_basis.jl:64-70 "evaluate_d"; profile/profile_basis.jl:69: cost(dB) ~ 2 J cost(B))
WORKLOADS["1d"] = Workload("1d", "SymmetricBasis evaluate_d (Jacobian dB), Invariant, ord=3, maxdeg=10, wL=1.5 SparseBasis, 30 neighbours (BASELINE config 1, evaluate_d)",
                           lambda: sparse_basis(Invariant(), 3, 10), "dB", 1, 30, 100_000, seed=20248, cache_key="inv_3_10")

WORKLOADS["4ad"] = Workload("4ad", "SymmetricBasis evaluate_d (Jacobian dB), EuclideanVector, ord=3, maxdeg=10, wL=1.5 SparseBasis, 30 neighbours (BASELINE config 4, L=1, evaluate_d)",
                            lambda: sparse_basis(EuclideanVector(), 3, 10), "dB", 1, 30, 40_000, seed=20249, cache_key="vec_3_10")
WORKLOADS["4d"] = Workload("4d", "SymmetricBasis evaluate_d (Jacobian dB), EuclideanMatrix, ord=3, maxdeg=10, wL=1.5 SparseBasis, 30 neighbours (BASELINE config 4, L=2, evaluate_d)",
                           lambda: sparse_basis(EuclideanMatrix(), 3, 10), "dB", 1, 30, 5_000, seed=20250, cache_key="mat_3_10")

_BASIS_CACHE: Dict[str, SymmetricBasis] = {}


def build_basis(w: Workload) -> SymmetricBasis:
    if w.cache_key not in _BASIS_CACHE:
        _BASIS_CACHE[w.cache_key] = w.build()
    return _BASIS_CACHE[w.cache_key]


def coefficients(w: Workload, basis: SymmetricBasis) -> np.ndarray:
    """c ~ U(-0.5, 0.5) (test/test_linearmodel.jl:36), [nB][nprop]."""
    from .utils import philox
    return philox(w.seed + 1000).random((len(basis), w.nprop)) - 0.5


def algorithmic_work(basis: SymmetricBasis, J: int, call: str, nprop: int = 1) -> dict:
    """algorithmic flops per environment split by stage (1 add / mul = 1 flop, FMA = 2,
    complex mul = 6, complex x real = 2), and algorithmic HBM bytes per environment (inputs read once, outputs
    written once).  P = nprop * ncomp output channels."""
    b1p = basis.pibasis.basis1p
    Nn = len(b1p.component(0).R)
    L = max(b[b1p.sym_index("l")] for b in b1p.spec)
    sizeP, sizeY, nA = (L + 1) * (L + 2) // 2, (L + 1) ** 2, len(b1p)
    orders = np.asarray(basis.pibasis.spec.orders)
    ncomp = basis.A2Bmap.ncomp
    P = nprop * ncomp
    nsp = 4 if b1p.component(2) is not None else 0
    counts = {nu: int((orders == nu).sum()) for nu in range(1, int(orders.max()) + 1)}
    in_bytes = 24 * J + 8 + (4 * J if nsp else 0)
    if call == "dB":
        cs = 1 if basis.real else 2
        nnz, maxo = basis.A2Bmap.nnz, int(orders.max())
        pool = J * (28 + 5 * Nn + 15 * sizeP + 4 * nA)                            # A, as in the B call
        dpool = J * (6 * Nn + 22 * sizeP + 12 * nA)                               # dA: 3 complex components per one-particle function
        prod_B = float(sum(n * 6 * (nu - 1) for nu, n in counts.items()))
        couple_B = float(nnz * ncomp * (2 if basis.pibasis.real else 4 if basis.real else 8))
        # per neighbour and non-zero of A2Bmap: the local adjoints (2 (nu - 1) - 1 complex products) + nu complex x complex-3-vector
        # multiply-adds + the coupling coefficient applied to the 3-vector
        adj = float(J * sum(n * 6 * max(2 * (nu - 1) - 1, 0) for nu, n in counts.items()))
        jac = float(J * sum(n * nu * 3 * (4 if basis.pibasis.real else 8) for nu, n in counts.items()) + J * nnz * ncomp * 3 * 2)
        flops = {"pool": float(pool), "product_B": prod_B, "coupling_B": couple_B, "jacobian": float(dpool) + adj + jac}
        out_bytes = J * len(basis) * ncomp * 3 * 8 * cs + len(basis) * ncomp * 8 * cs
        return {"flops": flops, "flops_total": float(sum(flops.values())), "bytes": int(in_bytes + out_bytes),
                "in_bytes": int(in_bytes), "out_bytes": int(out_bytes)}
    if call == "B":
        cs = 1 if basis.real else 2
        pool = J * (28 + 5 * Nn + 15 * sizeP + 4 * nA)
        prod = float(sum(n * 6 * (nu - 1) for nu, n in counts.items()))
        # A2B . AA per non-zero and component: real x real FMA (real AA), Re(complex x complex) = 2 FMAs (real B),
        # or a full complex multiply-add
        couple = float(basis.A2Bmap.nnz * ncomp * (2 if basis.pibasis.real else 4 if basis.real else 8))
        flops = {"pool": float(pool), "product": prod, "coupling": couple}
        out_bytes = len(basis) * ncomp * 8 * cs
    else:
        f = lambda nu: 6 * (nu - 1) + 18 * max(nu - 2, 0) + (4 * nu + 2) * P      # noqa: E731
        fe = lambda nu: 6 * (nu - 1) + 2 * P                                      # noqa: E731  (energy only: product + readout)
        pool = J * (28 + 5 * Nn + 15 * sizeP + 4 * nA)
        if call == "EF":
            adj = float(sum(n * f(nu) for nu, n in counts.items()))
            forces = J * (6 * Nn + 22 * sizeP + (8 * nA + 16 * sizeY) * P)
            flops = {"pool": float(pool), "adjoint": adj, "forces": float(forces)}
            out_bytes = 24 * J * P + 8 * P
        else:
            flops = {"pool": float(pool), "adjoint": float(sum(n * fe(nu) for nu, n in counts.items())), "forces": 0.0}
            out_bytes = 8 * P
    return {"flops": flops, "flops_total": float(sum(flops.values())), "bytes": int(in_bytes + out_bytes),
            "in_bytes": int(in_bytes), "out_bytes": int(out_bytes)}

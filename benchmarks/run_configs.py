#!/usr/bin/env python
"""All five BASELINE.json configs on one B200: throughput (device-resident inputs, CUDA events) and a sampled
parity check against the oracle.  bench.py measures config 2 (the headline); this script is the evidence for the
other rows.  Usage: python benchmarks/run_configs.py [--scale 1.0] [--out profiles/rN_configs.json]"""
import argparse
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import ace_jl_b200 as ace  # noqa: E402
from ace_jl_b200.descriptor import basis_descriptor  # noqa: E402
from ace_jl_b200.utils import RnYlm_1pbasis, philox, rand_envs  # noqa: E402
from oracle import Oracle  # noqa: E402


def sparse(phi, order, deg, wL=1.5, species=0):
    Bsel = ace.SparseBasis(maxorder=order, p=1, default_maxdeg=deg, weight={"n": 1.0, "l": wL})
    b1p = RnYlm_1pbasis(maxdeg=deg, maxL=math.ceil(deg / wL), Bsel=None if species else Bsel)
    if species:
        b1p = ace.Product1pBasis((ace.Categorical1pBasis(list(range(species)), varsym="mu", idxsym="q"),) + b1p.bases)
    return ace.SymmetricBasis(phi, b1p, Bsel)


CONFIGS = [
    # name, builder, nprop, J, nenv, call
    ("config1: SymmetricBasis values, Invariant, ord 3, deg 10, J=30", lambda: sparse(ace.Invariant(), 3, 10), 1, 30, 400_000, "B"),
    ("config2: LinearACEModel E+F, ord 3, deg 12, J=40", lambda: sparse(ace.Invariant(), 3, 12), 1, 40, 1_000_000, "EF"),
    ("config3: LinearACEModel E+F, ord 4, deg 14, J=60", lambda: sparse(ace.Invariant(), 4, 14), 1, 60, 200_000, "EF"),
    ("config4a: SymmetricBasis values, EuclideanVector, ord 3, deg 10, J=30", lambda: sparse(ace.EuclideanVector(), 3, 10), 1, 30, 200_000, "B"),
    ("config4b: SymmetricBasis values, EuclideanMatrix, ord 3, deg 10, J=30", lambda: sparse(ace.EuclideanMatrix(), 3, 10), 1, 30, 100_000, "B"),
    ("config5: 16-property LinearACEModel energies, 4 species, ord 3, deg 12, J=40", lambda: sparse(ace.Invariant(), 3, 12, species=4), 16, 40, 100_000, "E"),
    ("config5f: 16-property LinearACEModel E+F, 4 species, ord 3, deg 12, J=40", lambda: sparse(ace.Invariant(), 3, 12, species=4), 16, 40, 50_000, "EF"),
]


def relerr(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--out", default="")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    results = []
    cache = {}
    for name, build, nprop, J, nenv, call in CONFIGS:
        if args.only and args.only not in name:
            continue
        nenv = max(64, int(nenv * args.scale))
        t0 = time.time()
        key = name.split(":")[0].rstrip("f")
        basis = cache.get(key) or build()
        cache[key] = basis
        tb = time.time() - t0
        rng = philox(20240 + len(results))
        c = rng.random((len(basis), nprop)) - 0.5
        model = ace.LinearACEModel(basis, c if nprop > 1 else c[:, 0])
        h = model.evaluator.handle
        cat = basis.pibasis.basis1p.component(2)
        R, off, sp = rand_envs(rng, basis.pibasis.basis1p.component(0), nenv, J, len(cat) if cat else 0)
        b = ace.B200Batch(torch.from_numpy(R).cuda(), torch.from_numpy(off).cuda(), None if sp is None else torch.from_numpy(sp).cuda())
        fn = {"B": h.eval_B, "E": h.energy, "EF": h.energy_forces}[call]
        out = fn(b)
        torch.cuda.synchronize()
        reps = 3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = fn(b)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        # sampled parity against the oracle
        ns = min(nenv, 200 if len(basis.pibasis) > 20000 else 1000)
        sel = np.sort(rng.choice(nenv, size=ns, replace=False))
        Rs = np.concatenate([R[off[e]:off[e + 1]] for e in sel])
        sps = None if sp is None else np.concatenate([sp[off[e]:off[e + 1]] for e in sel])
        offs = np.arange(ns + 1) * J
        o = Oracle(basis_descriptor(basis, c))
        if call == "B":
            err = {"B": relerr(out.cpu().numpy()[sel], o.eval_B(Rs, offs, sps))}
        elif call == "E":
            err = {"E": relerr(out.cpu().numpy()[sel], o.energy(Rs, offs, sps))}
        else:
            Eo, Go = o.energy_forces(Rs, offs, sps)
            G = out[1].cpu().numpy().reshape(nenv, J, *out[1].shape[1:])[sel].reshape(ns * J, *out[1].shape[1:])
            err = {"E": relerr(out[0].cpu().numpy()[sel], Eo), "G": relerr(G, Go)}
        o_ = basis.pibasis.spec.orders
        res = {"config": name, "call": call, "nenv": nenv, "J": J, "nprop": nprop,
               "nA": len(basis.pibasis.basis1p), "nAA": len(basis.pibasis), "nB": len(basis), "ncomp": basis.A2Bmap.ncomp,
               "nAA_by_order": [int((o_ == k).sum()) for k in range(int(o_.max()) + 1)],
               "ms": ms, "env_per_s": nenv / (ms * 1e-3), "stage_ms": h.last_stage_ms() if call != "B" else None,
               "rel_err_vs_oracle": err, "parity_ok": all(v < 1e-12 for v in err.values()), "build_s": tb}
        print(json.dumps(res), flush=True)
        results.append(res)
        del model, h, b, out
        torch.cuda.empty_cache()
    if args.out:
        json.dump(results, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()

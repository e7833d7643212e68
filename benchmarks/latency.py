#!/usr/bin/env python
"""Per-call latency of the C ABI at small batch sizes (the reference's natural usage is one evaluate(model, cfg) per
environment): wall-clock time per call for nenv = 1, 32, 1024, 32768 with 40 neighbours, device-resident and host
buffers, energy + forces and energy only.  Usage: python benchmarks/latency.py > profiles/r2_latency.txt"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import ace_jl_b200 as ace  # noqa: E402
from ace_jl_b200.utils import philox, rand_envs  # noqa: E402
from ace_jl_b200.workloads import WORKLOADS, build_basis, coefficients  # noqa: E402

w = WORKLOADS["2"]
basis = build_basis(w)
c = coefficients(w, basis)
h = ace.LinearACEModel(basis, c[:, 0]).evaluator.handle
print("config 2 (ord 3, deg 12, 40 neighbours); microseconds per call, median of 200 calls (wall clock, call returns synchronised)")
print(f"{'nenv':>7} {'E+F device':>12} {'E+F host':>12} {'E device':>12} {'E host':>12} {'launches/call':>14}")
for nenv in (1, 32, 1024, 32768):
    R, off, _ = rand_envs(philox(3), basis.pibasis.basis1p.component(0), nenv, 40)
    bd = ace.B200Batch(torch.from_numpy(R).cuda(), torch.from_numpy(off).cuda())
    bh = ace.B200Batch(R, off)
    res = []
    for b, fn in ((bd, h.energy_forces), (bh, h.energy_forces), (bd, h.energy), (bh, h.energy)):
        for _ in range(20):
            fn(b)
        torch.cuda.synchronize()
        ts = []
        for _ in range(200 if nenv <= 1024 else 50):
            t0 = time.perf_counter()
            fn(b)
            ts.append(time.perf_counter() - t0)
        res.append(1e6 * float(np.median(ts)))
    l0 = h.launch_count()
    h.energy_forces(bd)
    print(f"{nenv:>7} {res[0]:>12.1f} {res[1]:>12.1f} {res[2]:>12.1f} {res[3]:>12.1f} {h.launch_count() - l0:>14}")

#!/usr/bin/env python
"""Does co-scheduling the three kernels of different sub-batches help?  Config 2, 10^6 environments in HBM:
one call on one stream vs K threads, each with its own torch stream and 1/K of the batch (the handle's context
pool gives every thread private workspaces).  Usage: python benchmarks/overlap_probe.py"""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import ace_jl_b200 as ace  # noqa: E402
from ace_jl_b200.utils import philox, rand_envs  # noqa: E402
from ace_jl_b200.workloads import WORKLOADS, build_basis, coefficients  # noqa: E402

w = WORKLOADS["2"]
basis = build_basis(w)
h = ace.LinearACEModel(basis, coefficients(w, basis)[:, 0]).evaluator.handle
nenv = 1_000_000
R, off, _ = rand_envs(philox(3), basis.pibasis.basis1p.component(0), nenv, w.J)
Rd, offd = torch.from_numpy(R).cuda(), torch.from_numpy(off).cuda()


def run(K, reps=5):
    per = nenv // K
    parts = []
    for k in range(K):
        o = offd[k * per: (k + 1) * per + 1]
        parts.append((ace.B200Batch(Rd[k * per * w.J: (k + 1) * per * w.J], o - o[0]), torch.cuda.Stream(),
                      torch.empty((per, 1, 1), dtype=torch.float64, device="cuda"),
                      torch.empty((per * w.J, 1, 3, 1), dtype=torch.float64, device="cuda")))
    torch.cuda.synchronize()

    def work(p):
        b, s, E, G = p
        with torch.cuda.stream(s):
            for _ in range(reps):
                h.energy_forces(b, E, G)
    th = [threading.Thread(target=work, args=(p,)) for p in parts]      # warm-up, concurrent: every thread's context allocates its workspaces
    for t in th:
        t.start()
    for t in th:
        t.join()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(p,)) for p in parts]
    for t in th:
        t.start()
    for t in th:
        t.join()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"K = {K}: {1e3 * dt:.3f} ms per 10^6 environments, {nenv / dt:.3e} env/s", flush=True)


for K in (1, 2, 3, 4, 8):
    run(K)

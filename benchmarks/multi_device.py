#!/usr/bin/env python
"""One handle, N GPUs (aceb200_set_devices): end-to-end env/s of the per-environment call with pinned HOST buffers as a
function of the number of devices, config 2.  Usage (on a multi-GPU box): python benchmarks/multi_device.py [--envs 1000000]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import ace_jl_b200 as ace  # noqa: E402
from ace_jl_b200 import _lib  # noqa: E402
from ace_jl_b200.utils import philox, rand_envs  # noqa: E402
from ace_jl_b200.workloads import WORKLOADS, build_basis, coefficients  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1_000_000, help="environments per device")
args = ap.parse_args()
ndev = _lib.load().aceb200_device_count()
w = WORKLOADS["2"]
basis = build_basis(w)
c = coefficients(w, basis)
h = ace.LinearACEModel(basis, c[:, 0]).evaluator.handle
out = []
for n in [k for k in (1, 2, 4, 8) if k <= ndev]:
    nenv = args.envs * n
    R, off, _ = rand_envs(philox(7), basis.pibasis.basis1p.component(0), nenv, w.J)
    Rh, offh = torch.from_numpy(R).pin_memory(), torch.from_numpy(off).pin_memory()
    E = torch.empty((nenv, 1, 1), dtype=torch.float64).pin_memory()
    G = torch.empty((nenv * w.J, 1, 3, 1), dtype=torch.float64).pin_memory()
    b = ace.B200Batch(Rh.numpy(), offh.numpy())
    h.set_devices(list(range(n)))
    h.energy_forces(b, E.numpy(), G.numpy())
    t0 = time.perf_counter()
    for _ in range(5):
        h.energy_forces(b, E.numpy(), G.numpy())
    dt = (time.perf_counter() - t0) / 5
    r = {"devices": n, "envs": nenv, "ms_per_call": 1e3 * dt, "env_per_s": nenv / dt,
         "host_bytes_per_s": (R.nbytes + off.nbytes + E.numel() * 8 + G.numel() * 8) / dt}
    print(json.dumps(r), flush=True)
    out.append(r)
    del Rh, offh, E, G
for r in out:
    r["efficiency_vs_1"] = r["env_per_s"] / (out[0]["env_per_s"] * r["devices"])
print(json.dumps({"summary": out}))

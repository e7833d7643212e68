#!/usr/bin/env python
"""Where the time of aceb200_structure_energy_forces goes (config 2 model, 10^6-atom FCC crystal, pinned host buffers):
wall clock per call vs the device interval of its kernels, with and without the caller's reverse table, packed and
unpacked neighbour words.  Usage: python benchmarks/struct_profile.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import ace_jl_b200 as ace  # noqa: E402
from ace_jl_b200.structure import B200Structure, pack_neighbours, reverse_pairs  # noqa: E402
from ace_jl_b200.utils import fcc_structure, philox  # noqa: E402
from ace_jl_b200.workloads import WORKLOADS, build_basis, coefficients  # noqa: E402

w = WORKLOADS["2"]
basis = build_basis(w)
h = ace.LinearACEModel(basis, coefficients(w, basis)[:, 0]).evaluator.handle
X, cell, first, nbr, img = fcc_structure(philox(1), 63)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()   # noqa: E731
E = torch.empty((len(X), 1, 1), dtype=torch.float64).pin_memory().numpy()
F = torch.empty((len(X), 1, 3, 1), dtype=torch.float64).pin_memory().numpy()
W = torch.empty((1, 3, 3), dtype=torch.float64).pin_memory().numpy()
rev = None
cases = {
    "unpacked (7 B/pair), reverse pairs found on the device": B200Structure(pin(X), pin(first), pin(nbr), pin(img), cell),
    "packed (4 B/pair), reverse pairs found on the device": B200Structure(pin(X), pin(first), pin(pack_neighbours(nbr, img)), None, cell, packed=True),
}
print(f"{len(X)} atoms, {len(nbr)} pairs")
for name, st in cases.items():
    for _ in range(3):
        h.structure_energy_forces(st, True, E, F, W)
    ts, ks = [], []
    for _ in range(10):
        t0 = time.perf_counter()
        h.structure_energy_forces(st, True, E, F, W)
        ts.append(time.perf_counter() - t0)
        ks.append(h.last_kernel_ms())
    print(f"{name}: wall {1e3 * np.median(ts):.2f} ms per call, device interval of the kernels {np.median(ks):.2f} ms, "
          f"{len(X) / np.median(ts):.3e} atoms/s")
Xd, fd, nd = (torch.from_numpy(a).cuda() for a in (X, first, pack_neighbours(nbr, img)))
sd = B200Structure(Xd, fd, nd, None, cell, packed=True)
for _ in range(3):
    h.structure_energy_forces(sd)
torch.cuda.synchronize()
ts = []
for _ in range(10):
    t0 = time.perf_counter()
    h.structure_energy_forces(sd)
    torch.cuda.synchronize()
    ts.append(time.perf_counter() - t0)
print(f"device-resident structure: wall {1e3 * np.median(ts):.2f} ms per call, kernels {h.last_kernel_ms():.2f} ms")

"""Sweep the chunk size of aceb200_structure_energy_forces (ACEB200_STRUCT_MB) on the 10^6-atom benchmark structure."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, ace_jl_b200 as ace
from ace_jl_b200.structure import B200Structure, reverse_pairs
from ace_jl_b200.utils import fcc_structure, philox

basis, c = bench.build_model()
model = ace.LinearACEModel(basis, c); h = model.evaluator.handle
X, cell, first, nbr, img = fcc_structure(philox(1), 63)
pin = lambda a: torch.from_numpy(a).pin_memory()
pX, pf, pn, pi = pin(X), pin(first), pin(nbr), pin(img)
st = B200Structure(pX.numpy(), pf.numpy(), pn.numpy(), pi.numpy(), cell)
E = torch.empty((st.natoms, 1, 1), dtype=torch.float64).pin_memory()
F = torch.empty((st.natoms, 1, 3, 1), dtype=torch.float64).pin_memory()
W = torch.empty((1, 3, 3), dtype=torch.float64).pin_memory()
for mb in (2, 4, 8, 16, 32, 64, 1000):
    os.environ["ACEB200_STRUCT_MB"] = str(mb)
    h.structure_energy_forces(st, True, E.numpy(), F.numpy(), W.numpy())
    t0 = time.perf_counter()
    for _ in range(5):
        h.structure_energy_forces(st, True, E.numpy(), F.numpy(), W.numpy())
    dt = (time.perf_counter() - t0) / 5
    print(f"STRUCT_MB {mb}: {dt*1e3:.2f} ms  {st.natoms/dt:.3g} atoms/s  kernels {h.last_kernel_ms():.2f} ms", flush=True)
# device-resident structure: no copies at all
t = lambda a: torch.from_numpy(a).cuda()
sd = B200Structure(t(X), t(first), t(nbr), t(img), cell)
Ed, Fd, Wd = h.structure_energy_forces(sd)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    h.structure_energy_forces(sd, True, Ed, Fd, Wd)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(f"device-resident: {dt*1e3:.2f} ms  {sd.natoms/dt:.3g} atoms/s", flush=True)

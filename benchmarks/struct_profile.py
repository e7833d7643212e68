"""One device-resident call of aceb200_structure_energy_forces on the 10^6-atom benchmark structure (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, ace_jl_b200 as ace
from ace_jl_b200.structure import B200Structure
from ace_jl_b200.utils import fcc_structure, philox
basis, c = bench.build_model()
h = ace.LinearACEModel(basis, c).evaluator.handle
X, cell, first, nbr, img = fcc_structure(philox(1), int(sys.argv[1]) if len(sys.argv) > 1 else 63)
t = lambda a: torch.from_numpy(a).cuda()
sd = B200Structure(t(X), t(first), t(nbr), t(img), cell)
for _ in range(3):
    E, F, W = h.structure_energy_forces(sd)
torch.cuda.synchronize()
print("ok", float(E.sum()))

"""evaluate(model, cfg) alone (aceb200_energy) on BASELINE config 2, device-resident: per-stage times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, ace_jl_b200 as ace
from ace_jl_b200.utils import philox, rand_envs
basis, c = bench.build_model()
h = ace.LinearACEModel(basis, c).evaluator.handle
nenv = 1_000_000
R, off, _ = rand_envs(philox(3), basis.pibasis.basis1p.component(0), nenv, 40)
b = ace.B200Batch(torch.from_numpy(R).cuda(), torch.from_numpy(off).cuda())
for _ in range(3):
    E = h.energy(b)
torch.cuda.synchronize()
print("energy only:", h.last_stage_ms(), "kernel ms", h.last_kernel_ms())
E2, G = h.energy_forces(b)
torch.cuda.synchronize()
print("energy+forces:", h.last_stage_ms(), "max |E - E2| / max|E|", float((E - E2).abs().max() / E2.abs().max()))

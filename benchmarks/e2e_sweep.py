import os, sys, time, json
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench, ace_jl_b200 as ace
from ace_jl_b200.utils import philox, rand_envs
basis, c = bench.build_model()
nenv, J = 1_000_000, 40
R, off, _ = rand_envs(philox(1), basis.pibasis.basis1p.component(0), nenv, J)
Rh = torch.from_numpy(R).pin_memory(); offh = torch.from_numpy(off).pin_memory()
Eh = torch.empty((nenv,1,1), dtype=torch.float64).pin_memory(); Gh = torch.empty((nenv*J,1,3,1), dtype=torch.float64).pin_memory()
for lanes in (2,3,4):
    for mb in (8, 16, 32, 64, 128):
        os.environ['ACEB200_LANES']=str(lanes); os.environ['ACEB200_PIPE_MB']=str(mb)
        model = ace.LinearACEModel(basis, c); h = model.evaluator.handle
        hb = ace.B200Batch(Rh.numpy(), offh.numpy())
        h.energy_forces(hb, Eh.numpy(), Gh.numpy())
        t0=time.perf_counter()
        for _ in range(3): h.energy_forces(hb, Eh.numpy(), Gh.numpy())
        dt=(time.perf_counter()-t0)/3
        print(lanes, mb, '%.1f ms  %.3g env/s'%(dt*1e3, nenv/dt), flush=True)
        del model, h

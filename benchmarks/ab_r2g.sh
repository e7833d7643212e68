#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q > $OUT/abg_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/abg_tests.log
B="python bench.py --no-cpu --no-e2e --steps 5"
ACEB200_VERBOSE=1 $B --config 3 > $OUT/abg_c3.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/abg_c*.log')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, '%.3e env/s'%d['value'], '%.3f ms'%d['ms_per_step'], d['roofline'].get('ms_per_launch'), d['roofline']['bound'], round(d['roofline']['frac'],3), d['parity'])
    except Exception as e:
        print(f, 'FAILED', open(f).read()[-1200:])
PY
python - <<'PY'
# energy-only call, config 2, 10^6 environments: readout pass of k_basis_stream vs the Euler-identity energy stream
import os, subprocess, sys, json
code = r'''
import torch, ace_jl_b200 as ace, numpy as np
from ace_jl_b200.workloads import *
from ace_jl_b200.utils import philox, rand_envs
w=WORKLOADS["2"]; basis=build_basis(w); c=coefficients(w,basis)
h=ace.LinearACEModel(basis,c[:,0]).evaluator.handle
R,off,_=rand_envs(philox(1),basis.pibasis.basis1p.component(0),1000000,40)
b=ace.B200Batch(torch.from_numpy(R).cuda(),torch.from_numpy(off).cuda())
E=h.energy(b); torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): h.energy(b,E)
e1.record(); torch.cuda.synchronize()
print("energy-only config 2: %.3f ms per 1e6 env"%(e0.elapsed_time(e1)/10), h.last_stage_ms())
'''
for env in ({}, {"ACEB200_NO_ENERGY_BSTREAM": "1"}):
    e = dict(os.environ); e.update(env)
    print(env, subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True).stdout.strip())
PY

#!/bin/bash
# config-2 A/B of the DMMA kernels against the FMA kernels + one ncu capture (run under gpurun)
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-x}
B="python bench.py --no-cpu --no-e2e --steps 10 --config 2"
$B > $OUT/ab2_${TAG}_mma.log 2>&1
ACEB200_POOL_MMA=0 ACEB200_FORCES_MMA=0 $B > $OUT/ab2_${TAG}_fma.log 2>&1
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/ab2_${TAG}_*.log')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, '%.3e env/s'%d['value'], '%.3f ms'%d['ms_per_step'], d['roofline'].get('ms_per_launch'), d['parity']['ok'])
    except Exception as e:
        print(f, 'FAILED', open(f).read()[-800:])
PY
ncu --set full --clock-control none --import-source on -k regex:"k_pool_mma|k_forces_mma" -s 4 -c 2 -f -o $OUT/r2_mma_${TAG} python bench.py --config 2 --envs 200000 --steps 1 --warmup 3 --no-cpu --no-e2e > $OUT/r2_mma_${TAG}_ncu.log 2>&1

#!/bin/bash
# A/B of the round-2 kernels on one B200 (run under gpurun): logs -> gpurun_out/ab_*.log
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/ab_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/ab_tests.log
B="python bench.py --no-cpu --no-e2e --steps 10"
$B --config 2 > $OUT/ab_c2_mma.log 2>&1
ACEB200_POOL_MMA=0 $B --config 2 > $OUT/ab_c2_poolfma.log 2>&1
ACEB200_FORCES_MMA=0 $B --config 2 > $OUT/ab_c2_forcesfma.log 2>&1
for c in 1 3 4a 4 5; do $B --config $c --steps 5 > $OUT/ab_c$c.log 2>&1; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/ab_c*.log')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, '%.3e env/s'%d['value'], '%.3f ms'%d['ms_per_step'], d['roofline'].get('ms_per_launch'), d.get('parity'))
    except Exception as e:
        print(f, 'FAILED', open(f).read()[-600:])
PY

OUT=gpurun_out
B="python bench.py --no-cpu --no-e2e --steps 5 --config 3"
ACEB200_EPL=2 ACEB200_VERBOSE=1 $B > $OUT/exp_c3_epl2.log 2>&1
$B > $OUT/exp_c3_epl1.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/exp_c3_*.log')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, '%.3e'%d['value'], d['roofline']['ms_per_launch'])
    except Exception as e: print(f,'FAILED',open(f).read()[-600:])
PY
grep -h "stream:" $OUT/exp_c3_epl2.log

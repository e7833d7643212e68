#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/abe_tests.log 2>&1; echo "tests rc=$?"; tail -4 $OUT/abe_tests.log
python benchmarks/latency.py > $OUT/r2_latency.txt 2>&1; cat $OUT/r2_latency.txt
B="python bench.py --no-cpu --steps 5"
ACEB200_VERBOSE=1 $B --config 5 --no-e2e > $OUT/abe_c5.log 2>&1
ACEB200_NO_ENERGY_BSTREAM=1 $B --config 5 --no-e2e > $OUT/abe_c5_old.log 2>&1
$B --config 2 --steps 10 > $OUT/abe_c2.log 2>&1
$B --config 3 --no-e2e > $OUT/abe_c3.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/abe_c*.log')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, '%.3e env/s'%d['value'], '%.3f ms'%d['ms_per_step'], d['roofline'].get('ms_per_launch'), d['roofline']['bound'], round(d['roofline']['frac'],3), d['parity']['ok'], 'e2e', d.get('e2e',{}).get('value'), 'e2e_env', d.get('e2e_per_environment',{}).get('value'))
    except Exception as e:
        print(f, 'FAILED', open(f).read()[-1200:])
PY
grep -h "energy stream" $OUT/abe_c*.log | sort -u

#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
B="python bench.py --no-cpu --no-e2e --steps 5"
for c in 1 4a 4; do ACEB200_VERBOSE=1 $B --config $c > $OUT/abd_c${c}.log 2>&1; done
ACEB200_BASIS_WARPS=8 $B --config 4 > $OUT/abd_c4_w8.log 2>&1
ACEB200_BASIS_EPL=1 $B --config 4 > $OUT/abd_c4_epl1.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/abd_c*.log')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, '%.3e env/s'%d['value'], '%.3f ms'%d['ms_per_step'], d['roofline'].get('ms_per_launch'), d['roofline']['bound'], round(d['roofline']['frac'],3), d['parity']['ok'])
    except Exception as e:
        print(f, 'FAILED', open(f).read()[-800:])
PY
grep -h "basis stream" $OUT/abd_c*.log | sort -u
ncu --set full --clock-control none --import-source on -k regex:"k_basis_stream" -s 3 -c 1 -f -o $OUT/r2_basis_c4d python bench.py --config 4 --envs 50000 --steps 1 --warmup 3 --no-cpu --no-e2e > $OUT/r2_basis_c4d_ncu.log 2>&1

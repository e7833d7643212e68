#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q > $OUT/abf_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/abf_tests.log
B="python bench.py --no-cpu --no-e2e --steps 5"
for c in 1 4a 4 5; do ACEB200_VERBOSE=1 $B --config $c > $OUT/abf_c${c}.log 2>&1; done
ACEB200_BASIS_WARPS=8 $B --config 5 > $OUT/abf_c5_w8.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/abf_c*.log')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, '%.3e env/s'%d['value'], '%.3f ms'%d['ms_per_step'], d['roofline'].get('ms_per_launch'), d['roofline']['bound'], round(d['roofline']['frac'],3), d['parity']['ok'])
    except Exception as e:
        print(f, 'FAILED', open(f).read()[-1200:])
PY
grep -h "stream:" $OUT/abf_c*.log | sort -u
ncu --set full --clock-control none --import-source on -k regex:"k_basis_stream" -s 3 -c 1 -f -o $OUT/r2_basis_c5 python bench.py --config 5 --envs 20000 --steps 1 --warmup 3 --no-cpu --no-e2e > $OUT/r2_basis_c5_ncu.log 2>&1

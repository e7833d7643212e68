#!/bin/bash
# round 2, second GPU pass: whole -m gpu suite (incl. full-size configs), value-path benches, ncu of the fused basis kernel
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q --durations=8 > $OUT/abb_tests.log 2>&1; echo "tests rc=$?"; tail -14 $OUT/abb_tests.log
B="python bench.py --no-cpu --no-e2e --steps 5"
for c in 1 4a 4; do ACEB200_VERBOSE=1 $B --config $c > $OUT/abb_c$c.log 2>&1; ACEB200_NO_BASIS_STREAM=1 $B --config $c > $OUT/abb_c${c}_old.log 2>&1; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/abb_c*.log')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, '%.3e env/s'%d['value'], '%.3f ms'%d['ms_per_step'], d['roofline'].get('ms_per_launch'), d['roofline']['bound'], round(d['roofline']['frac'],3), d['parity']['ok'])
    except Exception as e:
        print(f, 'FAILED', open(f).read()[-800:])
PY
grep -h "basis stream" $OUT/abb_c*.log | sort -u
ncu --set full --clock-control none --import-source on -k regex:"k_basis_stream" -s 3 -c 1 -f -o $OUT/r2_basis_c4 python bench.py --config 4 --envs 50000 --steps 1 --warmup 3 --no-cpu --no-e2e > $OUT/r2_basis_c4_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_basis_stream|k_pool_mma" -s 6 -c 2 -f -o $OUT/r2_basis_c1 python bench.py --config 1 --envs 400000 --steps 1 --warmup 3 --no-cpu --no-e2e > $OUT/r2_basis_c1_ncu.log 2>&1

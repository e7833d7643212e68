OUT=gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q 2>&1 | tail -2
B="python bench.py --no-cpu --no-e2e --steps 5"
for c in 1 4a 4; do $B --config $c > $OUT/exp_basis_c$c.log 2>&1; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/exp_basis_c*.log')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, '%.3e'%d['value'], d['roofline']['ms_per_launch'], d['parity']['ok'])
    except Exception as e: print(f,'FAILED',open(f).read()[-600:])
PY

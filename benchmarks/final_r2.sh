#!/bin/bash
# Round-2 evidence pass on one B200: bench lines of every BASELINE config, reference arm, launch list, ncu captures,
# latency, sanitizers on the new kernels.  Everything lands in gpurun_out/ (copied to profiles/ by hand).
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/fin_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/fin_tests.log
python bench.py > $OUT/fin_bench_n1.json 2> $OUT/fin_bench_n1.err; tail -c 400 $OUT/fin_bench_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/fin_bench_ref.json 2>&1
for c in 1 1d 3 4a 4 5 5f; do python bench.py --config $c --steps 5 > $OUT/fin_bench_c$c.json 2> $OUT/fin_bench_c$c.err; done
python benchmarks/latency.py > $OUT/fin_latency.txt 2>&1
# launch list of the bench command (shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 60 --csv --log-file $OUT/fin_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --envs 1000000 > $OUT/fin_launches.log 2>&1
# full captures: the three kernels of the headline step, and the fused value kernel
ncu --set full --clock-control none --import-source on -k regex:"k_pool_mma|k_adjoint_stream|k_forces" -s 9 -c 3 -f -o $OUT/fin_c2 python bench.py --config 2 --envs 200000 --steps 1 --warmup 3 --no-cpu --no-e2e > $OUT/fin_c2_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_pool_mma|k_basis_stream" -s 6 -c 2 -f -o $OUT/fin_c1 python bench.py --config 1 --envs 400000 --steps 1 --warmup 3 --no-cpu --no-e2e > $OUT/fin_c1_ncu.log 2>&1
ncu --set full --clock-control none -k regex:"k_basis_stream" -s 3 -c 1 -f -o $OUT/fin_c4 python bench.py --config 4 --envs 50000 --steps 1 --warmup 3 --no-cpu --no-e2e > $OUT/fin_c4_ncu.log 2>&1
ncu --set full --clock-control none -k regex:"k_adjoint_stream" -s 3 -c 1 -f -o $OUT/fin_c3 python bench.py --config 3 --envs 100000 --steps 1 --warmup 3 --no-cpu --no-e2e > $OUT/fin_c3_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_dA|k_dB_env" -s 2 -c 2 -f -o $OUT/fin_c1d python bench.py --config 1d --envs 20000 --steps 1 --warmup 1 --no-cpu --no-e2e > $OUT/fin_c1d_ncu.log 2>&1
# the reports are large (gpurun brings back at most 64 MiB): keep the raw and source pages as (gzipped) CSV instead
for r in fin_c2 fin_c1 fin_c4 fin_c3 fin_c1d; do
  [ -f $OUT/$r.ncu-rep ] || continue
  ncu -i $OUT/$r.ncu-rep --page raw --csv > $OUT/${r}_raw.csv 2>/dev/null
  ncu -i $OUT/$r.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null | gzip > $OUT/${r}_src.csv.gz
  rm -f $OUT/$r.ncu-rep
done
# sanitizers on the kernels that are new this round (pooling on DMMA, fused basis stream, two-level adjoint stream, dp pullback)
SAN=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  timeout 900 $SAN --tool $tool --error-exitcode 0 --log-file $OUT/fin_sanitizer_$tool.log python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "all_entry_points or edge or multichunk or adjoint_eval" > $OUT/fin_sanitizer_$tool.pytest.log 2>&1
  echo "$tool: $(tail -1 $OUT/fin_sanitizer_$tool.pytest.log) | $(tail -1 $OUT/fin_sanitizer_$tool.log)"
done

#!/bin/bash
# compute-sanitizer passes over the C-ABI parity suite (run under gpurun; logs -> gpurun_out/sanitizer_*.log)
# usage: benchmarks/sanitize.sh [tag]
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
SAN=/usr/local/cuda/bin/compute-sanitizer
SEL='not config2 and not full_size and not concurrent'
for tool in memcheck racecheck synccheck initcheck; do
  extra=""
  [ $tool = memcheck ] && extra="--leak-check no"
  [ $tool = racecheck ] && extra="--racecheck-report all"
  echo "=== $tool" 
  timeout 900 $SAN --tool $tool $extra --error-exitcode 0 --log-file $OUT/sanitizer_${TAG}_$tool.log \
      python -m pytest tests/test_gpu_parity.py tests/test_structure.py -q -m gpu -k "$SEL" -x > $OUT/sanitizer_${TAG}_$tool.pytest.log 2>&1
  echo "rc=$?"; tail -3 $OUT/sanitizer_${TAG}_$tool.pytest.log; tail -5 $OUT/sanitizer_${TAG}_$tool.log
done

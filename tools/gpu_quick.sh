#!/bin/bash
# quick GPU visit: a pytest selection + one bench line.  bash tools/gpu_quick.sh <tag> <pytest -k expr or ''> <bench args...>
TAG=$1; shift
KEXPR=$1; shift
OUT=gpurun_out
mkdir -p $OUT
if [ -n "$KEXPR" ]; then
  timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "$KEXPR" > $OUT/pytest_$TAG.log 2>&1
  echo "pytest exit $?" >> $OUT/pytest_$TAG.log
  tail -25 $OUT/pytest_$TAG.log
fi
timeout 1200 python bench.py "$@" > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench exit $?"
cat $OUT/bench_$TAG.json
tail -5 $OUT/bench_$TAG.err

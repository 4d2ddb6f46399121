#!/bin/bash
# One GPU visit: [pytest selection] + [bench line(s)].  Usage (through gpurun):
#   bash tools/gpu_visit.sh <tag> "<pytest -k expression | all | none>" [bench args...]
TAG=$1; shift
KEXPR=$1; shift
OUT=gpurun_out
mkdir -p $OUT
if [ "$KEXPR" = "all" ]; then
  timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $OUT/pytest_$TAG.log 2>&1
  echo "pytest exit $?" >> $OUT/pytest_$TAG.log
  tail -40 $OUT/pytest_$TAG.log
elif [ "$KEXPR" != "none" ]; then
  timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -k "$KEXPR" > $OUT/pytest_$TAG.log 2>&1
  echo "pytest exit $?" >> $OUT/pytest_$TAG.log
  tail -40 $OUT/pytest_$TAG.log
fi
if [ $# -gt 0 ]; then
  timeout 1200 python bench.py "$@" > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
  echo "bench exit $?"
  cat $OUT/bench_$TAG.json
  tail -5 $OUT/bench_$TAG.err
fi

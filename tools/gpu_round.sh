#!/bin/bash
# One GPU visit: parity suite, smoke, full-size bench, ncu launch list + one full capture of the top kernel.
# Usage (through gpurun): bash tools/gpu_round.sh <tag> [sampler]
TAG=${1:-r01}
SAMPLER=${2:-tc}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > $OUT/pytest_$TAG.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_$TAG.log
tail -15 $OUT/pytest_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1
echo "smoke exit $?" >> $OUT/smoke_$TAG.log
tail -3 $OUT/smoke_$TAG.log
timeout 900 python bench.py --sampler $SAMPLER > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench exit $?"
cat $OUT/bench_$TAG.json
tail -3 $OUT/bench_$TAG.err
SMALL="python bench.py --sampler $SAMPLER --batch 64 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"   # full 100-iteration schedule: kernel shares comparable with the bench line
timeout 300 $SMALL > $OUT/plain_$TAG.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/launches_$TAG.csv $SMALL > $OUT/ncu_list_$TAG.log 2>&1
echo "ncu list exit $?"

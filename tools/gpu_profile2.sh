#!/bin/bash
# full ncu capture of one kernel family on a short bench command.  bash tools/gpu_profile2.sh <tag> <kernel regex> [skip] [count]
TAG=$1; REGEX=$2; SKIP=${3:-3}; CNT=${4:-1}
OUT=gpurun_out
mkdir -p $OUT
SMALL="python bench.py --sampler tc --batch 128 --niter 3 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 300 $SMALL > $OUT/plain_$TAG.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -s $SKIP -c $CNT -o $OUT/prof_$TAG $SMALL > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full exit $?"
tail -2 $OUT/ncu_full_$TAG.log

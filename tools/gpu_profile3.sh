#!/bin/bash
# full ncu capture of kernels matching a regex while running an arbitrary command.  bash tools/gpu_profile3.sh <tag> <regex> <skip> <count> <cmd...>
TAG=$1; REGEX=$2; SKIP=$3; CNT=$4; shift 4
OUT=gpurun_out
mkdir -p $OUT
timeout 300 "$@" > $OUT/plain_$TAG.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -s $SKIP -c $CNT -o $OUT/prof_$TAG "$@" > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full exit $?"
tail -2 $OUT/ncu_full_$TAG.log

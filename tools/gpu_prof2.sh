#!/bin/bash
# two full ncu captures of one kernel at different launch indices of the same command
#   bash tools/gpu_prof2.sh <tag> <kernel regex> <skipA> <skipB> <cmd...>
TAG=$1; REGEX=$2; SA=$3; SB=$4; shift 4
OUT=gpurun_out
mkdir -p $OUT
timeout 600 "$@" > $OUT/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain_$TAG.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -s $SA -c 1 -o $OUT/prof_${TAG}_a "$@" > $OUT/ncu_${TAG}_a.log 2>&1
echo "ncu a exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -s $SB -c 1 -o $OUT/prof_${TAG}_b "$@" > $OUT/ncu_${TAG}_b.log 2>&1
echo "ncu b exit $?"

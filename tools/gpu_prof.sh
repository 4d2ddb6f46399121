#!/bin/bash
# ncu visit: launch list (device time of every launch) + one full capture of the kernels matching a regex.
#   bash tools/gpu_prof.sh <tag> <kernel regex> <skip> <count> <cmd...>
TAG=$1; REGEX=$2; SKIP=$3; CNT=$4; shift 4
OUT=gpurun_out
mkdir -p $OUT
timeout 600 "$@" > $OUT/plain_$TAG.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches_$TAG.csv "$@" > $OUT/ncu_list_$TAG.log 2>&1
echo "ncu list exit $?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -s $SKIP -c $CNT -o $OUT/prof_$TAG "$@" > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full exit $?"
tail -2 $OUT/ncu_full_$TAG.log
python tools/summarize_launches.py $OUT/launches_$TAG.csv | head -40

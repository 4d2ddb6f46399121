#!/bin/bash
# ncu launch list + full captures of the top kernels on a short bench command.  bash tools/gpu_profile.sh <tag> <sampler>
TAG=${1:-r01}
SAMPLER=${2:-tc}
OUT=gpurun_out
mkdir -p $OUT
SMALL="python bench.py --sampler $SAMPLER --batch 128 --niter 3 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 300 $SMALL > $OUT/plain_$TAG.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches_$TAG.csv $SMALL > $OUT/ncu_list_$TAG.log 2>&1
echo "ncu list exit $?"
timeout 300 $SMALL > $OUT/plain2_$TAG.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"decoder_tc_kernel|decode_ws_kernel|nmf_hg2_kernel" -s 9 -c 3 -o $OUT/prof_$TAG $SMALL > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full exit $?"
tail -3 $OUT/ncu_full_$TAG.log
ls -la $OUT/*.ncu-rep

#!/bin/bash
# stage tests + M1 bench + per-kernel times of the M-step kernels (ncu launch list of a short B=512 run).  bash tools/gpu_wfs.sh <tag>
TAG=$1
OUT=gpurun_out
bash tools/gpu_try.sh $TAG "stages or e2e or tc"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"w_from_frame_stats|nmf_hg5|nmf_norm|decode_stats" -c 24 --csv --log-file $OUT/launches_$TAG.csv python bench.py --steps 1 --warmup 1 --niter 4 --no-cpu-baseline --no-e2e > $OUT/ncu_list_$TAG.log 2>&1
python tools/summarize_launches.py $OUT/launches_$TAG.csv | head -8

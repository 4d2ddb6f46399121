#!/bin/bash
# A/B of the sampler generations on the bench workloads + the configs[4] microbench sweep.  bash tools/gpu_ab.sh <tag>
TAG=${1:-ab}
OUT=gpurun_out
mkdir -p $OUT
Q="--steps 1 --warmup 2 --no-cpu-baseline --no-e2e"
for S in v2 v4 v3; do
  DVAE_TC_SAMPLER=$S timeout 300 python bench.py $Q > $OUT/ab_${TAG}_M1_$S.json 2> $OUT/ab_${TAG}_M1_$S.err
  python - <<P
import json
try:
    d = json.loads(open("$OUT/ab_${TAG}_M1_$S.json").read().strip().splitlines()[-1])
    print("M1 b512 $S", round(d["value"], 1), d["ms_per_step"], d["stage_share"], d["mean_final_cost"])
except Exception as e:
    print("M1 $S failed", e)
P
done
for S in v2 v4; do
  DVAE_TC_SAMPLER=$S timeout 600 python bench.py $Q --variant M2 --batch 4096 > $OUT/ab_${TAG}_M2_$S.json 2> $OUT/ab_${TAG}_M2_$S.err
  python - <<P
import json
try:
    d = json.loads(open("$OUT/ab_${TAG}_M2_$S.json").read().strip().splitlines()[-1])
    print("M2 b4096 $S", round(d["value"], 1), d["ms_per_step"], d["stage_share"], d["mean_final_cost"])
except Exception as e:
    print("M2 $S failed", e)
P
done
timeout 600 python tools/microbench_stft.py 1024 2048 4096 8192 16384 32768 65536 > $OUT/microbench_$TAG.jsonl 2> $OUT/microbench_$TAG.err
cut -c1-400 $OUT/microbench_$TAG.jsonl

#!/bin/bash
# tc tests + M1 bench with DVAE_TC_POLY=1 and =2 + phase clocks.  bash tools/gpu_try2.sh <tag>
TAG=$1
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "tc or e2e" > $OUT/pytest_$TAG.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_$TAG.log
tail -6 $OUT/pytest_$TAG.log
Q="--steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
show() { python - "$1" "$2" <<P
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], round(d["value"], 1), round(d["ms_per_step"], 1), d["stage_share"], d["roofline"]["frac"], d["mean_final_cost"])
except Exception as e:
    print(sys.argv[2], "failed", e)
P
}
for PM in 1 2; do
  DVAE_TC_POLY=$PM timeout 300 python bench.py $Q > $OUT/try_${TAG}_M1_p$PM.json 2> $OUT/try_${TAG}_M1_p$PM.err; show $OUT/try_${TAG}_M1_p$PM.json "M1 b512 poly=$PM"
done
python tools/tc_phase_clocks.py 512 > $OUT/phase_$TAG.log 2>&1; cat $OUT/phase_$TAG.log
DVAE_TC_POLY=2 python tools/tc_phase_clocks.py 512 > $OUT/phase_${TAG}_p2.log 2>&1; cat $OUT/phase_${TAG}_p2.log

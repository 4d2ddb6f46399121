#!/bin/bash
# Quick check of a kernel change: a pytest selection, one short M1 bench line (and optionally M2 b4096), microbench at 4096.
# bash tools/gpu_try.sh <tag> <pytest -k expr> [m2]
TAG=$1; KEXPR=$2; M2=$3
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "$KEXPR" > $OUT/pytest_$TAG.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_$TAG.log
tail -8 $OUT/pytest_$TAG.log
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
timeout 300 python bench.py $Q > $OUT/try_${TAG}_M1.json 2> $OUT/try_${TAG}_M1.err; show $OUT/try_${TAG}_M1.json "M1 b512"
if [ -n "$M2" ]; then
  timeout 600 python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-e2e --variant M2 --batch 4096 > $OUT/try_${TAG}_M2.json 2> $OUT/try_${TAG}_M2.err; show $OUT/try_${TAG}_M2.json "M2 b4096"
fi
timeout 300 python tools/microbench_stft.py 4096 > $OUT/microbench_$TAG.jsonl 2> $OUT/microbench_$TAG.err
cat $OUT/microbench_$TAG.jsonl; tail -3 $OUT/microbench_$TAG.err

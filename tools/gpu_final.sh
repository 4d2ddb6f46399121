#!/bin/bash
# Round-end GPU visit: parity suite, smoke, default bench line, reference arm, launch list (B=64, full schedule), ncu captures of the
# three dominant kernels, the other configs' bench lines and the configs[4] microbench.  bash tools/gpu_final.sh <tag>
TAG=${1:-final}
OUT=gpurun_out
bash tools/gpu_round.sh $TAG tc
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "reference arm exit $?"; cut -c1-300 $OUT/bench_ref_$TAG.json
bash tools/gpu_profile3.sh ${TAG}_mh2 mh2_kernel 1 1 python tools/tc_phase_clocks.py 512
bash tools/gpu_profile3.sh ${TAG}_ds "decode_stats_kernel|nmf_hg5_kernel" 2 2 python tools/ws_phase_clocks.py 512
for V in M2 M2v3; do
  timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --variant $V --batch 4096 > $OUT/bench_${V}_$TAG.json 2> $OUT/bench_${V}_$TAG.err
  echo "$V exit $?"; cut -c1-200 $OUT/bench_${V}_$TAG.json
done
timeout 600 python tools/microbench_stft.py 1024 4096 16384 65536 > $OUT/microbench_$TAG.jsonl 2> $OUT/microbench_$TAG.err
cut -c1-300 $OUT/microbench_$TAG.jsonl

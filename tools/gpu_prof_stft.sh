#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python tools/microbench_stft.py 4096 > $OUT/plain_stft.log 2>&1 || { echo plain failed; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^stft_kernel" -s 2 -c 1 -o $OUT/prof_stft_v2 python tools/microbench_stft.py 4096 > $OUT/ncu_stft_v2.log 2>&1; echo "ncu stft $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'istft_kernel' -s 2 -c 1 -o $OUT/prof_istft_v2 python tools/microbench_stft.py 4096 > $OUT/ncu_istft_v2.log 2>&1; echo "ncu istft $?"

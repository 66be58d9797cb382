#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python tools/tune.py all > gpurun_out/tune1.txt 2>&1
grep "== best" gpurun_out/tune1.txt
# ncu: full capture of the two kernels of the current 2^20 f64 plan (1 launch each, after warm-up)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft_pass_kernel -s 8 -c 2 -o gpurun_out/prof_r01_2p20_v0 \
   python bench.py --steps 32 --warmup 4 --no-graph --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log

#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/cufft_yardstick.py > gpurun_out/cufft.txt 2>&1; cat gpurun_out/cufft.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fft_pass_kernel -s 6 -c 3 -o gpurun_out/prof_r01_2p26_v0 \
   python bench.py --workload c2c_f64_2p26 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_2p26.log 2>&1
tail -2 gpurun_out/ncu_2p26.log | cut -c1-600

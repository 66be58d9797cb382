#!/bin/bash
# ncu full capture (with source) of the 1024-row middle pass, default vs radix-32 two-stage build
set -u
mkdir -p gpurun_out
export PHASTFT_FACTORS="24:7,10,7"
PHASTFT_PASS_C="16,8,16" timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft_pass_kernel -s 4 -c 1 -f -o gpurun_out/r01_mid1024_default python tools/one_fft.py f64 24 3 > gpurun_out/ncu_mid_def.log 2>&1
PHASTFT_PASS_VARIANT="0,32,0" PHASTFT_PASS_C="16,8,16" timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft_pass_kernel -s 4 -c 1 -f -o gpurun_out/r01_mid1024_r32 python tools/one_fft.py f64 24 3 > gpurun_out/ncu_mid_r32.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -2 gpurun_out/ncu_mid_def.log gpurun_out/ncu_mid_r32.log

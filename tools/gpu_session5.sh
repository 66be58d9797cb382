#!/bin/bash
set -u
mkdir -p gpurun_out
SECS="--section LaunchStats --section Occupancy --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section SchedulerStats --section ComputeWorkloadAnalysis"
i=0
for cfg in "24:8,8,8" "24:7,9,8" "24:6,10,8"; do
  i=$((i+1))
  PHASTFT_FACTORS=$cfg PHASTFT_PASS_C=16,8,16 timeout 600 ncu $SECS --clock-control none -k regex:fft_pass_kernel -s 3 -c 3 -o /tmp/prof_mid_$i python tools/one_fft.py f64 24 2 > gpurun_out/ncu_mid_$i.log 2>&1
  tail -1 gpurun_out/ncu_mid_$i.log
  ncu -i /tmp/prof_mid_$i.ncu-rep --page raw --csv > gpurun_out/prof_mid_$i.csv 2>/dev/null
  ls -la /tmp/prof_mid_$i.ncu-rep
done

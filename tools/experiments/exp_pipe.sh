#!/bin/bash
run() { env "$@" PHASTFT_CLUSTER=0 python tools/timing.py $SFX $LN $B "$*" 2>&1 | tail -1 | cut -c1-175; }
SFX=f32; LN=16; B=4096
run PHASTFT_PIPE=0
run PHASTFT_PIPE=1 PHASTFT_PIPE_DISCARD=0
run PHASTFT_PIPE=1 PHASTFT_PIPE_DISCARD=1
run PHASTFT_PIPE=1 PHASTFT_PIPE_RING_MB=16
run PHASTFT_PIPE=1 PHASTFT_PIPE_RING_MB=48
for cfg in "f64 16 2048" "f64 14 8192" "f64 18 512" "f64 20 128" "f32 14 8192" "f32 18 512" "f32 20 128"; do
  set -- $cfg; SFX=$1; LN=$2; B=$3
  run PHASTFT_PIPE=0
  run PHASTFT_PIPE=1
done
mkdir -p /tmp/ncu
for D in 0 1; do
PHASTFT_PIPE=1 PHASTFT_PIPE_DISCARD=$D timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,launch__registers_per_thread --clock-control none -k regex:fft_pipe2 -s 1 -c 1 python tools/one_fft.py f32 16 2 1024 2>&1 | grep -E "fft_pipe2|duration|dram__|issue_active|inst_executed|registers" | cut -c1-150
done

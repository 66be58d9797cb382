#!/bin/bash
# round-2 experiment: TMA tile input for the passes of 3-pass plans and for the pipelined batch launch
run() { env "$@" python tools/timing.py $SFX $LN $B "$*" 2>&1 | tail -1 | cut -c1-200; }
B=1
for cfg in "f64 26" "f64 24" "f32 26" "f32 24" "f64 22"; do
  set -- $cfg; SFX=$1; LN=$2
  run PHASTFT_TMA_MID=0
  run PHASTFT_TMA_MID=1
  run PHASTFT_TMA_MID=1 PHASTFT_TMA_ENDS=1
done
SFX=f32; LN=16; B=4096
run PHASTFT_PIPE=0
run PHASTFT_PIPE=1
run PHASTFT_PIPE=1 PHASTFT_PIPE_TMA=1
SFX=f64; LN=16; B=2048
run PHASTFT_PIPE=0
run PHASTFT_PIPE=1
run PHASTFT_PIPE=1 PHASTFT_PIPE_TMA=1

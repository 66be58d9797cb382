#!/bin/bash
# round-2 experiment: TMA tile input for the middle pass of 3-pass plans
run() { env "$@" python tools/timing.py $SFX $LN 1 "$*" 2>&1 | tail -1 | cut -c1-330; }
for cfg in "f64 26" "f64 25" "f32 26" "f32 25" "f64 27"; do
  set -- $cfg; SFX=$1; LN=$2
  run PHASTFT_TMA_MID=0
  run PHASTFT_TMA_MID=1
done

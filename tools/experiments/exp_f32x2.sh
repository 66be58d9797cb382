#!/bin/bash
# round-2 experiment: packed f32x2 arithmetic (FADD2/FMUL2/FFMA2) in the f32 passes
run() { env "$@" python tools/timing.py $SFX $LN $B "$*" 2>&1 | tail -1 | cut -c1-175; }
SFX=f32; LN=16; B=4096
run PHASTFT_PIPE=0
run PHASTFT_PIPE=1
for cfg in "f32 10 16384" "f32 12 4096" "f32 14 1024" "f32 18 512" "f32 20 128" "f32 20 1" "f32 16 1" "f32 24 1" "f64 20 1"; do
  set -- $cfg; SFX=$1; LN=$2; B=$3
  run PHASTFT_PIPE=0
  if [ $B -gt 1 ] && [ $LN -ge 13 ]; then run PHASTFT_PIPE=1; fi
done

#!/bin/bash
# round-2 experiment: TMA tile input for the 128- / 256-row middle passes of 2^21..2^24 (PHASTFT_TMA_MID=2 adds the id-310 tiles)
run() { env "$@" python tools/timing.py $SFX $LN 1 "$*" 2>&1 | tail -1 | cut -c1-260; }
for SFX in f64 f32; do
  for LN in 21 22 23 24; do
    run PHASTFT_TMA_MID=1; run PHASTFT_TMA_MID=2
  done
done

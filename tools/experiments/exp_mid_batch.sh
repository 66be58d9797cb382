#!/bin/bash
# round-2 experiment: (a) 2^27 with one 512-row pass instead of two; (b) batches of 2^17..2^21 with narrower tiles (more CTAs per SM)
run() { env "$@" python tools/timing.py $SFX $LN $B "$*" 2>&1 | tail -1 | cut -c1-250; }
for SFX in f64 f32; do
  LN=27; B=1
  run X=default; run PHASTFT_FACTORS="27:8,10,9"; run PHASTFT_FACTORS="27:9,10,8"; run PHASTFT_FACTORS="27:8,9,10"
  for LN in 17 18 19 20 21; do
    B=$(( (1<<24) >> LN ))
    run X=default
    if [ $SFX = f64 ]; then CS="4 8 16"; else CS="8 16 32"; fi
    for C in $CS; do run PHASTFT_TILE_C=$C; done
  done
done

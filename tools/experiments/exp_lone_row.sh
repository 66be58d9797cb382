#!/bin/bash
# round-2 experiment, after the twiddle change: lone transforms of 2^11..2^13 points in ONE CTA (one launch) vs two many-CTA passes
run() { env "$@" python tools/timing.py $SFX $LN 1 "$*" 2>&1 | tail -1 | cut -c1-200; }
for SFX in f64 f32; do
  for LN in 11 12 13; do
    run X=default
    for V in 0 70 81 90 91; do run PHASTFT_FACTORS="$LN:$LN" PHASTFT_VARIANT=$V; done
  done
done

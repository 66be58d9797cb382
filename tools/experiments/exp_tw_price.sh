#!/bin/bash
# round-2 experiment: coalesced [i][m] stage-twiddle tables in the one-CTA kernels (batches of 2^24 points, and lone transforms)
run() { env "$@" python tools/timing.py $SFX $LN $B "$*" 2>&1 | tail -1 | cut -c1-200; }
for SFX in f64 f32; do for LN in 5 6 7 8 9 10 11 12; do B=$(( (1<<24) >> LN )); run X=0; done; done
for SFX in f64 f32; do for LN in 8 10 12; do B=1; run X=0; done; done

#!/bin/bash
# round-2 experiment, after the [i][m] / product stage twiddles: do the 128 KB one-CTA kernels (2^13 f64, 2^13-2^14 f32) now beat
# the two-pass plan for batches?  And every compiled one-CTA variant again at 2^9..2^12.
run() { env "$@" python tools/timing.py $SFX $LN $B "$*" 2>&1 | tail -1 | cut -c1-230; }
for SFX in f64 f32; do
  for LN in 13 14; do
    B=$(( (1<<24) >> LN ))
    run X=two-pass
    for V in 0 90 91; do run PHASTFT_ROW_VARIANT=$V PHASTFT_ONE_CTA_MAX=14; done
  done
  for LN in 9 10 11 12; do
    B=$(( (1<<24) >> LN ))
    for V in 0 70 81; do run PHASTFT_ROW_VARIANT=$V; done
  done
done

#!/bin/bash
# round-2 experiment: one-CTA batch kernels 2^9..2^13, every compiled variant (packed f32 arithmetic may have moved the optimum)
run() { env "$@" python tools/timing.py $SFX $LN $B "$*" 2>&1 | tail -1 | cut -c1-260; }
for SFX in f32 f64; do
  for LN in 9 10 11 12 13; do
    B=$(( (1<<24) >> LN ))
    for V in 0 70 81 90 91; do run PHASTFT_ROW_VARIANT=$V PHASTFT_ONE_CTA_MAX=13; done
  done
done

#!/bin/bash
# round-2 experiment: batches of 16..128-point transforms, more transforms per CTA (variants 82 / 83) vs the current choices
run() { env "$@" python tools/timing.py $SFX $LN $B "$*" 2>&1 | tail -1 | cut -c1-170; }
for SFX in f64 f32; do
  for LN in 4 5 6 7; do
    B=$(( (1<<24) >> LN ))
    run X=default
    for V in 0 80 81 82 83; do run PHASTFT_ROW_VARIANT=$V; done
  done
done

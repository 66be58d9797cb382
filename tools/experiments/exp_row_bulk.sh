#!/bin/bash
# round-2 experiment: one-CTA batch kernels, tile in / out by cp.async.bulk (PHASTFT_ROW_BULK=1; it was the default build when this log was taken, label X=bulk) vs per-lane loads / stores (PHASTFT_ROW_BULK=0)
run() { env "$@" python tools/timing.py $SFX $LN $B "$*" 2>&1 | tail -1 | cut -c1-140; }
for SFX in f64 f32; do
  for LN in 2 3 4 5 6 7 8 9 10 11 12; do
    B=$(( (1<<24) >> LN ))
    run PHASTFT_ROW_BULK=1; run PHASTFT_ROW_BULK=0
  done
done

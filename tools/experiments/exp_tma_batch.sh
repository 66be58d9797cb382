#!/bin/bash
# round-2 experiment: batches of 2^19 / 2^20 points with the asynchronous-input (TMA / bulk) 1024- and 512-row tiles
run() { env "$@" python tools/timing.py $SFX $LN $B "$*" 2>&1 | tail -1 | cut -c1-250; }
for SFX in f64 f32; do
  for LN in 18 19 20; do
    B=$(( (1<<24) >> LN ))
    run X=default
    run PHASTFT_TMA=1 PHASTFT_TMA_BATCH=1
    run PHASTFT_TMA=1 PHASTFT_TMA_BATCH=1 PHASTFT_TMA_VARIANT=301
  done
done

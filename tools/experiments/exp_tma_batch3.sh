#!/bin/bash
# where does the asynchronous-input pair start to pay for small batches?  (default = TMA for batch * N >= 2^21)
run() { env "$@" python tools/timing.py $SFX $LN $B "$*" 2>&1 | tail -1 | cut -c1-110; }
for cfg in "f64 20 2" "f64 20 4" "f64 20 8" "f64 18 8" "f64 18 16" "f64 16 32" "f64 16 64" "f32 20 2" "f32 20 4" "f32 20 8" "f32 16 32" "f32 16 64" "f32 16 128"; do
  set -- $cfg; SFX=$1; LN=$2; B=$3
  run X=tma-default; run PHASTFT_TMA_BATCH=0
done

#!/bin/bash
# round-2 experiment: lone 2^18..2^20 transforms, plain vs asynchronous-input (TMA) kernels, with / without PDL
run() { env "$@" python tools/timing.py $SFX $LN 1 "$*" 2>&1 | tail -1 | cut -c1-400; }
for SFX in f64 f32; do
  for LN in 20 19 18; do
    run PHASTFT_TMA=0
    run PHASTFT_TMA=0 PHASTFT_PDL=1
    run PHASTFT_TMA=1 PHASTFT_PDL=0
    run PHASTFT_TMA=1
    run PHASTFT_TMA_VARIANT=301 PHASTFT_PDL=0
    run PHASTFT_TMA_VARIANT=301
  done
done

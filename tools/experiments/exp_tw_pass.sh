#!/bin/bash
# round-2 experiment: what the stage twiddles cost in the strided pass kernels (COL / TRANS), and whether deriving them by
# products pays there too.  Variants built by tools/experiments/build_variant.py:
#   notw   -DPHAST_EXP_STAGE_TW=1  no stage twiddles in COL/TRANS kernels (wrong results: the price of the loads + products)
#   twprod (at the time: -DPHAST_EXP_STAGE_TW=2; the default build since)  loads for i = 1,2,4,.. only, the rest by complex products
#   to reproduce "main" of this log now: build_variant.py tw_loads -DPHAST_EXP_STAGE_TW=2
run() { env "$@" python tools/timing.py $SFX $LN $B "$*" 2>&1 | tail -1 | cut -c1-160; }
for SFX in f64 f32; do
  for cfg in "20 1" "18 1" "22 1" "24 1" "26 1" "16 256" "14 1024"; do
    set -- $cfg; LN=$1; B=$2
    for V in main notw twprod; do
      if [ $V = main ]; then run V=main; else run V=$V PHASTFT_LIB=build/variants/$V/libphastft_cuda.so; fi
    done
  done
done

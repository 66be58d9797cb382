#!/bin/bash
# round-2 experiment: batched 2^13..2^16 transforms (2^24 points per call), cluster / one-CTA / two-launch plans
run() { env "$@" python tools/timing.py $SFX $LN $B "$*" 2>&1 | tail -1; }
for SFX in f64 f32; do
  for LN in 13 14 15 16; do
    B=$(( (1<<24) >> LN ))
    run PHASTFT_CLUSTER=0 PHASTFT_ONE_CTA_MAX=12
    run PHASTFT_CLUSTER=1
    if [ $SFX = f64 ]; then VAR="100"; else VAR="1 100"; fi
    for V in $VAR; do run PHASTFT_CLUSTER_VARIANT=$V PHASTFT_ONE_CTA_MAX=12; done
    run PHASTFT_ONE_CTA_MAX=12 PHASTFT_CLUSTER_VARIANT=0
  done
done
SFX=f32; LN=16; B=4096
run PHASTFT_CLUSTER=0
run PHASTFT_CLUSTER=1
run PHASTFT_CLUSTER_VARIANT=1
run PHASTFT_CLUSTER_VARIANT=100

#!/bin/bash
# round-2 experiment: one-CTA kernels, stage twiddles: 4 loads + 11 products per radix-16 task (batches of 2^24 points)
run() { env "$@" python tools/timing.py $SFX $LN $B "$*" 2>&1 | tail -1 | cut -c1-200; }
for SFX in f64 f32; do for LN in 8 9 10 11 12; do B=$(( (1<<24) >> LN )); run X=0; done; done

run() { env "$@" python tools/timing.py $SFX $LN $B "$*" 2>&1 | tail -1 | cut -c1-150; }
for SFX in f64 f32; do
  for LN in 16 17; do
    B=$(( (1<<24) >> LN ))
    run X=default
    run PHASTFT_TMA=1 PHASTFT_TMA_BATCH=1
  done
done
SFX=f32; LN=16; B=4096; run X=default; run PHASTFT_TMA=1 PHASTFT_TMA_BATCH=1

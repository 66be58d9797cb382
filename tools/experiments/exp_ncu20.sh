#!/bin/bash
mkdir -p /tmp/ncu gpurun_out
for cfg in "plain PHASTFT_TMA=0" "tma300 PHASTFT_TMA=1" ; do
  set -- $cfg
  env $2 PHASTFT_PDL=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:fft_pass -s 8 -c 2 -f -o /tmp/ncu/p20_$1 python tools/one_fft.py f64 20 8 > gpurun_out/ncu_$1.log 2>&1
  ncu -i /tmp/ncu/p20_$1.ncu-rep --page raw --csv > gpurun_out/r02_ncu_2p20_$1.raw.csv 2>/dev/null
  ncu -i /tmp/ncu/p20_$1.ncu-rep --page source --csv --print-source sass > gpurun_out/r02_ncu_2p20_$1.source.csv 2>/dev/null
  tail -2 gpurun_out/ncu_$1.log
done

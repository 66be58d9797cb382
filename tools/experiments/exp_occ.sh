#!/bin/bash
for cfg in "f32 16 0" "f32 16 1" "f32 16 100" "f32 15 0" "f64 14 0" "f64 15 0" "f64 16 0" "f64 16 100"; do
  set -- $cfg
  PHASTFT_CLUSTER_VARIANT=$3 PHASTFT_ONE_CTA_MAX=12 python - <<PY
import phastft_b200 as pf
P = pf.PlannerDit64 if "$1" == "f64" else pf.PlannerDit32
print(P(1 << $2, 0).describe().split("||")[-1])
PY
done
mkdir -p /tmp/ncu
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fft_cluster2 -s 2 -c 1 -f -o /tmp/ncu/cl_f32_16 python tools/one_fft.py f32 16 3 256 > gpurun_out/ncu_cl.log 2>&1
ncu -i /tmp/ncu/cl_f32_16.ncu-rep --page raw --csv > gpurun_out/r02_ncu_cluster_f32_2p16.raw.csv 2>/dev/null
ncu -i /tmp/ncu/cl_f32_16.ncu-rep --page details > gpurun_out/r02_ncu_cluster_f32_2p16.details.txt 2>/dev/null
ncu -i /tmp/ncu/cl_f32_16.ncu-rep --page source --csv --print-source sass > gpurun_out/r02_ncu_cluster_f32_2p16.source.csv 2>/dev/null
tail -3 gpurun_out/ncu_cl.log

#!/bin/bash
set -u
mkdir -p gpurun_out
bash tools/gpu_session6.sh
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200 --csv --log-file gpurun_out/r01_launches_2p20.csv python bench.py --steps 96 --warmup 16 --no-graph --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/r01_launches_2p20.csv | cut -c1-250
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft_pass_kernel -s 8 -c 2 -o gpurun_out/r01_prof_2p20_final python bench.py --steps 32 --warmup 4 --no-graph --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out/r01_prof_2p20_final.ncu-rep

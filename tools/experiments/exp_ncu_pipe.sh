#!/bin/bash
mkdir -p /tmp/ncu gpurun_out
PHASTFT_CLUSTER=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:fft_pipe2 -s 1 -c 1 -f -o /tmp/ncu/pipe python tools/one_fft.py f32 16 2 1024 > gpurun_out/ncu_pipe.log 2>&1
ncu -i /tmp/ncu/pipe.ncu-rep --page raw --csv > gpurun_out/r02_ncu_pipe_f32_2p16.raw.csv 2>/dev/null
ncu -i /tmp/ncu/pipe.ncu-rep --page source --csv --print-source sass > gpurun_out/r02_ncu_pipe_f32_2p16.source.csv 2>/dev/null
tail -2 gpurun_out/ncu_pipe.log
PHASTFT_CLUSTER=0 PHASTFT_PIPE=0 timeout 400 ncu --set full --clock-control none -k regex:fft_pass -s 2 -c 2 -f -o /tmp/ncu/two python tools/one_fft.py f32 16 2 1024 > gpurun_out/ncu_two.log 2>&1
ncu -i /tmp/ncu/two.ncu-rep --page raw --csv > gpurun_out/r02_ncu_two_f32_2p16.raw.csv 2>/dev/null

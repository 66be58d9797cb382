#!/bin/bash
# final-kernel profiles: launch list of the default bench command, full capture of the 2^20 passes,
# section-limited capture of the 2^26 passes (a full set on 3 GiB of state replays for ~10 minutes).
# The .ncu-rep files embed the whole cubin (~45 MB each), so only CSV/text exports leave the box.
set -u
mkdir -p gpurun_out
O=/tmp/ncu_out; mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200 --csv --log-file gpurun_out/r01_launches_2p20.csv python bench.py --steps 96 --warmup 16 --no-graph --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/r01_launches_2p20.csv | cut -c1-250
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft_pass_kernel -s 8 -c 2 -f -o $O/p2p20 python tools/one_fft.py f64 20 8 > gpurun_out/ncu_full_2p20.log 2>&1
timeout 900 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section LaunchStats --section WarpStateStats --section SchedulerStats --section ComputeWorkloadAnalysis --clock-control none -k regex:fft_pass_kernel -s 3 -c 3 -f -o $O/p2p26 python tools/one_fft.py f64 26 2 > gpurun_out/ncu_2p26.log 2>&1
timeout 600 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section LaunchStats --section WarpStateStats --clock-control none -k regex:fft_pass_kernel -s 2 -c 2 -f -o $O/pbatch python tools/one_fft.py f32 16 2 4096 > gpurun_out/ncu_batch.log 2>&1
for r in p2p20 p2p26 pbatch; do
  ncu -i $O/$r.ncu-rep --page raw --csv > gpurun_out/r01_ncu_$r.raw.csv 2>/dev/null
  ncu -i $O/$r.ncu-rep --page details > gpurun_out/r01_ncu_$r.details.txt 2>/dev/null
done
ncu -i $O/p2p20.ncu-rep --page source --csv --print-source sass > gpurun_out/r01_ncu_p2p20.source.csv 2>/dev/null
ls -la gpurun_out/ | head -30; du -sh gpurun_out

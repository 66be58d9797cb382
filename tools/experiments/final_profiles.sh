#!/bin/bash
# round-2 final evidence: launch list of the default bench command, ncu captures of the final kernels (CSV exports only leave the box)
set -u
mkdir -p gpurun_out /tmp/ncu
O=/tmp/ncu
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200 --csv --log-file gpurun_out/r02_launches_2p20.csv python bench.py --steps 96 --warmup 16 --no-graph --no-cpu-baseline --no-batched > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/r02_launches_2p20.csv | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft_pass -s 8 -c 2 -f -o $O/p2p20 python tools/one_fft.py f64 20 8 > gpurun_out/ncu_full_2p20.log 2>&1
SECT="--section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section LaunchStats --section WarpStateStats --section SchedulerStats --section ComputeWorkloadAnalysis"
timeout 900 ncu $SECT --clock-control none -k regex:fft_pass -s 3 -c 3 -f -o $O/p2p26 python tools/one_fft.py f64 26 2 > gpurun_out/ncu_2p26.log 2>&1
timeout 600 ncu $SECT --clock-control none -k regex:fft_pass -s 2 -c 2 -f -o $O/pbatch python tools/one_fft.py f32 16 2 4096 > gpurun_out/ncu_batch.log 2>&1
timeout 600 ncu $SECT --clock-control none -k regex:"fft_pass|r2c_untangle|c2r_preprocess" -s 14 -c 7 -f -o $O/pr2c python - > gpurun_out/ncu_r2c.log 2>&1 <<'PY'
import torch, phastft_b200 as pf
n = 1 << 24
p = pf.PlannerR2c64(n, 0)
x = torch.rand(n, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
a = torch.empty(n // 2 + 1, dtype=torch.float64, device="cuda"); b = torch.empty_like(a)
for _ in range(4):
    pf.r2c_fft_f64_with_planner(x, a, b, p); pf.c2r_fft_f64_with_planner(a, b, y, p)
torch.cuda.synchronize()
PY
for r in p2p20:2p20 p2p26:2p26 pbatch:batch_f32 pr2c:r2c; do
  src=${r%%:*}; dst=${r##*:}
  ncu -i $O/$src.ncu-rep --page raw --csv > gpurun_out/r02_ncu_${dst}_final.raw.csv 2>/dev/null
  ncu -i $O/$src.ncu-rep --page details > gpurun_out/r02_ncu_${dst}_final.details.txt 2>/dev/null
done
ncu -i $O/p2p20.ncu-rep --page source --csv --print-source sass > gpurun_out/r02_ncu_2p20_final.source.csv 2>/dev/null
# (the cluster and pipelined launches were captured earlier in the round: profiles/r02_ncu_cluster_*.csv, r02_ncu_pipe_*.csv)
# DRAM bytes per launch from the captures above -> the table bench.py reads, then the bench lines themselves, the sweep and the
# criterion-shaped series
mkdir -p /tmp/prof && cp gpurun_out/r02_ncu_*_final.raw.csv profiles/ 2>/dev/null
python tools/ncu_traffic.py > gpurun_out/r02_traffic.json && cp gpurun_out/r02_traffic.json profiles/r02_traffic.json
python bench.py > gpurun_out/r02_bench_c2c_f64_2p20.json 2> gpurun_out/bench_2p20.err
for W in c2c_f64_2p26 batch_f32 r2c_f64_2p24; do python bench.py --workload $W --no-cpu-baseline > gpurun_out/r02_bench_$W.json 2> gpurun_out/bench_$W.err; done
python bench.py --impl reference > gpurun_out/r02_bench_reference.json 2> gpurun_out/bench_ref.err
python tools/sweep.py > gpurun_out/r02_sweep.txt 2> gpurun_out/sweep.err
python tools/sweep.py --criterion gpurun_out/criterion > gpurun_out/criterion.log 2>&1 && tar czf gpurun_out/r02_criterion.tar.gz -C gpurun_out criterion && rm -rf gpurun_out/criterion
ls -la gpurun_out | head -40; du -sh gpurun_out

#!/bin/bash
# round-2 experiment: c2r with the pre-processing folded into the first pass's loads (MODE_C2R_IN) vs the separate sweep
mkdir -p /tmp/ncu gpurun_out
export PYTHONPATH=$PWD
cat > /tmp/c2r_time.py <<'PY'
import os, sys, torch, phastft_b200 as pf
def t(dt, ln, fuse):
    os.environ["PHASTFT_C2R_FUSE"] = fuse
    n = 1 << ln
    P = pf.PlannerR2c64 if dt == torch.float64 else pf.PlannerR2c32
    f = pf.c2r_fft_f64_with_planner if dt == torch.float64 else pf.c2r_fft_f32_with_planner
    p = P(n, 0)
    nb = max(2, (512 << 20) // (n * 8))
    ins = [(torch.rand(n // 2 + 1, dtype=dt, device="cuda"), torch.rand(n // 2 + 1, dtype=dt, device="cuda")) for _ in range(nb)]
    y = torch.empty(n, dtype=dt, device="cuda")
    for a, b in ins[:2]: f(a, b, y, p)
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for a, b in ins: f(a, b, y, p)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / nb * 1e3)
    return best
for dt in (torch.float64, torch.float32):
    for ln in (14, 16, 18, 20, 21, 22, 23, 24, 25, 26):
        a, b = t(dt, ln, "1"), t(dt, ln, "0")
        print(f"c2r {str(dt)[6:]} 2^{ln}: fused {a:9.2f} us   separate sweep {b:9.2f} us   {b / a:5.2f}x", flush=True)
PY
python /tmp/c2r_time.py
cat > /tmp/c2r_one.py <<'PY'
import torch, phastft_b200 as pf
n = 1 << 24
p = pf.PlannerR2c64(n, 0)
y = torch.empty(n, dtype=torch.float64, device="cuda")
a = torch.rand(n // 2 + 1, dtype=torch.float64, device="cuda"); b = torch.rand(n // 2 + 1, dtype=torch.float64, device="cuda")
for _ in range(3): pf.c2r_fft_f64_with_planner(a, b, y, p)
torch.cuda.synchronize()
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fft_pass_async -s 1 -c 1 -f -o /tmp/ncu/c2r python /tmp/c2r_one.py > gpurun_out/ncu_c2r.log 2>&1
ncu -i /tmp/ncu/c2r.ncu-rep --page raw --csv > gpurun_out/r02_ncu_c2r_fused.raw.csv 2>/dev/null
ncu -i /tmp/ncu/c2r.ncu-rep --page source --csv --print-source sass > gpurun_out/r02_ncu_c2r_fused.source.csv 2>/dev/null
tail -2 gpurun_out/ncu_c2r.log

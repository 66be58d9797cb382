#!/usr/bin/env python
"""PDL on/off and R=1024 variants for the 2^20 headline (each config in a subprocess: PHASTFT_PDL is read once)."""
import os, subprocess, sys
for pdl in (0, 1):
    for n, fac, pv in ((20, "20:10,10", "0,0"), (20, "20:10,10", "20,20"), (20, "20:10,10", "25,25"), (20, "20:10,10", "1,1"),
                       (18, "18:9,9", "0,0"), (16, "16:8,8", "0,0"), (24, "24:8,8,8", "0,0,0"), (26, "26:8,10,8", "0,0,0")):
        env = dict(os.environ, PHASTFT_PDL=str(pdl), PHASTFT_FACTORS=fac, PHASTFT_PASS_VARIANT=pv)
        nbuf = 16 if n <= 20 else 1
        reps = 50 if n <= 20 else 4
        code = (f"import sys; sys.path.insert(0,'tools'); from tune import time_plan; ms,d=time_plan('f64',1<<{n},1,{nbuf},{reps});"
                f"print('pdl={pdl} 2^{n} {fac} v={pv}: %.2f us | %s' % (ms*1e3, d[:150]))")
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
        print(out.stdout.strip() or out.stderr[-400:], flush=True)

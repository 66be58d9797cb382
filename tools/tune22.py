import os, subprocess, sys
for fuse in (0, 1, 2):
    env = dict(os.environ, PHASTFT_FUSE=str(fuse))
    code = ("import sys; sys.path.insert(0,'tools'); from tune import time_plan\n"
            "for sfx,n in (('f64',20),('f64',18),('f64',16),('f64',13),('f32',20),('f32',16)):\n"
            "    ms,d=time_plan(sfx,1<<n,1,16,40)\n"
            f"    print('fuse={fuse}', sfx, '2^%d: %.2f us | %s' % (n, ms*1e3, d[-40:]), flush=True)\n")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(out.stdout.strip() or out.stderr[-800:], flush=True)

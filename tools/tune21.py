import os, subprocess, sys
for fuse in (0, 1):
    env = dict(os.environ, PHASTFT_FUSE=str(fuse))
    code = ("import sys; sys.path.insert(0,'tools'); from tune import time_plan\n"
            "for sfx in ('f64','f32'):\n"
            "    for n in range(11 if sfx=='f64' else 13, 21):\n"
            "        ms,d=time_plan(sfx,1<<n,1,16,40)\n"
            f"        print('fuse={fuse}', sfx, '2^%d: %.2f us %.1f Gpt/s | %s' % (n, ms*1e3, (1<<n)/ms/1e6, d[12:140]), flush=True)\n")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(out.stdout.strip() or out.stderr[-800:], flush=True)

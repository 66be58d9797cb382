import os, subprocess, sys
def run(sfx, n, batch, mb, reps=3, extra=None):
    env = dict(os.environ, PHASTFT_L2_CHUNK_MB=str(mb))
    if extra: env.update(extra)
    code = (f"import sys; sys.path.insert(0,'tools'); from tune import time_plan; ms,d=time_plan('{sfx}',1<<{n},{batch},1,{reps});"
            f"print('{sfx} {batch}x2^{n} chunk {mb} MiB: %.1f us total, %.2f us/transform, %.1f Gpt/s | %s' % (ms*1e3, ms*1e3/{batch}, {batch}*(1<<{n})/ms/1e6, d[:110]))")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(out.stdout.strip() or out.stderr[-300:], flush=True)
for mb in (48, 96, 192, 384, 768, 4096):
    run("f32", 16, 4096, mb)
for mb in (48, 96, 192, 384, 1024):
    run("f64", 20, 64, mb)
for mb in (48, 192, 1024):
    run("f64", 16, 1024, mb)
    run("f64", 13, 8192, mb)

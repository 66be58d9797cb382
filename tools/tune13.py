import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from tune import time_plan
for sfx, n, nbuf, reps in (("f64", 20, 16, 60), ("f64", 18, 16, 60), ("f64", 16, 16, 60), ("f64", 14, 16, 60), ("f64", 24, 1, 5), ("f64", 26, 1, 3), ("f32", 20, 16, 60), ("f32", 16, 16, 60), ("f32", 24, 1, 5)):
    ms, d = time_plan(sfx, 1 << n, 1, nbuf, reps)
    print(f"{sfx} 2^{n}: {ms*1e3:8.2f} us {(1<<n)/ms/1e6:6.1f} Gpt/s | {d[:170]}", flush=True)
ms, d = time_plan("f32", 1 << 16, 4096, 1, 3)
print(f"f32 4096x2^16: {ms*1e3:8.1f} us {4096*65536/ms/1e6:6.1f} Gpt/s", flush=True)

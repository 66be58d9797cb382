import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from tune import time_plan
import tune4_lib as t4
def setenv(env):
    for k in ("PHASTFT_FACTORS", "PHASTFT_TILE_C", "PHASTFT_PASS_C", "PHASTFT_PASS_VARIANT", "PHASTFT_VARIANT"):
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in env.items()})
for sfx in ("f64", "f32"):
    for n, fac in ((16, "16:8,8"), (15, "15:7,8"), (17, "17:9,8"), (17, "17:8,9")):
        for pv in ("0,0", "40,40", "41,41", "42,42"):
            setenv({"PHASTFT_FACTORS": fac, "PHASTFT_PASS_VARIANT": pv})
            ms, d = time_plan(sfx, 1 << n, 1, 16, 50)
            print(f"{sfx} 2^{n} {fac} v={pv}: {ms*1e3:7.2f} us | {d[:140]}", flush=True)
F = "PHASTFT_FACTORS"; PC = "PHASTFT_PASS_C"; PV = "PHASTFT_PASS_VARIANT"
t4.prof("f64", 24, {F: "24:8,8,8", PC: "16,8,16", PV: "40,41,42"}, reps=5)
t4.prof("f64", 24, {F: "24:8,8,8", PC: "16,8,16", PV: "40,41,40"}, reps=5)
t4.prof("f64", 26, {F: "26:8,10,8", PC: "16,8,16", PV: "40,0,42"}, reps=3)
t4.prof("f64", 26, {F: "26:8,10,8", PC: "16,8,16", PV: "40,0,40"}, reps=3)
t4.prof("f64", 23, {F: "23:8,7,8", PC: "16,8,16", PV: "40,0,40"}, reps=5)
t4.prof("f64", 22, {F: "22:7,8,7", PC: "16,8,16", PV: "0,41,0"}, reps=5)
t4.prof("f64", 22, {F: "22:8,6,8", PC: "16,8,16", PV: "40,0,40"}, reps=5)
t4.prof("f32", 24, {F: "24:8,8,8", PC: "32,16,32", PV: "40,40,40"}, reps=5)
t4.prof("f32", 26, {F: "26:8,10,8", PC: "32,16,32", PV: "40,0,40"}, reps=3)

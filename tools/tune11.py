import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
import tune4_lib as t4
F = "PHASTFT_FACTORS"; PC = "PHASTFT_PASS_C"; PV = "PHASTFT_PASS_VARIANT"
for v, c in ((0, 8), (50, 8), (51, 8), (52, 8), (53, 8), (58, 8), (54, 4), (55, 4), (56, 4), (57, 4)):
    t4.prof("f64", 24, {F: "24:6,10,8", PC: f"16,{c},16", PV: f"0,{v},40"}, reps=5)
for v, c in ((0, 16), (50, 16), (53, 16), (58, 16), (54, 8), (55, 8), (56, 8)):
    t4.prof("f32", 24, {F: "24:6,10,8", PC: f"32,{c},32", PV: f"0,{v},40"}, reps=5)

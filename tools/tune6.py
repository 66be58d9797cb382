import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
import tune4_lib as t4
F = "PHASTFT_FACTORS"; PC = "PHASTFT_PASS_C"; PV = "PHASTFT_PASS_VARIANT"
VS = (0, 2, 3, 9, 11, 4)
print("### f64 first pass (COL, large stride) and last pass (TRANS) variants, 2^24")
for fac, pc in (("24:8,8,8", "16,8,16"), ("24:6,10,8", "16,8,16"), ("24:7,9,8", "16,8,16"), ("24:9,7,8", "8,8,16"), ("24:8,7,9", "16,8,8")):
    for v in VS:
        t4.prof("f64", 24, {F: fac, PC: pc, PV: f"{v},{v},{v}"}, reps=5)
print("### f32 batch 4096 x 2^16")
for fac, pc in (("16:8,8", "32,32"), ("16:7,9", "32,16"), ("16:8,8", "16,16"), ("16:9,7", "16,32"), ("16:6,10", "32,16")):
    for v in VS:
        t4.prof("f32", 16, {F: fac, PC: pc, PV: f"{v},{v}"}, batch=4096, reps=3)
print("### f32 2^24 single")
for fac, pc in (("24:8,8,8", "32,16,32"), ("24:7,10,7", "32,16,32"), ("24:8,9,7", "32,16,32")):
    for v in VS:
        t4.prof("f32", 24, {F: fac, PC: pc, PV: f"{v},{v},{v}"}, reps=5)
print("### f64 2^20")
for fac, pc in (("20:10,10", "8,8"),):
    for v in (0, 2, 3, 9, 11, 4):
        t4.prof("f64", 20, {F: fac, PC: pc, PV: f"{v},{v}"}, reps=20)

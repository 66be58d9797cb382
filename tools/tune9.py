import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
import tune4_lib as t4
F = "PHASTFT_FACTORS"; PC = "PHASTFT_PASS_C"; PV = "PHASTFT_PASS_VARIANT"
for v in (0, 40, 41, 42, 43):
    t4.prof("f64", 24, {F: "24:8,8,8", PC: "16,8,16", PV: f"{v},{v},{v}"}, reps=5)
for v in (0, 40, 41, 42, 43):
    t4.prof("f32", 24, {F: "24:8,8,8", PC: "32,16,32", PV: f"{v},{v},{v}"}, reps=5)
for v in (0, 40, 41, 42, 43):
    t4.prof("f32", 16, {F: "16:8,8", PC: "32,16", PV: f"{v},{v}"}, batch=4096, reps=3)
for v in (0, 40, 41, 42):
    t4.prof("f32", 16, {F: "16:8,8", PC: "16,16", PV: f"{v},{v}"}, batch=4096, reps=3)
t4.prof("f32", 16, {F: "16:7,9", PC: "32,16"}, batch=4096, reps=3)

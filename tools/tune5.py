import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
os.environ.setdefault("X", "1")
import importlib
t4 = importlib.import_module("tune4_lib")
F = "PHASTFT_FACTORS"; PC = "PHASTFT_PASS_C"; PV = "PHASTFT_PASS_VARIANT"
for v in (0, 1, 2, 3, 4, 5, 6, 7, 9, 10, 11, 30, 31, 32):
    t4.prof("f64", 24, {F: "24:7,9,8", PC: "16,8,16", PV: f"0,{v},0"}, reps=6)
for v in (0, 20, 21, 22, 23, 24, 25, 26):
    t4.prof("f64", 24, {F: "24:6,10,8", PC: "16,8,16", PV: f"0,{v},0"}, reps=6)
for v in (0, 1, 2, 3, 9, 10, 11):
    t4.prof("f64", 24, {F: "24:8,8,8", PC: "16,8,16", PV: f"0,{v},0"}, reps=6)

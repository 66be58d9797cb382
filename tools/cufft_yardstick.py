#!/usr/bin/env python
"""cuFFT (via torch.fft) as an informal GPU yardstick -- NOT part of the product or of bench.py's
arms; interleaved complex layout, out-of-place.  Prints us per transform."""
import torch
dev = torch.device("cuda", 0)


def t(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


xs = [torch.randn(1 << 20, dtype=torch.complex128, device=dev) for _ in range(16)]
i = [0]
def f20():
    torch.fft.fft(xs[i[0] % 16]); i[0] += 1
us = t(f20, 400)
print(f"cuFFT Z2Z 2^20: {us:.2f} us  {(1<<20)/us/1e3:.1f} Gpt/s")
del xs
x = torch.randn(1 << 26, dtype=torch.complex128, device=dev)
us = t(lambda: torch.fft.fft(x), 5)
print(f"cuFFT Z2Z 2^26: {us:.2f} us  {(1<<26)/us/1e3:.1f} Gpt/s")
del x
x = torch.randn(4096, 1 << 16, dtype=torch.complex64, device=dev)
us = t(lambda: torch.fft.fft(x, dim=1), 5)
print(f"cuFFT C2C 4096x2^16: {us:.2f} us  {4096*(1<<16)/us/1e3:.1f} Gpt/s")
del x
x = torch.randn(1 << 24, dtype=torch.float64, device=dev)
us = t(lambda: torch.fft.rfft(x), 10)
print(f"cuFFT D2Z 2^24: {us:.2f} us  {(1<<24)/us/1e3:.1f} Greal-pt/s")

"""Sustained GEMM throughput with SM clock and board power sampled during the run (is the kernel power-capped?).
    [RBA_TC_DEBUG=..] python tools/gemm_power.py M N K [seconds]"""
import os
import sys
import threading
import time

import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rba_b200 import ops

M, N, K = (int(v) for v in sys.argv[1:4])
secs = float(sys.argv[4]) if len(sys.argv) > 4 else 1.5
dev = torch.device("cuda", 0)
a = torch.randn(M, K, device=dev)
w = torch.randn(N, K, device=dev) / K ** 0.5
ap, wp = ops.split_planes(a), ops.split_planes(w)
c = torch.empty(M, N, device=dev)
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
clk, pw, stop = [], [], threading.Event()


def sample():
    while not stop.is_set():
        clk.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
        pw.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
        stop.wait(0.05)


for _ in range(3):
    ops.gemm(ap, wp, out=c, backend=ops.RBA_GEMM_TC)
torch.cuda.synchronize()
t = threading.Thread(target=sample, daemon=True)
t.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 0
t0 = time.perf_counter()
e0.record()
while time.perf_counter() - t0 < secs:
    for _ in range(50):
        ops.gemm(ap, wp, out=c, backend=ops.RBA_GEMM_TC)
    n += 50
    torch.cuda.synchronize()
e1.record()
torch.cuda.synchronize()
stop.set()
t.join()
ms = e0.elapsed_time(e1) / n
half = len(clk) // 2
print(f"M={M} N={N} K={K} dbg={os.environ.get('RBA_TC_DEBUG', '0')}: {ms:.4f} ms  {2.0 * M * N * K / ms / 1e9:.0f} TF/s  "
      f"SM clock median(second half) {sorted(clk[half:])[len(clk[half:]) // 2]} MHz  power {sorted(pw[half:])[len(pw[half:]) // 2]:.0f} W")

"""profiles/<prefix>_ncu_kernels.md from a directory of .ncu-rep captures (tools/gpu_ncu_kernels.sh):
   python tools/summarize_ncu.py gpurun_out/ncuk > profiles/r2f_ncu_kernels.md"""
import csv
import glob
import io
import os
import subprocess
import sys

d = sys.argv[1]
M = {
    "time_us": "gpu__time_duration.sum",
    "dram_rd_MB": "dram__bytes_read.sum", "dram_wr_MB": "dram__bytes_write.sum",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "tensor_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "xu_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "issue_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "occ_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "regs": "launch__registers_per_thread",
}
print("| capture | kernel | time (us, under ncu) | DRAM read / write (MB) | achieved DRAM GB/s | DRAM % | tensor pipe % | XU % | issue % | occupancy % | regs |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for rep in sorted(glob.glob(os.path.join(d, "*.ncu-rep"))):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        continue
    hdr, units, val = rows[0], rows[1], rows[2]
    rec = dict(zip(hdr, val))
    un = dict(zip(hdr, units))

    def g(k):
        v = rec.get(M[k], "")
        try:
            return float(v.replace(",", ""))
        except ValueError:
            return float("nan")

    def to(k, target):   # normalise units
        v, u = g(k), un.get(M[k], "")
        scale = {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}.get(u, 1.0)
        return v * scale
    t, rd, wr = to("time_us", "us"), to("dram_rd_MB", "MB"), to("dram_wr_MB", "MB")
    name = rec.get("Kernel Name", "?")[:70]
    gbs = (rd + wr) / t * 1e3 if t == t and t > 0 else float("nan")
    print(f"| {os.path.basename(rep)[:-8]} | `{name}` | {t:.1f} | {rd:.1f} / {wr:.1f} | {gbs:.0f} | {g('dram_pct'):.1f} | {g('tensor_pct'):.1f} | {g('xu_pct'):.1f} | {g('issue_pct'):.1f} | {g('occ_pct'):.1f} | {g('regs'):.0f} |")

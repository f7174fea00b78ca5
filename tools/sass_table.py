"""Per-kernel SASS mnemonic table of librba_b200.so (cuobjdump -sass): evidence of which kernels are tcgen05 / TMA / TMEM
(UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, UBLKCP = bulk copy, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit)
and which run on the legacy pipes (HMMA = mma.sync, MUFU, FFMA).   python tools/sass_table.py > profiles/r2_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "rba_b200", "lib", "librba_b200.so")
TAGS = ["UTCHMMA", "UTMALDG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "HMMA", "MUFU", "FFMA", "FHFMA", "LDGSTS", "LDSM", "ATOM", "RED"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
fn, counts, total = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        fn = re.sub(r"^void ", "", fn)
        counts[fn] = collections.Counter()
        total[fn] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and fn:
        total[fn] += 1
        op = m.group(1)
        for t in TAGS:
            if op.startswith(t):
                counts[fn][t] += 1
print(f"SASS mnemonic counts per kernel, {os.path.relpath(LIB, ROOT)} (sm_100a)")
print(f"{'kernel':88s} {'instr':>7s} " + " ".join(f"{t:>7s}" for t in TAGS))
for fn, c in sorted(counts.items()):
    if total[fn] < 8:
        continue
    print(f"{fn[:88]:88s} {total[fn]:7d} " + " ".join(f"{c[t]:7d}" if c[t] else f"{'.':>7s}" for t in TAGS))

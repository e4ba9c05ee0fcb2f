"""Per-kernel SASS evidence of the in-tree library: which kernels use the Blackwell tensor / copy engines.
    python tools/sass_summary.py > profiles/r02_sass_summary.txt
Mnemonics (B200_PROFILING.md): UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = cp.async.bulk.tensor (TMA, tensor
map), UBLKCP = cp.async.bulk (TMA, linear), IMMA/HMMA = mma.sync (legacy warp-level tensor path), LDGSTS = cp.async."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "qwen3-tts-apple-silicon_b200", "qwen3_tts_b200", "libq3tts_b200.so")
PAT = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "IMMA", "HMMA", "LDGSTS", "LDSM", "SYNCS", "UCGABAR", "MAPA"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[cur] = collections.Counter()
        continue
    if cur is None or "/*" not in line:
        continue
    total[cur] += 1
    for p in PAT:
        if re.search(r"\b" + p + r"[\w.]*", line):
            counts[cur][p] += 1
print(f"# {os.path.relpath(LIB, ROOT)}  ({os.path.getsize(LIB)} bytes), cuobjdump -sass, instruction counts per kernel")
print(f"{'kernel':58s} {'instr':>7s}  " + " ".join(f"{p:>7s}" for p in PAT))
for k, c in counts.items():
    print(f"{k[:58]:58s} {total[k]:7d}  " + " ".join(f"{c.get(p, 0):7d}" for p in PAT))

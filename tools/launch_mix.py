"""Kernel mix of ONE batched frame from an ncu launch list (ncu --metrics gpu__time_duration.sum --csv of tools/bs64_probe.py):
the `n` launches up to the last advance_kernel before the codec = the last graph replay.  Times are cold-cache and serialised: shares, not absolutes.
usage: python tools/launch_mix.py gpurun_out/launches.csv <launches per frame>"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10 and r[0].isdigit()]
n = int(sys.argv[2])
names = [r[4] for r in rows]
first_codec = next(i for i, k in enumerate(names) if "rvq_gather_sum" in k)
last = max(i for i, k in enumerate(names[:first_codec]) if "advance_kernel" in k)      # a frame ends with the talker step's advance
frame = rows[last + 1 - n:last + 1]
agg = collections.OrderedDict()
for r in frame:
    key = (r[4].split("(")[0].replace("q3t::", ""), r[8] if "gemm" in r[4] or "attn" in r[4] else "")
    us = float(r[-1].replace(",", "")) / (1e3 if r[-2] in ("ns", "nsecond") else 1.0)
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1; a[1] += us
tot = sum(v[1] for v in agg.values())
print(f"{len(frame)} launches, {tot:.0f} us serialised")
for (k, g), (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:36s} {g:16s} x{c:4d} {us:8.0f} us  avg {us / c:6.1f}  {100 * us / tot:5.1f} %")

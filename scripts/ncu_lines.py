#!/usr/bin/env python
"""Warp-instructions and stall samples per CUDA source line of one kernel from an ncu report:
python scripts/ncu_lines.py rep kernel_regex [min_share_percent]"""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
minshare = float(sys.argv[3]) if len(sys.argv) > 3 else 0.7
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, seen_fn, hdr = None, 0, None
agg = collections.OrderedDict()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; ie = hdr.index("Instructions Executed"); iss = hdr.index("# Samples"); it = hdr.index("Thread Instructions Executed"); continue
    if hdr is None or len(r) < len(hdr) - 3:
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    if r[2] == "-":   # the CUDA line row itself carries the aggregate over its SASS
        key = (fname, line)
        a = agg.setdefault(key, [r[1].strip(), 0, 0, 0])
        a[1] += int(r[ie] or 0); a[2] += int(r[iss] or 0); a[3] += int(r[it] or 0)
tot_e = sum(v[1] for v in agg.values()); tot_s = sum(v[2] for v in agg.values()); tot_t = sum(v[3] for v in agg.values())
print(f"# {kern}: {tot_e} warp-instructions, {tot_t} thread-instructions ({tot_t / max(tot_e,1):.1f} thr/inst), {tot_s} samples")
for (f, l), (src, e, s, t) in agg.items():
    if e >= tot_e * minshare / 100 or s >= tot_s * minshare / 100:
        print(f"{f}:{l:4d} inst {100 * e / tot_e:5.1f}%  smp {100 * s / max(tot_s,1):5.1f}%  thr/inst {t / max(e,1):4.1f}  {src[:110]}")

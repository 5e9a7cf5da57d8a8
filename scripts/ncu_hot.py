#!/usr/bin/env python
"""Hot SASS lines of one kernel from an ncu report (source page): python scripts/ncu_hot.py rep kernel_regex [top]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# only the first kernel block
hdr = rows[1]
body = []
for r in rows[2:]:
    if len(r) < len(hdr) - 5:
        break
    body.append(r)
ia, ie, it, iss = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Avg. Predicated-On Threads Executed"), hdr.index("# Samples")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
si = [hdr.index(h) for h in stalls]
tot_s = sum(int(r[iss]) for r in body)
tot_e = sum(int(r[ie]) for r in body)
agg = {h: sum(int(r[i]) for r in body) for h, i in zip(stalls, si)}
print(f"# {kern}: {len(body)} SASS lines, {tot_e} warp-instructions, {tot_s} samples")
print("# stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v > tot_s * 0.01})
idx = sorted(range(len(body)), key=lambda k: -int(body[k][iss]))[:top]
for k in sorted(idx):
    r = body[k]
    st = sorted(((int(r[i]), h[6:]) for h, i in zip(stalls, si)), reverse=True)[:2]
    print(f"{k:5d} exe={int(r[ie]):10d} thr={r[it]:>5s} smp={int(r[iss]):6d} ({100 * int(r[iss]) / tot_s:4.1f}%) {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]}  {r[ia].strip()[:70]}")

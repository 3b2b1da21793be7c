"""Correlate an ncu SASS-level source page (ncu -i X.ncu-rep --page source --csv) with nvdisasm -g line info and print the
hottest source lines per kernel with their dominant stall reasons.
usage: ncu_lines.py <src_sass.csv> <kernel-substring> <file.sass> [top]"""
import csv
import re
import sys
from collections import defaultdict

csv_path, kname, sass_path = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
rows, hdr, on = [], None, False
for r in csv.reader(open(csv_path)):
    if r and r[0] == "Kernel Name":
        on = kname in r[1]
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if on and hdr and len(r) == len(hdr):
        rows.append(r)
ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
lines, cur, infn = [], None, False
pat_fn = re.compile(r"^\s*\.text\.(\S+):")
for ln in open(sass_path):
    m = pat_fn.match(ln)
    if m:
        infn = kname in m.group(1)
        continue
    if not infn:
        continue
    if "//## File" in ln:
        mm = re.search(r'File "([^"]+)", line (\d+)(.*)', ln)
        cur = (mm.group(1).split("/")[-1], int(mm.group(2)))
        continue
    if re.match(r"^\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
n = min(len(lines), len(rows))
print(f"kernel {kname}: {len(rows)} ncu rows, {len(lines)} sass instrs")
agg = defaultdict(lambda: [0, 0, 0, defaultdict(int)])
tot = [0, 0, 0, defaultdict(int)]
for i in range(n):
    r = rows[i]
    s = int(r[ci["Warp Stall Sampling (All Samples)"]] or 0); ie = int(r[ci["Instructions Executed"]] or 0); te = int(r[ci["Thread Instructions Executed"]] or 0)
    a = agg[lines[i]]
    a[0] += s; a[1] += ie; a[2] += te
    tot[0] += s; tot[1] += ie; tot[2] += te
    for st in stalls:
        v = int(r[ci[st]] or 0)
        if v:
            a[3][st] += v; tot[3][st] += v
print(f"total samples {tot[0]}  warp-instr {tot[1]}  thread-instr {tot[2]}  avg threads/instr {tot[2]/max(1,tot[1]):.1f}")
print("stall mix: " + "  ".join(f"{k[6:]} {100*v/max(1,tot[0]):.1f}%" for k, v in sorted(tot[3].items(), key=lambda kv: -kv[1])[:8]))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    mix = " ".join(f"{a[6:]}:{100*b/max(1,v[0]):.0f}" for a, b in sorted(v[3].items(), key=lambda kv: -kv[1])[:3])
    print(f"{str(k):36s} samples {v[0]:7d} ({100*v[0]/max(1,tot[0]):5.1f}%)  winstr {v[1]:10d} ({100*v[1]/max(1,tot[1]):5.1f}%)  thr/instr {v[2]/max(1,v[1]):5.1f}  {mix}")

"""Aggregate an ncu source page (ncu -i X.ncu-rep --page source --csv) per device function and per source line, using nvdisasm -g -c
of the same cubin for the function labels and line info.
usage: ncu_funcs.py <src.csv> <csv-kernel-substring> <file.sass> <mangled-kernel-substring> [top-lines]"""
import csv
import re
import sys
from collections import defaultdict

csv_path, kcsv, sass_path, ksass = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 25
rows, hdr, on, seen = [], None, False, False
for r in csv.reader(open(csv_path)):
    if r and r[0] == "Kernel Name":
        on = (kcsv in r[1]) and not seen
        if on:
            seen = True
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if on and hdr and len(r) == len(hdr):
        rows.append(r)
ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
funcs, lines, cur_fn, cur_ln, infn = [], [], None, None, False
for ln in open(sass_path):
    m = re.match(r"^\s*\.text\.(\S+):", ln)
    if m:
        infn = ksass in m.group(1)
        cur_fn = "kernel body"
        continue
    if not infn:
        continue
    m = re.match(r"^(\$\S+):", ln)
    if m:
        lab = m.group(1)
        mm = re.search(r"\d+(eval_single|eval_generic|eval_fd|lm_solve_dim|nn_scan_split|nn_scan|finalize_object|make_pair|sym_oplus|noinline_\w+?)I?[A-Z]", lab)
        cur_fn = mm.group(1) if mm else lab.split("$")[-1][:40]
        if "eval_fd" in cur_fn:
            cur_fn += "<%s>" % re.search(r"eval_fdILi\d+ELi(\d)", lab).group(1)
        continue
    if "//## File" in ln:
        mm = re.search(r'File "([^"]+)", line (\d+)', ln)
        cur_ln = (mm.group(1).split("/")[-1], int(mm.group(2)))
        continue
    if re.match(r"^\s+/\*[0-9a-f]{4,}\*/", ln):
        funcs.append(cur_fn); lines.append(cur_ln)
n = min(len(funcs), len(rows))
print(f"{len(rows)} ncu rows, {len(funcs)} sass instrs")
agg = defaultdict(lambda: [0, 0, 0, defaultdict(int), 0])
lagg = defaultdict(lambda: [0, 0, 0, defaultdict(int)])
tot = [0, 0, 0, defaultdict(int)]
for i in range(n):
    r = rows[i]
    s = int(r[ci["Warp Stall Sampling (All Samples)"]] or 0); ie = int(r[ci["Instructions Executed"]] or 0); te = int(r[ci["Thread Instructions Executed"]] or 0)
    a = agg[funcs[i]]; b = lagg[(funcs[i], lines[i])]
    a[0] += s; a[1] += ie; a[2] += te; a[4] += 1
    b[0] += s; b[1] += ie; b[2] += te
    tot[0] += s; tot[1] += ie; tot[2] += te
    for st in stalls:
        v = int(r[ci[st]] or 0)
        if v:
            a[3][st] += v; b[3][st] += v; tot[3][st] += v
print(f"total samples {tot[0]}  warp-instr {tot[1]}  thread-instr {tot[2]}  avg threads/instr {tot[2]/max(1,tot[1]):.1f}")
print("stall mix: " + "  ".join(f"{k[6:]} {100*v/max(1,tot[0]):.1f}%" for k, v in sorted(tot[3].items(), key=lambda kv: -kv[1])[:8]))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    mix = " ".join(f"{a[6:]}:{100*b/max(1,v[0]):.0f}" for a, b in sorted(v[3].items(), key=lambda kv: -kv[1])[:3])
    print(f"{k:28s} static {v[4]:5d}  samples {v[0]:7d} ({100*v[0]/max(1,tot[0]):5.1f}%)  winstr {v[1]:10d} ({100*v[1]/max(1,tot[1]):5.1f}%)  thr/instr {v[2]/max(1,v[1]):5.1f}  {mix}")
print("---- hottest lines")
for k, v in sorted(lagg.items(), key=lambda kv: -kv[1][0])[:top]:
    mix = " ".join(f"{a[6:]}:{100*b/max(1,v[0]):.0f}" for a, b in sorted(v[3].items(), key=lambda kv: -kv[1])[:3])
    print(f"{k[0][:18]:18s} {str(k[1]):28s} samples {v[0]:7d} ({100*v[0]/max(1,tot[0]):5.1f}%)  winstr {v[1]:10d} ({100*v[1]/max(1,tot[1]):5.1f}%)  thr/instr {v[2]/max(1,v[1]):5.1f}  {mix}")

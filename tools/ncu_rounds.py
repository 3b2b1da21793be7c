"""Per-round device times of the symmetry round pipeline from an ncu launch list (--metrics gpu__time_duration.sum --csv):
usage: ncu_rounds.py <launches.csv>   -> eval(r) / step(r) durations in launch order, totals per kernel"""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = {}
seq = []
for r in rows[1:]:
    name = r[ki].split("(")[0].replace("void ", "").replace("eslam::", "").replace("<unnamed>::", "")
    v = float(r[vi].replace(",", "")); u = r[ui]
    us = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
    tot.setdefault(name, [0, 0.0]); tot[name][0] += 1; tot[name][1] += us
    seq.append((name, us))
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:50]:50s} launches {n:4d}  total {t:10.1f} us")
ev = [u for n, u in seq if "sym_eval" in n]
stp = [u for n, u in seq if "sym_step" in n]
per = len(ev) // max(1, sum(1 for n, _ in seq if "sym_final" in n))
print("rounds per sequence", per)
print("eval us :", " ".join(f"{u:.0f}" for u in ev[:per]))
print("step us :", " ".join(f"{u:.0f}" for u in stp[:per + 1]))

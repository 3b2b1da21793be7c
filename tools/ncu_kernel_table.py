"""Per-kernel table (time, DRAM bytes) from an ncu CSV launch list:  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file X.csv ...
usage: ncu_kernel_table.py X.csv [frames]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
frames = float(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
acc = {}
for r in rows[hdr + 1:]:
    d = dict(zip(H, r))
    k = d["Kernel Name"].split("(")[0]
    acc.setdefault(k, {}).setdefault(d["Metric Name"], []).append(float(d["Metric Value"].replace(",", "")))
tot_t = tot_b = 0.0
print(f"{'kernel':36s} {'launches':>8s} {'ms':>9s} {'read GB':>9s} {'write GB':>9s}")
for k, v in acc.items():
    t = sum(v.get("gpu__time_duration.sum", [0])) / 1e6
    rd = sum(v.get("dram__bytes_read.sum", [0])) / 1e9; wr = sum(v.get("dram__bytes_write.sum", [0])) / 1e9
    tot_t += t; tot_b += rd + wr
    print(f"{k[:36]:36s} {len(v.get('gpu__time_duration.sum', [])):8d} {t:9.3f} {rd:9.3f} {wr:9.3f}")
print(f"{'total':36s} {'':8s} {tot_t:9.3f} {tot_b:9.3f} GB" + (f"  = {tot_b * 1e3 / frames:.1f} MB/frame" if frames else ""))

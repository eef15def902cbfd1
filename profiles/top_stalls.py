"""Summarise an `ncu --page source --csv` dump: hottest SASS lines by stall samples.
    ncu -i x.ncu-rep --page source --csv > src.csv ; python profiles/top_stalls.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[col["# Samples"]] or 0) for r in data)
inst = sum(int(r[col["Instructions Executed"]] or 0) for r in data)
print(f"total samples {tot}, warp instructions {inst}, SASS lines {len(data)}")
agg = {s: 0 for s in stall_cols}
for r in data:
    for s in stall_cols:
        agg[s] += int(r[col[s]] or 0)
print("stall mix:", ", ".join(f"{k[6:]}={v*100//max(tot,1)}%" for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v * 50 > tot))
idx = sorted(range(len(data)), key=lambda i: -int(data[i][col["# Samples"]] or 0))[:n]
for i in sorted(idx):
    r = data[i]
    top = sorted(((int(r[col[s]] or 0), s[6:]) for s in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {int(r[col['# Samples']]):7d} {int(r[col['Instructions Executed']]):10d}  {r[col['Source']].strip():60s} "
          + " ".join(f"{nm}:{v}" for v, nm in top if v))

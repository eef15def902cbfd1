"""Summarise `ncu --page source --csv --print-source cuda,sass`: stall samples, dominant stall
reasons and executed warp instructions per CUDA source line (rows that carry a line number are
the per-line totals).
    ncu -i x.ncu-rep --page source --csv --print-source cuda,sass > src.csv
    python profiles/by_line.py src.csv [N] [first_line last_line]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
lo = int(sys.argv[3]) if len(sys.argv) > 4 else 0
hi = int(sys.argv[4]) if len(sys.argv) > 4 else 10 ** 9
out = []
cur_file = ""
hdr = None
stall_cols = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = {}
        for i, h in enumerate(r):
            hdr.setdefault(h, i)
        stall_cols = [(h, i) for h, i in hdr.items() if h.startswith("stall_") and "Not Issued" not in h]
    elif hdr and len(r) > 8 and r[0].isdigit():
        try:
            smp = int(r[hdr["# Samples"]] or 0)
            ins = int(r[hdr["Instructions Executed"]] or 0)
        except ValueError:
            continue
        if (smp or ins) and lo <= int(r[0]) <= hi:
            st = sorted(((int(r[i] or 0), h[6:]) for h, i in stall_cols), reverse=True)[:2]
            out.append((smp, ins, cur_file, int(r[0]), r[1].strip()[:80], " ".join(f"{nm}:{v}" for v, nm in st if v)))
tot = sum(o[0] for o in out)
toti = sum(o[1] for o in out)
print(f"total samples {tot}, warp instructions {toti}")
for smp, ins, f, ln, src, st in sorted(out, reverse=True)[:n]:
    print(f"{smp*100/max(tot,1):5.1f}% {ins*100/max(toti,1):5.1f}%i {f}:{ln:4d}  {src:80s} {st}")

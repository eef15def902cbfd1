"""Small DecompressAll run for ncu captures (kept short: ncu replays every kernel ~40x).
    python profiles/prof_small.py [reads] [chunk] [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import corpus  # noqa: E402
import parallelparsing_b200 as pp  # noqa: E402

reads = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
fq = corpus.fastq(reads, fixed=150)
gz = corpus.gz_parallel(fq, 6, segment=8 << 20, threads=os.cpu_count())
ix = pp.Core.BuildDeflateIndex(gz, chunk)
dev = pp.Device(0)
job = pp.Job(dev, ix, gz.size)
import ctypes as C  # noqa: E402
L = pp.lib()
ph = (C.c_ulonglong * 20)()
L.pp_internal_phase_cycles(ph, 20)
for _ in range(iters):
    info = job.run(gz)
n = L.pp_internal_phase_cycles(ph, 20)
names = ["stage", "header", "guess", "sync", "scan", "emit", "resolve", "stored", "other", "r.expand", "r.expand+gather", "r.chase+store+scatter", "h.parse", "h.lit", "wait"]
tot = sum(ph[i] for i in range(n)) or 1
print("phase share of CTA time:", ", ".join(f"{names[i]} {100*ph[i]/tot:.1f}%" for i in range(n)),
      f"| cycles/iter/CTA-sum {tot/iters:.3e}")
assert info.status == 0
print(f"chunks {info.n_chunks} records {info.total_records} bytes {info.total_bytes} "
      f"inflate {info.inflate_ms:.3f} ms parse {info.parse_ms:.3f} ms")

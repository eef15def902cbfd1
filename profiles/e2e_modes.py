"""Where does an end-to-end step go?  Every upload mode on the bench file, wall clock per step and the
library's own CUDA-event times (upload / inflate incl. window pre-pass / parse).
    python profiles/e2e_modes.py [reads] [chunk] [steps]"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import parallelparsing_b200 as pp  # noqa: E402

reads = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
bench.make_native(["tools", os.path.join("parallelparsing_b200", "csrc")])
gz_path = bench.make_gz(reads, 150, 0, chunk)
idx_path = bench.corpus_paths(reads, 150, 0, chunk)[2]
if not os.path.exists(idx_path):
    pp.IndexIO.Serialize(pp.Core.BuildDeflateIndex(gz_path, chunk), idx_path)
ix = pp.IndexIO.Deserialize(idx_path)
gz_np = np.fromfile(gz_path, np.uint8)
gz, ptr = pp.pinned_copy(gz_np)
dev = pp.Device(0)
U = None
for name, kw, src in (("staged", {}, ptr), ("staged+compact", dict(compact_windows=True), ptr),
                      ("pipelined", dict(pipeline=True), ptr), ("pipelined+compact", dict(pipeline=True, compact_windows=True), ptr),
                      ("pipelined pageable", dict(pipeline=True), gz_np.ctypes.data_as(C.c_void_p)),
                      ("pull", dict(zero_copy=True), ptr), ("pull+compact", dict(zero_copy=True, compact_windows=True), ptr)):
    job = pp.Job(dev, ix, gz.size, **kw)
    for _ in range(2):
        job.upload(src); job.execute(); job.download()
    ph = (C.c_ulonglong * 20)()
    pp.lib().pp_internal_phase_cycles(ph, 20)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tu = te = td = 0.0
    for _ in range(steps):
        a = time.perf_counter(); job.upload(src)
        b = time.perf_counter(); job.execute()
        c = time.perf_counter(); job.download()
        d = time.perf_counter()
        tu += b - a; te += c - b; td += d - c
    dt = (time.perf_counter() - t0) / steps
    i = job.info()
    U = i.total_bytes
    print(f"{name:20s} step {dt*1e3:7.2f} ms = {U/dt/1e9:6.1f} GB/s | host: upload {tu/steps*1e3:5.2f} execute {te/steps*1e3:5.2f} "
          f"download {td/steps*1e3:6.2f} | events: upload {i.upload_ms:6.2f} inflate {i.inflate_ms:6.2f} parse {i.parse_ms:5.2f} "
          f"| h2d {i.h2d_bytes/1e6:7.1f} MB")
    nph = pp.lib().pp_internal_phase_cycles(ph, 20)
    names = ["stage", "header", "guess", "sync", "scan", "emit", "resolve", "stored", "other", "r.expand", "r.expand+gather", "r.chase+store", "h.parse", "h.lit", "wait"]
    tot = sum(ph[i] for i in range(nph)) or 1
    print("      CTA-time ms/step (sum over CTAs / 296 / 1.965 GHz): " + ", ".join(
        f"{names[i]} {ph[i]/steps/296/1.965e6:.2f}" for i in range(nph) if ph[i] * 200 > tot))
    job.free()

"""CreateIndex on the GPU (pp_index_create_gpu) beside the host pass it replaces (pp_index_create: zlib
inflate(Z_BLOCK), one thread, as Core.BuildDeflateIndex does) on bench.py's corpus: Generator seed 0,
N reads x 150 bp, gzip -6, chunk 10 000.  Checks that both indexes serialize to identical files, then
times `steps` GPU builds from pinned host memory (wall clock, everything included: H2D of the file, the
kernels, host stitching and planning, D2H of windows and offsets) and one host build.  One JSON line.
    python profiles/createindex_bench.py [reads] [steps]"""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
import parallelparsing_b200 as pp  # noqa: E402

reads = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
bench.make_native(["tools", os.path.join("parallelparsing_b200", "csrc")])
gz_path = bench.make_gz(reads, 150, 0, 10000)
pin, ptr = pp.pinned_copy(np.fromfile(gz_path, np.uint8))
dev = pp.Device(0)
t0 = time.perf_counter()
host = pp.Core.BuildDeflateIndex(gz_path, 10000)
host_s = time.perf_counter() - t0
ix, st = pp.Core.BuildDeflateIndexGpu(pin, 10000, dev, want_stats=True)     # warm-up + parity
with tempfile.TemporaryDirectory() as d:
    pp.IndexIO.Serialize(host, os.path.join(d, "h"))
    pp.IndexIO.Serialize(ix, os.path.join(d, "g"))
    same = open(os.path.join(d, "h"), "rb").read() == open(os.path.join(d, "g"), "rb").read()
assert same, "GPU index differs from the host index"
walls, stats = [], []
for _ in range(steps):
    t0 = time.perf_counter()
    ix, st = pp.Core.BuildDeflateIndexGpu(pin, 10000, dev, want_stats=True)
    walls.append(time.perf_counter() - t0)
    stats.append(st)
best = int(np.argmin(walls))
U = stats[best]["total_out"]
print(json.dumps({
    "metric": "CreateIndex uncompressed GB/s", "value": U / walls[best] / 1e9, "unit": "GB/s",
    "gpu_wall_ms": walls[best] * 1e3, "gpu_wall_ms_all": [w * 1e3 for w in walls],
    "host_zlib_1thread_s": host_s, "host_GBps": U / host_s / 1e9, "speedup": host_s / walls[best],
    "identical_index_file": same, "stages_ms": {k: v for k, v in stats[best].items() if k.endswith("_ms")},
    "blocks": stats[best]["blocks"], "segments": stats[best]["segments"], "points": stats[best]["points"],
    "scan_passes": stats[best]["scan_passes"],
    "config": {"workload": f"Generator seed 0, {reads} reads x 150bp, gzip -6, chunk 10000", "uncompressed_bytes": U,
               "compressed_bytes": int(pin.size)}}))

"""pp_scan_blocks over segment sizes on bench.py's corpus: kernel ms, passes, wall ms per size.
    python profiles/scan_sweep.py [reads] [sizes KiB ...]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
import parallelparsing_b200 as pp  # noqa: E402

reads = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
sizes = [int(x) for x in sys.argv[2:]] or [128, 256, 512, 1024, 2048, 4096]
bench.make_native(["tools", os.path.join("parallelparsing_b200", "csrc")])
pin, ptr = pp.pinned_copy(np.fromfile(bench.make_gz(reads, 150, 0, 10000), np.uint8))
dev = pp.Device(0)
pp.Core.ScanBlocks(pin, dev)
out = {}
for kb in sizes:
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        b, o, end, tot, ms, passes = pp.Core.ScanBlocks(pin, dev, kb << 10)
        w = (time.perf_counter() - t0) * 1e3
        if best is None or ms < best[0]:
            best = (ms, passes, w, len(b))
    out[f"{kb}KiB"] = {"kernel_ms": best[0], "passes": best[1], "wall_ms": best[2], "blocks": best[3]}
print(json.dumps({"reads": reads, "compressed_bytes": int(pin.size), "scan": out}))

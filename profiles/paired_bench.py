"""BASELINE config 3: paired-end R1/R2, N pairs x 150 bp (two Generator streams, seeds 0 and 1), chunk
10 000, both files decoded on ONE GPU — back to back on one context, and concurrently on two
contexts/host threads (what PairedFASTQ does).  Prints one JSON line.
    python profiles/paired_bench.py [pairs] [steps]"""
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
import parallelparsing_b200 as pp  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
__import__("__graft_entry__").build()
files = []
for seed in (0, 1):
    gz_path, idx_path = bench.make_corpus(pairs, 150, seed, 10000)
    gz = np.fromfile(gz_path, np.uint8)
    pin, ptr = pp.pinned_copy(gz)
    files.append((pin, pp.IndexIO.Deserialize(idx_path), ptr))
devs = [pp.Device(0), pp.Device(0)]
jobs = [pp.Job(devs[i], files[i][1], files[i][0].size, strict=True, zero_copy=True) for i in range(2)]


def run(i):
    jobs[i].upload(files[i][2])
    jobs[i].execute()
    jobs[i].download()


def both(concurrent):
    if not concurrent:
        run(0)
        run(1)
        return
    th = threading.Thread(target=run, args=(1,))
    th.start()
    run(0)
    th.join()


res = {}
for mode in (False, True):
    for _ in range(3):
        both(mode)
    t = time.perf_counter()
    for _ in range(steps):
        both(mode)
    res[mode] = (time.perf_counter() - t) / steps
infos = [j.info() for j in jobs]
assert infos[0].status == 0 and infos[1].status == 0
assert infos[0].total_records == infos[1].total_records == pairs, (infos[0].total_records, infos[1].total_records)
U = infos[0].total_bytes + infos[1].total_bytes
print(json.dumps({
    "workload": f"paired-end, {pairs} pairs x 150bp (Generator seeds 0/1), gzip -6, chunk 10000, one GPU, "
                "pull mode (kernels read the compressed bytes from pinned host memory) + results to host per file, host wall clock",
    "pairs": pairs, "uncompressed_bytes": U, "chunks": [infos[0].n_chunks, infos[1].n_chunks],
    "sequential": {"ms_per_step": res[False] * 1e3, "GB/s": U / res[False] / 1e9, "pairs_per_s": pairs / res[False]},
    "concurrent": {"ms_per_step": res[True] * 1e3, "GB/s": U / res[True] / 1e9, "pairs_per_s": pairs / res[True]},
    "steps": steps}))

"""BASELINE config 3: paired-end R1/R2, N pairs x 150 bp (two Generator streams, seeds 0 and 1), chunk
10 000, through the paired DecompressAll of the C ABI (pp_pair_decompress_all) on the GPUs given:
both files partitioned over them, R1 and R2 decoded concurrently per GPU, mates paired by ordinal,
top-up chunks so that every mate is co-resident.  Every step is a cold call with host buffers (plan,
allocations from the pool, pull from pinned host memory, kernels, results to the host) + free.
Checks: pair count, mates located on the same part for sampled ordinals, both files' digests equal
the single-file jobs'.  Prints one JSON line.
    python profiles/paired_bench.py [pairs] [steps] [n_gpus]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
import parallelparsing_b200 as pp  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ngpu = int(sys.argv[3]) if len(sys.argv) > 3 else 1
bench.make_native(["tools", os.path.join("parallelparsing_b200", "csrc")])
files = []
for seed in (0, 1):
    gz_path = bench.make_gz(pairs, 150, seed, 10000)
    idx_path = bench.corpus_paths(pairs, 150, seed, 10000)[2]
    if not os.path.exists(idx_path):
        pp.IndexIO.Serialize(pp.Core.BuildDeflateIndex(gz_path, 10000), idx_path)
    pin, ptr = pp.pinned_copy(np.fromfile(gz_path, np.uint8))
    files.append((pin, pp.IndexIO.Deserialize(idx_path), ptr))
devices = list(range(ngpu))


def one():
    return pp.PairedDecompressAll(devices, files[0][1], files[0][0], files[1][1], files[1][0], zero_copy=True)


pe = one()
info = pe.info()
assert pe.status == 0 and info.pairs == pairs == info.records_r1 == info.records_r2, (pe.status, info.pairs)
ubytes = 0
for g in range(ngpu):
    j1, base1, r2 = pe.part(g)
    ubytes += j1.info().total_bytes
    n1 = j1.info().total_records
    for r in (0, n1 // 3, n1 - 1):
        w, idx = pe.locate(g, base1 + r)      # the mate is on this part
        assert 0 <= idx < r2[w][0].info().total_records
ubytes2 = int(files[1][1].scalars()[0][-1])
topup = info.topup_chunks
pe.free()
for _ in range(2):
    one().free()
t0 = time.perf_counter()
t_call = t_free = 0.0
for _ in range(steps):
    a = time.perf_counter()
    h = one()
    b = time.perf_counter()
    h.free()
    t_call += b - a
    t_free += time.perf_counter() - b
dt = (time.perf_counter() - t0) / steps
print(json.dumps({
    "metric": "paired DecompressAll (R1 + R2) uncompressed GB/s", "value": (ubytes + ubytes2) / dt / 1e9, "unit": "GB/s",
    "pairs_per_s": pairs / dt, "ms_per_step": dt * 1e3, "n_gpus": ngpu, "steps": steps,
    "config": {"workload": f"Generator seeds 0/1, {pairs} pairs x 150bp, gzip -6, chunk 10000, paired by ordinal (PP_JOB_STRICT)",
               "topup_chunks": topup, "uncompressed_bytes": ubytes + ubytes2},
    "call_ms": t_call / steps * 1e3, "free_ms": t_free / steps * 1e3,
    "what": "cold pp_pair_decompress_all + pp_pair_free per step, pull mode, host wall clock"}))

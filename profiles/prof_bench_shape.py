"""The bench workload (10 M reads x 150 bp, chunk 10 000) for ncu captures: inputs resident, two warm
steps, then ONE profiled step (inflate kernel, scan, parse kernel).  Run as
    ncu --set full --clock-control none --import-source on -k regex:'pp_(inflate|parse)_kernel' \
        --launch-skip 4 -c 2 -o gpurun_out/r02_bench10M python profiles/prof_bench_shape.py
(two kernels match per step: skip the four of the warm steps)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
import parallelparsing_b200 as pp  # noqa: E402

reads = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
bench.make_native(["tools", os.path.join("parallelparsing_b200", "csrc")])
gz_path = bench.make_gz(reads, 150, 0, chunk)
idx_path = bench.corpus_paths(reads, 150, 0, chunk)[2]
if not os.path.exists(idx_path):
    pp.IndexIO.Serialize(pp.Core.BuildDeflateIndex(gz_path, chunk), idx_path)
ix = pp.IndexIO.Deserialize(idx_path)
gz = np.fromfile(gz_path, np.uint8)
job = pp.Job(pp.Device(0), ix, gz.size)
job.upload(gz.ctypes.data_as(__import__("ctypes").c_void_p))
for _ in range(3):
    job.execute()
job.download()
i = job.info()
print(f"chunks {i.n_chunks} records {i.total_records} bytes {i.total_bytes} compressed {i.compressed_bytes} "
      f"inflate {i.inflate_ms:.3f} ms parse {i.parse_ms:.3f} ms")

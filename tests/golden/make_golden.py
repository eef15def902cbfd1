"""Generates the golden fixtures in this directory.  Run from the repo root:

    python tests/golden/make_golden.py

The reference (C#) holds no tests, golden vectors or fixtures for this path and cannot
run in this image, so these vectors come from the oracle (oracle/pp_oracle.c: the C
restatement of the reference over the same zlib) and are cross-checked at generation
time against zlib itself (concat(chunks) == zlib stream inflate == Generator output).
They pin the oracle against drift and give the GPU tests fixed expected values that do
not depend on the oracle being rebuilt on the GPU box.

Fixtures:
  gen600.fastq.gz      Generator(seed 0, 600 reads, native U[128,512) lengths), zlib level 6 with a
                       Z_SYNC_FLUSH every 60000 bytes (more deflate blocks => more checkpoints, and
                       empty stored blocks in the stream), one gzip member
  gen600.chunk50.gzi   IndexIO file (Common/IndexIO.cs format) of CreateIndex(chunksize 50) by the oracle
  golden.json          KATs: .NET Random, generator md5s, per-chunk record counts / byte md5 / field digests
"""
import hashlib
import json
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import corpus  # noqa: E402
import oracle_lib as O  # noqa: E402


def main():
    import ctypes as C
    tools = C.CDLL(os.path.join(ROOT, "tools", "_build", "libpptools.so"))
    tools.ppgen_kat_next.restype = C.c_int32
    tools.ppgen_kat_next_range.restype = C.c_int32
    tools.ppgen_kat_next_range.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_double)]
    d = C.c_double(0)
    g = {"dotnet_random": {
        "Random(0).Next()": tools.ppgen_kat_next(0),
        "Random(42).Next()": tools.ppgen_kat_next(42),
        "Random(0).Next(128,512)": tools.ppgen_kat_next_range(0, 128, 512, C.byref(d)),
        "then NextDouble()": d.value}}
    fq = corpus.fastq(600)
    g["generator"] = {
        "first_line_seed0": fq.split(b"\n", 1)[0].decode(),
        "md5_600_native": hashlib.md5(fq).hexdigest(), "bytes_600_native": len(fq),
        "md5_20000_fixed150": hashlib.md5(corpus.fastq(20000, fixed=150)).hexdigest(),
    }
    gz = corpus.gz_member(fq, 6, flush_every=60000)
    gz.tofile(os.path.join(HERE, "gen600.fastq.gz"))
    assert zlib.decompress(gz.tobytes(), 31) == fq
    ox = O.OracleIndex.build(gz, 50)
    ox.serialize(os.path.join(HERE, "gen600.chunk50.gzi"))
    chunks, cat = [], []
    for k in range(ox.count - 1):
        n, recs, buf, dg = O.chunk(gz, ox, k, want_digest=True)
        p = ox.point(k)
        chunks.append({"output": p["output"], "input": p["input"], "bits": p["bits"], "offset_len": int(p["offset"].size),
                       "inflated": int(buf.size), "bytes_md5": hashlib.md5(buf.tobytes()).hexdigest(),
                       "records": int(n), "fields_md5": hashlib.md5(recs.astype("<i8").tobytes()).hexdigest()})
        cat.append(buf.tobytes())
    assert b"".join(cat) == fq, "concat(chunks) != generator output"
    end = ox.point(ox.count - 1)
    g["index"] = {"chunksize": 50, "points": ox.count, "chunk_max_bytes": ox.chunk_max_bytes,
                  "end_output": end["output"], "end_input": end["input"],
                  "gzi_md5": hashlib.md5(open(os.path.join(HERE, "gen600.chunk50.gzi"), "rb").read()).hexdigest(),
                  "gz_md5": hashlib.md5(gz.tobytes()).hexdigest()}
    g["chunks"] = chunks
    g["total_records"] = int(sum(c["records"] for c in chunks))
    json.dump(g, open(os.path.join(HERE, "golden.json"), "w"), indent=1)
    print(json.dumps(g["index"]), g["total_records"])


if __name__ == "__main__":
    main()

"""Synthetic corpora for the tests: Generator-exact FASTQ (tools/ppgen) compressed in
the ways the decoder has to cope with (dynamic / fixed / stored blocks, sync flushes,
gzip vs zlib-made members)."""
import functools
import os
import subprocess
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PPGEN = os.path.join(ROOT, "tools", "_build", "ppgen")
PPGZIP = os.path.join(ROOT, "tools", "_build", "ppgzip")


@functools.lru_cache(maxsize=None)
def fastq(reads, fixed=0, seed=0, lognormal=None, cap=0):
    cmd = [PPGEN, str(reads), "--seed", str(seed)]
    if fixed:
        cmd += ["--fixed", str(fixed)]
    if lognormal:
        cmd += ["--lognormal", str(lognormal[0]), str(lognormal[1])]
    if cap:
        cmd += ["--cap", str(cap)]
    return subprocess.run(cmd, check=True, stdout=subprocess.PIPE).stdout


def gz_member(data: bytes, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, flush_every=0, mem_level=9) -> np.ndarray:
    """One gzip member made with zlib (wbits 31)."""
    co = zlib.compressobj(level, zlib.DEFLATED, 31, mem_level, strategy)
    parts = []
    if flush_every:
        for i in range(0, len(data), flush_every):
            parts.append(co.compress(data[i:i + flush_every]))
            parts.append(co.flush(zlib.Z_SYNC_FLUSH))
    else:
        parts.append(co.compress(data))
    parts.append(co.flush())
    return np.frombuffer(b"".join(parts), np.uint8).copy()


def gz_system(data: bytes, level=6, tmpdir="/tmp") -> np.ndarray:
    """`gzip -<level>` as the reference's inputs are made (FNAME header included)."""
    p = os.path.join(tmpdir, f"pp_corpus_{os.getpid()}.fastq")
    with open(p, "wb") as f:
        f.write(data)
    subprocess.check_call(["gzip", f"-{level}", "-k", "-f", p])
    out = np.fromfile(p + ".gz", np.uint8)
    os.remove(p)
    os.remove(p + ".gz")
    return out


def gz_parallel(data: bytes, level=6, segment=1 << 20, threads=4, tmpdir="/tmp") -> np.ndarray:
    """tools/ppgzip: one member, segments joined by empty stored blocks."""
    p = os.path.join(tmpdir, f"pp_corpus_{os.getpid()}.fq")
    with open(p, "wb") as f:
        f.write(data)
    subprocess.check_call([PPGZIP, "-l", str(level), "-t", str(threads), "-s", str(segment), p, p + ".gz"])
    out = np.fromfile(p + ".gz", np.uint8)
    os.remove(p)
    os.remove(p + ".gz")
    return out

"""ctypes access to the CPU oracle (oracle/pp_oracle.c) — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs import this.  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(ROOT, "oracle", "_build", "libpporacle.so")


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        p, i32, i64, u32, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_size_t
        L.ora_build_index.restype = C.c_int
        L.ora_build_index.argtypes = [p, sz, u32, C.c_int, C.POINTER(p)]
        L.ora_index_free.argtypes = [p]
        L.ora_index_count.restype = i32
        L.ora_index_count.argtypes = [p]
        L.ora_index_chunk_max_bytes.restype = i32
        L.ora_index_chunk_max_bytes.argtypes = [p]
        for name, rt in (("output", i64), ("input", i64), ("bits", i32), ("offset_len", i32)):
            f = getattr(L, "ora_point_" + name)
            f.restype = rt
            f.argtypes = [p, C.c_int]
        L.ora_point_window.restype = C.POINTER(C.c_uint8)
        L.ora_point_window.argtypes = [p, C.c_int]
        L.ora_point_offset.restype = C.POINTER(C.c_uint8)
        L.ora_point_offset.argtypes = [p, C.c_int]
        L.ora_index_serialize.restype = C.c_int
        L.ora_index_serialize.argtypes = [p, C.c_char_p]
        L.ora_index_deserialize.restype = C.c_int
        L.ora_index_deserialize.argtypes = [C.c_char_p, C.POINTER(p)]
        L.ora_extract.restype = i64
        L.ora_extract.argtypes = [p, i64, p, C.c_int, C.c_int, p]
        L.ora_rent_size.restype = i64
        L.ora_rent_size.argtypes = [i64]
        L.ora_parse.restype = i64
        L.ora_parse.argtypes = [p, i64, p, i64, p, i64, C.POINTER(C.c_uint64)]
        L.ora_chunk.restype = i64
        L.ora_chunk.argtypes = [p, sz, p, C.c_int, p, p, i64, C.POINTER(C.c_uint64), C.POINTER(i64), C.c_int]
        L.ora_decompress_all_mt.restype = i64
        L.ora_decompress_all_mt.argtypes = [p, sz, p, C.c_int, C.c_int, C.c_int, C.POINTER(i64)]
        L.ora_naive_count.restype = i64
        L.ora_naive_count.argtypes = [p, sz, C.POINTER(i64)]
        L.ora_naive_records.restype = i64
        L.ora_naive_records.argtypes = [p, i64, p, i64]
        L.ora_digest_bytes.restype = C.c_uint64
        L.ora_digest_bytes.argtypes = [p, i64]
        L.ora_digest_fields.restype = C.c_uint64
        L.ora_digest_fields.argtypes = [p, i64]
        L.ora_chunk_digests_mt.restype = C.c_int
        L.ora_chunk_digests_mt.argtypes = [p, sz, p, C.c_int, C.c_int, C.c_int, p]
        L.ora_block_stops.restype = i64
        L.ora_block_stops.argtypes = [p, sz, p, p, p, i64, C.POINTER(i64), C.POINTER(i64)]
        L.ora_zcat.restype = i64
        L.ora_zcat.argtypes = [p, sz, p, i64]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class OracleIndex:
    """Common/Index.cs:5-49 as restated by the oracle."""

    def __init__(self, handle):
        self.h = handle

    def __del__(self):
        if getattr(self, "h", None):
            lib().ora_index_free(self.h)
            self.h = None

    @classmethod
    def build(cls, gz: np.ndarray, chunksize: int, lift_cap: bool = False):
        h = C.c_void_p()
        rc = lib().ora_build_index(_ptr(gz), gz.size, chunksize, int(lift_cap), C.byref(h))
        if rc != 0:
            raise RuntimeError(f"oracle build_index failed rc={rc}")
        return cls(h)

    @classmethod
    def deserialize(cls, path):
        h = C.c_void_p()
        rc = lib().ora_index_deserialize(path.encode(), C.byref(h))
        if rc != 0:
            raise RuntimeError("oracle deserialize failed")
        return cls(h)

    def serialize(self, path):
        assert lib().ora_index_serialize(self.h, path.encode()) == 0

    @property
    def count(self):
        return lib().ora_index_count(self.h)

    @property
    def chunk_max_bytes(self):
        return lib().ora_index_chunk_max_bytes(self.h)

    def point(self, i):
        L = lib()
        ol = L.ora_point_offset_len(self.h, i)
        win = np.ctypeslib.as_array(L.ora_point_window(self.h, i), shape=(32768,)).copy()
        off = np.ctypeslib.as_array(L.ora_point_offset(self.h, i), shape=(ol,)).copy() if ol else np.zeros(0, np.uint8)
        return dict(output=L.ora_point_output(self.h, i), input=L.ora_point_input(self.h, i),
                    bits=L.ora_point_bits(self.h, i), window=win, offset=off)

    def outputs(self):
        return [lib().ora_point_output(self.h, i) for i in range(self.count)]

    def inputs(self):
        return [lib().ora_point_input(self.h, i) for i in range(self.count)]


REC_FIELDS = ("start", "idnFrom", "idnLen", "seqFrom", "seqLen", "plsFrom", "plsLen", "qltFrom", "qltLen")


def parse(prepend: np.ndarray, rest: np.ndarray, cap=None, want_digest=False):
    """Parsing.Parse over CombinedMemory(prepend, rest) — rest is the whole rented array."""
    if cap is None:
        cap = (prepend.size + rest.size) // 4 + 4
    recs = np.zeros((cap, 9), np.int64)
    dg = C.c_uint64(0)
    n = lib().ora_parse(_ptr(prepend) if prepend.size else None, prepend.size, _ptr(rest), rest.size,
                        _ptr(recs), cap, C.byref(dg) if want_digest else None)
    recs = recs[: min(n, cap)]
    return (n, recs, dg.value) if want_digest else (n, recs)


def chunk(gz: np.ndarray, ix: OracleIndex, k: int, want_digest=False, do_copy=False):
    """One DecompressAll chunk in canonical order: returns (records, recs[n,9], inflated bytes, digest)."""
    L = lib()
    ln = L.ora_point_output(ix.h, k + 1) - L.ora_point_output(ix.h, k)
    rent = L.ora_rent_size(ln)
    buf = np.zeros(max(rent, 1), np.uint8)
    cap = (ln + L.ora_point_offset_len(ix.h, k)) // 4 + 4
    recs = np.zeros((cap, 9), np.int64)
    dg = C.c_uint64(0)
    produced = C.c_int64(0)
    n = L.ora_chunk(_ptr(gz), gz.size, ix.h, k, _ptr(buf), _ptr(recs), cap, C.byref(dg) if want_digest else None,
                    C.byref(produced), int(do_copy))
    if n < 0:
        raise RuntimeError(f"oracle chunk {k} failed rc={n}")
    return n, recs[:n], buf[: produced.value], dg.value


def extract(gz: np.ndarray, ix: OracleIndex, k: int):
    L = lib()
    fi, ti = L.ora_point_input(ix.h, k), L.ora_point_input(ix.h, k + 1)
    pos = max(fi - 1, 0)
    seg = gz[pos: pos + (ti - fi + 1)]
    ln = L.ora_point_output(ix.h, k + 1) - L.ora_point_output(ix.h, k)
    buf = np.zeros(max(ln, 1), np.uint8)
    seg = np.ascontiguousarray(seg)
    produced = L.ora_extract(_ptr(seg), seg.size, ix.h, k, k + 1, _ptr(buf))
    if produced < 0:
        raise RuntimeError(f"oracle extract failed rc={produced}")
    return buf[:produced]


def decompress_all_mt(gz, ix, first=0, n=None, threads=1):
    n = ix.count - 1 - first if n is None else n
    b = C.c_int64(0)
    r = lib().ora_decompress_all_mt(_ptr(gz), gz.size, ix.h, first, n, threads, C.byref(b))
    return r, b.value


def naive_count(gz):
    b = C.c_int64(0)
    r = lib().ora_naive_count(_ptr(gz), gz.size, C.byref(b))
    return r, b.value


def zcat(gz, cap):
    out = np.zeros(cap, np.uint8)
    n = lib().ora_zcat(_ptr(gz), gz.size, _ptr(out), cap)
    if n < 0:
        raise RuntimeError(f"zcat failed {n}")
    return out[:n]


def naive_records(data: np.ndarray):
    """SimpleDecompressor/Parsing.cs:9-49 over the inflated stream (+ zero tail): (count, recs[n,8]) with
    idnFrom, idnLen, seqFrom, seqLen, plsFrom, plsLen, qltFrom, qltLen as stream offsets; -1 = throws."""
    cap = data.size // 4 + 4
    recs = np.zeros((cap, 8), np.int64)
    n = lib().ora_naive_records(_ptr(data), data.size, _ptr(recs), cap)
    return n, recs[: max(n, 0)]


def digest_bytes(a: np.ndarray) -> int:
    a = np.ascontiguousarray(a, np.uint8)
    return int(lib().ora_digest_bytes(_ptr(a) if a.size else None, a.size))


def digest_fields(recs: np.ndarray) -> int:
    recs = np.ascontiguousarray(recs, np.int64)
    return int(lib().ora_digest_fields(_ptr(recs) if recs.size else None, recs.shape[0]))


def chunk_digests(gz, ix, first=0, n=None, threads=None):
    """Per chunk (inflated length, records, bytes digest, fields digest) from the oracle, `threads` workers."""
    import os
    n = ix.count - 1 - first if n is None else n
    out = np.zeros((max(n, 1), 4), np.uint64)
    rc = lib().ora_chunk_digests_mt(_ptr(gz), gz.size, ix.h, first, n, threads or (os.cpu_count() or 1), _ptr(out))
    if rc != 0:
        raise RuntimeError(f"oracle chunk digests failed rc={rc}")
    return out[:n]


def block_stops(gz: np.ndarray):
    """zlib's Z_BLOCK stops where a checkpoint may sit: (bits[], outs[], kinds[], end_bit, total_out)."""
    cap = max(1024, gz.size // 16)
    bits, outs, kinds = np.zeros(cap, np.int64), np.zeros(cap, np.int64), np.zeros(cap, np.uint8)
    end, tot = C.c_int64(-1), C.c_int64(-1)
    n = lib().ora_block_stops(_ptr(gz), gz.size, _ptr(bits), _ptr(outs), _ptr(kinds), cap, C.byref(end), C.byref(tot))
    if n < 0:
        raise RuntimeError(f"oracle block_stops failed rc={n}")
    assert n <= cap
    return bits[:n], outs[:n], kinds[:n], end.value, tot.value

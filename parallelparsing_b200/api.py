"""Host-side mirror of the reference's interface for the decode path, on top of the
C ABI (include/ppb200.h).  Same names, argument meaning and error behaviour as
the C# it stands in for, so tests read like tests of the reference:

    Index / Point             Common/Index.cs:5-82
    IndexIO.Serialize/...     Common/IndexIO.cs:7-53
    Core.BuildDeflateIndex    Decompressor/Core.cs:14    (CreateIndex, host)
    Core.ExtractDeflateIndex  Decompressor/Core.cs:133   (Decompress(checkpoint), GPU)
    Parsing.Parse             Decompressor/Parsing.cs:11 (GPU)
    BatchedFASTQ              Decompressor/BatchedFASTQ.cs:10 (DecompressAll, GPU)
    FastqRecord               Common/FastqRecord.cs:8

Python is used here only because the image has no .NET toolchain; the C# binding a
maintainer would add to Interop/ is in INTEGRATION.md.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import WINSIZE, ZException, check, lib

__all__ = ["Device", "Index", "Point", "IndexIO", "Core", "Parsing", "BatchedFASTQ", "PairedFASTQ", "FastqRecord", "Job",
           "MultiGpuDecompressAll", "PairedDecompressAll", "partition_chunks",
           "ZException", "pinned_copy"]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


class Device:
    """One GPU (pp_ctx).  Calls on one Device are serialised; use one per thread/GPU."""

    _default = {}

    def __init__(self, ordinal=0):
        h = C.c_void_p()
        check(lib().pp_open(ordinal, C.byref(h)), "pp_open")
        self.h = h
        self.ordinal = ordinal

    @classmethod
    def default(cls, ordinal=0):
        if ordinal not in cls._default:
            cls._default[ordinal] = cls(ordinal)
        return cls._default[ordinal]

    def close(self):
        if self.h:
            lib().pp_close(self.h)
            self.h = None


class Point:
    """Common/Index.cs:51-82."""

    __slots__ = ("Output", "Input", "Bits", "Window", "offset")

    def __init__(self, output, input, bits, window=None, offset=None):
        self.Output, self.Input, self.Bits = output, input, bits
        self.Window = window if window is not None else np.zeros(WINSIZE, np.uint8)
        self.offset = offset


class Index:
    """Common/Index.cs:5-49 — a list of Points with AddPoint's window un-rotation."""

    def __init__(self, points=None, _handle=None):
        if _handle is not None:
            self.h = _handle
        else:
            h = C.c_void_p()
            check(lib().pp_index_new(C.byref(h)), "pp_index_new")
            self.h = h
            for p in points or []:
                self.Add(p)

    def __del__(self):
        h = getattr(self, "h", None)
        if h:
            lib().pp_index_free(h)
            self.h = None

    @property
    def Count(self):
        return lib().pp_index_count(self.h)

    def __len__(self):
        return self.Count

    @property
    def ChunkMaxBytes(self):
        return lib().pp_index_chunk_max_bytes(self.h)

    def __getitem__(self, i):
        if i < 0:
            i += self.Count
        v = _lib.PPPoint()
        check(lib().pp_index_point(self.h, i, C.byref(v)), "pp_index_point")
        win = np.ctypeslib.as_array(v.window, shape=(WINSIZE,)).copy()
        off = np.ctypeslib.as_array(v.offset, shape=(v.offset_len,)).copy() if v.offset_len else np.zeros(0, np.uint8)
        return Point(v.output, v.input, v.bits, win, off)

    def scalars(self):
        """(Output[], Input[], Bits[], offsetLen[]) without copying windows."""
        n = self.Count
        out = np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros(n, np.int32), np.zeros(n, np.int32)
        v = _lib.PPPoint()
        for i in range(n):
            lib().pp_index_point(self.h, i, C.byref(v))
            out[0][i], out[1][i], out[2][i], out[3][i] = v.output, v.input, v.bits, v.offset_len
        return out

    def Add(self, p: Point):
        """Index.Add: append a finished Point as is (Common/Index.cs:22)."""
        win = np.ascontiguousarray(p.Window, np.uint8)
        off = np.ascontiguousarray(p.offset if p.offset is not None else np.zeros(0, np.uint8), np.uint8)
        check(lib().pp_index_add(self.h, p.Bits, p.Input, p.Output, _ptr(win), _ptr(off), off.size))

    def AddPoint(self, bits, input, output, left, window, offset):
        """Index.AddPoint (Common/Index.cs:24-48)."""
        win = np.ascontiguousarray(window, np.uint8)
        off = np.ascontiguousarray(offset if offset is not None else np.zeros(0, np.uint8), np.uint8)
        check(lib().pp_index_add_point(self.h, bits, input, output, left, _ptr(win), _ptr(off), off.size))


class IndexIO:
    """Common/IndexIO.cs:7-53 — same bytes on disk."""

    @staticmethod
    def Serialize(index: Index, path: str, compact: bool = False):
        """compact=True writes file version 1 (extension: zlib-compressed windows; this library's
        Deserialize reads both versions, the reference only version 0)."""
        f = lib().pp_index_serialize_v1 if compact else lib().pp_index_serialize
        check(f(index.h, str(path).encode()), "IndexIO.Serialize")

    @staticmethod
    def Deserialize(path: str) -> Index:
        h = C.c_void_p()
        check(lib().pp_index_deserialize(str(path).encode(), C.byref(h)), "IndexIO.Deserialize")
        return Index(_handle=h)


def _as_u8(buf):
    if isinstance(buf, np.ndarray):
        return np.ascontiguousarray(buf.reshape(-1).view(np.uint8))
    return np.frombuffer(buf, np.uint8)


class Core:
    """Decompressor/Core.cs:12-193."""

    @staticmethod
    def BuildDeflateIndex(file, chunksize: int, lift_record_cap=False) -> Index:
        """CreateIndex.  `file` is a path or the .gz bytes (the reference takes a FileStream)."""
        h = C.c_void_p()
        flags = _lib.PP_INDEX_LIFT_RECORD_CAP if lift_record_cap else 0
        if isinstance(file, (str, bytes)) and not isinstance(file, bytes):
            rc = lib().pp_index_create_file(str(file).encode(), chunksize, flags, C.byref(h))
        else:
            gz = _as_u8(file)
            rc = lib().pp_index_create(_ptr(gz), gz.size, chunksize, flags, C.byref(h))
        check(rc, "Core.BuildDeflateIndex")
        return Index(_handle=h)

    @staticmethod
    def BuildDeflateIndexGpu(file, chunksize: int, device=None, lift_record_cap=False, want_stats=False):
        """CreateIndex on the GPU (pp_index_create_gpu): the same Index as BuildDeflateIndex, without the
        serial inflate pass of Core.cs:41-127.  One gzip member; anything else raises ZException with
        PP_E_UNSUPPORTED (-106) and BuildDeflateIndex remains the way.  want_stats: (Index, dict of ms)."""
        dev = device or Device.default()
        gz = _as_u8(np.fromfile(file, np.uint8) if isinstance(file, str) else file)
        h = C.c_void_p()
        st = _lib.PPCreateStats()
        flags = _lib.PP_INDEX_LIFT_RECORD_CAP if lift_record_cap else 0
        check(lib().pp_index_create_gpu(dev.h, _ptr(gz), gz.size, chunksize, flags, C.byref(h), C.byref(st)),
              "Core.BuildDeflateIndexGpu")
        ix = Index(_handle=h)
        return (ix, {n: getattr(st, n) for n, _ in st._fields_ if n != "pad"}) if want_stats else ix

    @staticmethod
    def ScanBlocks(file, device=None, segment_bytes=0):
        """GPU-assisted CreateIndex, first slice (pp_scan_blocks): (start_bits[], out_offsets[], end_bit,
        total_out, kernel_ms, passes) — every deflate block's first bit and output offset, the stops
        BuildDeflateIndex's inflate(Z_BLOCK) pass makes (Core.cs:64,98)."""
        dev = device or Device.default()
        gz = _as_u8(np.fromfile(file, np.uint8) if isinstance(file, str) else file)
        cap = max(1024, gz.size // 64)
        bits, outs = np.zeros(cap, np.int64), np.zeros(cap, np.int64)
        n, end, tot, ms, passes = C.c_int64(), C.c_int64(), C.c_int64(), C.c_float(), C.c_int32()
        check(lib().pp_scan_blocks(dev.h, _ptr(gz), gz.size, segment_bytes, _ptr(bits), _ptr(outs), cap, C.byref(n),
                                   C.byref(end), C.byref(tot), C.byref(ms), C.byref(passes)), "pp_scan_blocks")
        return bits[: n.value], outs[: n.value], end.value, tot.value, ms.value, passes.value

    @staticmethod
    def ExtractDeflateIndex(fileBuffer, index: Index, from_point: int, buf: np.ndarray, device=None) -> int:
        """Decompress(checkpoint).  The reference passes the two Points; here the index
        and the ordinal of `from` are passed (`to` = index[from_point+1]).  `fileBuffer` is
        file[from.Input-1 : to.Input] as LazyFileReader.cs:63-69 reads it.  Returns the
        bytes produced; raises ZException where the reference throws (Core.cs:178-179)."""
        dev = device or Device.default()
        fb = _as_u8(fileBuffer)
        n = lib().pp_extract(dev.h, _ptr(fb), fb.size, index.h, from_point, _ptr(buf), buf.size)
        check(n, "Core.ExtractDeflateIndex")
        return n


class FastqRecord:
    """Common/FastqRecord.cs:8-84 — four byte slices of one record, decoded lazily."""

    __slots__ = ("_mem", "_f")

    def __init__(self, mem, idn, seq, pls, qlt):
        self._mem = mem
        self._f = (idn, seq, pls, qlt)

    def _get(self, i):
        a, n = self._f[i]
        return bytes(self._mem[a:a + n]).decode("ascii", "replace")

    Identifier = property(lambda s: s._get(0))
    Sequence = property(lambda s: s._get(1))
    Other = property(lambda s: s._get(2))
    Quality = property(lambda s: s._get(3))

    def Dispose(self):
        self._mem = None


def fields_from_line_starts(l0, l1, l2, l3, parse_end):
    """The reference's nine per-record integers (Parsing.cs:20-39) from the SoA line starts:
    start, idnFrom, idnLen, seqFrom, seqLen, plsFrom, plsLen, qltFrom, qltLen."""
    l0 = np.asarray(l0, np.int64)
    l1 = np.asarray(l1, np.int64)
    l2 = np.asarray(l2, np.int64)
    l3 = np.asarray(l3, np.int64)
    nxt = np.concatenate([l0[1:], np.array([parse_end], np.int64)]) if l0.size else l0
    return np.stack([l0 + 1, l0 + 1, l1 - l0 - 2, l1, l2 - l1 - 1, l2 + 1, l3 - l2 - 2, l3, nxt - l3 - 1], axis=1)


class Parsing:
    """Decompressor/Parsing.cs:8-70."""

    @staticmethod
    def ParseRaw(prepend, rest, device=None):
        """Returns (count, line_starts[count,4], parse_end) over CombinedMemory(prepend, rest)."""
        dev = device or Device.default()
        pre = _as_u8(prepend) if prepend is not None else np.zeros(0, np.uint8)
        rs = _as_u8(rest)
        cap = (pre.size + rs.size) // 4 + 16
        ls = np.zeros((cap, 4), np.uint32)
        pe = C.c_uint32(0)
        n = lib().pp_parse(dev.h, _ptr(pre), pre.size, _ptr(rs), rs.size, _ptr(ls), cap, C.byref(pe))
        check(n, "Parsing.Parse")
        return n, ls[:n], pe.value

    @staticmethod
    def Parse(prepend, rest, device=None):
        """Parsing.Parse(new CombinedMemory(prepend, rest)) -> list of FastqRecord."""
        pre = _as_u8(prepend) if prepend is not None else np.zeros(0, np.uint8)
        rs = _as_u8(rest)
        n, ls, pe = Parsing.ParseRaw(pre, rs, device)
        mem = np.concatenate([pre, rs])
        f = fields_from_line_starts(ls[:, 0], ls[:, 1], ls[:, 2], ls[:, 3], pe)
        return [FastqRecord(mem, (r[1], r[2]), (r[3], r[4]), (r[5], r[6]), (r[7], r[8])) for r in f]


class Job:
    """A DecompressAll plan over chunks [first, first+n) (pp_job)."""

    def __init__(self, device: Device, index: Index, gz_len: int, first=0, n=-1, strict=False, zero_copy=False,
                 pipeline=False, compact_windows=False):
        flags = ((_lib.PP_JOB_STRICT if strict else 0) | (_lib.PP_JOB_ZEROCOPY if zero_copy else 0) |
                 (_lib.PP_JOB_PIPELINE if pipeline else 0) | (_lib.PP_JOB_COMPACT_WINDOWS if compact_windows else 0))
        h = C.c_void_p()
        check(lib().pp_job_create(device.h, index.h, gz_len, first, n, flags, C.byref(h)), "pp_job_create")
        self.h = h
        self.device, self.index = device, index  # keep alive

    def free(self):
        if getattr(self, "h", None) and getattr(self, "_owned", True):
            lib().pp_job_free(self.h)
        self.h = None

    __del__ = free

    def upload(self, gz_ptr):
        check(lib().pp_job_upload(self.h, gz_ptr), "pp_job_upload")

    def file_range(self):
        """(file offset, length) of the .gz bytes this job reads."""
        lo, ln = C.c_int64(), C.c_int64()
        check(lib().pp_job_file_range(self.h, C.byref(lo), C.byref(ln)), "pp_job_file_range")
        return lo.value, ln.value

    def upload_range(self, range_ptr, range_file_offset, range_len):
        check(lib().pp_job_upload_range(self.h, range_ptr, range_file_offset, range_len), "pp_job_upload_range")

    def execute(self):
        check(lib().pp_job_execute(self.h), "pp_job_execute")

    def execute_to_host(self, dst_ptr, cap):
        """execute + every chunk's inflated bytes streamed to host memory while the decode runs."""
        check(lib().pp_job_execute_to_host(self.h, dst_ptr, cap), "pp_job_execute_to_host")

    def download(self):
        check(lib().pp_job_download(self.h), "pp_job_download")

    def run(self, gz: np.ndarray):
        self.upload(_ptr(gz))
        self.execute()
        self.download()
        return self.info()

    def info(self):
        i = _lib.PPJobInfo()
        check(lib().pp_job_info_get(self.h, C.byref(i)))
        return i

    def chunk(self, k):
        c = _lib.PPChunkInfo()
        check(lib().pp_job_chunk_info(self.h, k, C.byref(c)), "pp_job_chunk_info")
        return c

    def line_starts(self):
        n = self.info().total_records
        a = [np.zeros(n, np.uint32) for _ in range(4)]
        check(lib().pp_job_fetch_line_starts(self.h, *[_ptr(x) if n else None for x in a]), "fetch_line_starts")
        return a

    def base_histogram(self):
        """On-device consumer: byte histogram of all sequence lines (uint64[256])."""
        out = np.zeros(256, np.uint64)
        check(lib().pp_job_base_histogram(self.h, _ptr(out)), "pp_job_base_histogram")
        return out

    def count_pattern(self, pattern: bytes) -> int:
        """On-device consumer: records whose Sequence contains `pattern` (Benchmark/Naive.cs:167-180)."""
        pat = np.frombuffer(bytes(pattern), np.uint8)
        out = np.zeros(1, np.uint64)
        check(lib().pp_job_count_pattern(self.h, _ptr(pat) if pat.size else None, int(pat.size), _ptr(out)),
              "pp_job_count_pattern")
        return int(out[0])

    def digests(self):
        """Per-chunk (bytes_digest[n], fields_digest[n]) computed on the GPU (pp_job_digests)."""
        n = self.info().n_chunks
        b, f = np.zeros(max(n, 1), np.uint64), np.zeros(max(n, 1), np.uint64)
        check(lib().pp_job_digests(self.h, _ptr(b), _ptr(f)), "pp_job_digests")
        return b[:n], f[:n]

    def chunk_bytes(self, k):
        c = self.chunk(k)
        out = np.zeros(max(c.inflated, 1), np.uint8)
        check(lib().pp_job_fetch_chunk(self.h, k, _ptr(out), out.size), "fetch_chunk")
        return out[: c.inflated]

    def all_bytes(self):
        n = self.info().total_bytes
        out = np.zeros(max(n, 1), np.uint8)
        check(lib().pp_job_fetch_bytes(self.h, _ptr(out), out.size), "fetch_bytes")
        return out[:n]


def partition_chunks(index: Index, parts: int):
    """pp_partition_chunks: [(first_chunk, n_chunks)] per part — contiguous ranges of near-equal compressed size."""
    f = (C.c_int32 * parts)()
    n = (C.c_int32 * parts)()
    check(lib().pp_partition_chunks(index.h, parts, f, n), "pp_partition_chunks")
    return [(int(f[i]), int(n[i])) for i in range(parts)]


class MultiGpuDecompressAll:
    """DecompressAll over several GPUs of one box (pp_decompress_all_multi): the chunk list is cut into
    contiguous ranges of near-equal compressed size, one per GPU; every GPU reads only its byte range.
    `gz` must stay alive (and, for zero_copy, be pinned) while the object lives."""

    def __init__(self, devices, index: Index, gz, gz_len=None, strict=False, zero_copy=False, pipeline=False,
                 compact_windows=False):
        flags = ((_lib.PP_JOB_STRICT if strict else 0) | (_lib.PP_JOB_ZEROCOPY if zero_copy else 0) |
                 (_lib.PP_JOB_PIPELINE if pipeline else 0) | (_lib.PP_JOB_COMPACT_WINDOWS if compact_windows else 0))
        devs = (C.c_int32 * len(devices))(*devices)
        h = C.c_void_p()
        if isinstance(gz, np.ndarray):
            ptr, n = _ptr(gz), gz.size
        else:
            ptr, n = gz, gz_len
        rc = lib().pp_decompress_all_multi(devs, len(devices), index.h, ptr, n, flags, C.byref(h))
        if not h:
            check(rc, "pp_decompress_all_multi")
        self.h, self.status = h, rc
        self._keep = (index, gz)

    def info(self):
        i = _lib.PPMultiInfo()
        check(lib().pp_multi_info_get(self.h, C.byref(i)))
        return i

    def part(self, r):
        """(Job view, device ordinal, global ordinal of the part's first record)."""
        jh, dev, base = C.c_void_p(), C.c_int32(), C.c_int64()
        check(lib().pp_multi_part(self.h, r, C.byref(jh), C.byref(dev), C.byref(base)), "pp_multi_part")
        j = Job.__new__(Job)
        j.h = jh
        j._owned = False   # the multi handle frees it
        j._keep = self
        return j, dev.value, base.value

    def free(self):
        if getattr(self, "h", None):
            lib().pp_multi_free(self.h)
            self.h = None

    __del__ = free


class PairedDecompressAll:
    """Paired-end R1/R2 DecompressAll below the C ABI (pp_pair_decompress_all): both chunk lists are
    partitioned over `devices`; per part every R1 record's mate is resident on the same GPU."""

    def __init__(self, devices, index1: Index, gz1, index2: Index, gz2, zero_copy=False, pipeline=False,
                 compact_windows=False):
        flags = ((_lib.PP_JOB_ZEROCOPY if zero_copy else 0) | (_lib.PP_JOB_PIPELINE if pipeline else 0) |
                 (_lib.PP_JOB_COMPACT_WINDOWS if compact_windows else 0))
        devs = (C.c_int32 * len(devices))(*devices)
        h = C.c_void_p()
        rc = lib().pp_pair_decompress_all(devs, len(devices), index1.h, _ptr(gz1), gz1.size, index2.h, _ptr(gz2),
                                          gz2.size, flags, C.byref(h))
        if not h:
            check(rc, "pp_pair_decompress_all")
        self.h, self.status = h, rc
        self._keep = (index1, gz1, index2, gz2)

    def info(self):
        i = _lib.PPPairInfo()
        check(lib().pp_pair_info_get(self.h, C.byref(i)))
        return i

    def _view(self, jh):
        j = Job.__new__(Job)
        j.h, j._owned, j._keep = jh, False, self
        return j

    def part(self, g):
        """(R1 job, global ordinal of its first record, [(R2 job, global ordinal of its first record), ...])."""
        jh, base, n2 = C.c_void_p(), C.c_int64(), C.c_int32()
        check(lib().pp_pair_part(self.h, g, C.byref(jh), C.byref(base), C.byref(n2)), "pp_pair_part")
        r2 = []
        for w in range(n2.value):
            j2, b2 = C.c_void_p(), C.c_int64()
            check(lib().pp_pair_part_r2(self.h, g, w, C.byref(j2), C.byref(b2)), "pp_pair_part_r2")
            r2.append((self._view(j2), b2.value))
        return self._view(jh), base.value, r2

    def locate(self, g, ordinal):
        """(index into part g's R2 jobs, record index inside that job) of R2's record `ordinal`."""
        w, r = C.c_int32(), C.c_int64()
        check(lib().pp_pair_locate(self.h, g, ordinal, C.byref(w), C.byref(r)), "pp_pair_locate")
        return w.value, r.value

    def free(self):
        if getattr(self, "h", None):
            lib().pp_pair_free(self.h)
            self.h = None

    __del__ = free


class BatchedFASTQ:
    """DecompressAll: `for r in BatchedFASTQ(index, gzipPath, enableSsdOptimization)`
    (Decompressor/BatchedFASTQ.cs:10-27).  Records come out in canonical order — chunk
    ascending, file order inside a chunk (the reference's cross-chunk order is
    nondeterministic).  enableSsdOptimization only chose how many file
    handles the reference read with; it has no meaning here and is accepted for parity."""

    def __init__(self, index, gzipPath, enableSsdOptimization=False, device=None, strict=False):
        if isinstance(index, (str, bytes)):
            index = IndexIO.Deserialize(index)
        self._index = index
        self._gz = np.fromfile(gzipPath, np.uint8) if isinstance(gzipPath, str) else _as_u8(gzipPath)
        self._device = device or Device.default()
        self._strict = strict
        self._job = None

    def _run(self):
        if self._job is None:
            self._job = Job(self._device, self._index, self._gz.size, 0, -1, strict=self._strict)
            self._job.run(self._gz)
        return self._job

    def Count(self):
        """records.Count() — what the reference's benchmark measures (Benchmark/Naive.cs:158-162)."""
        info = self._run().info()
        if info.status < 0:
            raise ZException(info.status, "DecompressAll")
        return info.total_records

    def __iter__(self):
        job = self._run()
        info = job.info()
        if info.status < 0:
            raise ZException(info.status, "DecompressAll")
        l0, l1, l2, l3 = job.line_starts()
        for k in range(info.n_chunks):
            c = job.chunk(k)
            if c.records == 0:
                continue
            pre = self._index[info.first_chunk + k].offset
            mem = np.concatenate([pre, job.chunk_bytes(k)])
            s = slice(c.record_base, c.record_base + c.records)
            f = fields_from_line_starts(l0[s], l1[s], l2[s], l3[s], c.parse_end)
            for r in f:
                yield FastqRecord(mem, (r[1], r[2]), (r[3], r[4]), (r[5], r[6]), (r[7], r[8]))

    def Dispose(self):
        if self._job:
            self._job.free()
            self._job = None


def pinned_copy(a: np.ndarray):
    """Copy `a` into pinned host memory from pp_host_alloc; returns (ndarray view, raw pointer).
    Free with lib().pp_host_free(ptr)."""
    p = C.c_void_p()
    check(lib().pp_host_alloc(a.size, C.byref(p)), "pp_host_alloc")
    view = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(a.size,))
    view[:] = a
    return view, p


class PairedFASTQ:
    """Paired-end R1/R2 (BASELINE config 3).  The reference has no paired-end code (README.md:9
    only names it), so this is new design kept minimal: the two files are decoded as two
    DecompressAll jobs on the same GPU and records are paired by their GLOBAL ordinal — record r
    of R1 with record r of R2 — which the per-chunk `record_base` prefix sums provide without
    requiring both files to cut their chunks at the same records (checkpoints sit on deflate block
    boundaries, which differ between the files).  Quirk H1 (a duplicate record when a checkpoint
    falls on a record boundary) would shift the ordinals of one file only, so both jobs run with
    PP_JOB_STRICT.  Iteration yields (FastqRecord, FastqRecord) pairs.

    The two jobs run CONCURRENTLY: R2 gets its own context (stream) on the same GPU and its own host
    thread, so its CTAs fill the SMs the last, thin wave of R1's chunks leaves idle (DESIGN.md §4.1,
    wave quantisation) and the other way round."""

    def __init__(self, index1, gzip1, index2, gzip2, device=None, concurrent=True):
        dev = device or Device.default()
        self._dev2 = Device(dev.ordinal) if concurrent else None
        self._a = BatchedFASTQ(index1, gzip1, device=dev, strict=True)
        self._b = BatchedFASTQ(index2, gzip2, device=self._dev2 or dev, strict=True)

    def _run_both(self):
        if self._dev2 is None or (self._a._job is not None and self._b._job is not None):
            self._a._run()
            self._b._run()
            return
        import threading
        err = []

        def work(x):
            try:
                x._run()
            except BaseException as e:  # re-raised on the caller's thread
                err.append(e)
        th = threading.Thread(target=work, args=(self._b,))
        th.start()
        work(self._a)
        th.join()
        if err:
            raise err[0]

    def Count(self):
        self._run_both()
        n1, n2 = self._a.Count(), self._b.Count()
        if n1 != n2:
            raise ValueError(f"R1 holds {n1} records, R2 holds {n2}: not a paired-end pair of files")
        return n1

    def __iter__(self):
        self.Count()
        return zip(iter(self._a), iter(self._b))

    def Dispose(self):
        self._a.Dispose()
        self._b.Dispose()
        if self._dev2 is not None:
            self._dev2.close()
            self._dev2 = None

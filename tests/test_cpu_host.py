"""CPU tests (no GPU) of the product's HOST side: the C ABI exports what include/ppb200.h
declares, CreateIndex / IndexIO of libppb200.so are byte-compatible with the oracle's
restatement of the reference, the decode entry points fail loudly without a device (there is
no CPU fallback), and the multi-GPU chunk partitioning works under torch.distributed/gloo."""
import ctypes as C
import hashlib
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import corpus
import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden")
G = json.load(open(os.path.join(GOLD, "golden.json")))


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_abi_exports_every_declared_symbol():
    import parallelparsing_b200 as pp
    from parallelparsing_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "ppb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(pp_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], check=True, stdout=subprocess.PIPE, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    missing = declared - exported
    assert not missing, f"declared in ppb200.h but not exported: {sorted(missing)}"
    assert {n for n, _, _ in _lib.SYMBOLS} == declared  # the ctypes binding covers the whole header
    L = pp.lib()
    assert L.pp_abi_version() == 1
    assert b"data error" in L.pp_strerror(-3).lower() or L.pp_strerror(-3)


def _assert_same_index(ix, ox):
    assert ix.Count == ox.count
    assert ix.ChunkMaxBytes == ox.chunk_max_bytes
    for i in range(ox.count):
        a, b = ix[i], ox.point(i)
        assert (a.Output, a.Input, a.Bits) == (b["output"], b["input"], b["bits"]), i
        assert np.array_equal(a.Window, b["window"]), i
        assert np.array_equal(a.offset, b["offset"]), i


@pytest.mark.parametrize("kind", ["zlib6", "zlib1", "gzip_native", "ppgzip", "syncflush"])
def test_create_index_matches_oracle(kind):
    """Core.BuildDeflateIndex (Core.cs:14-131) of the product library vs the oracle: same points."""
    import parallelparsing_b200 as pp
    if kind == "gzip_native":
        gz, cs = corpus.gz_system(corpus.fastq(5000), 6), 1000
    elif kind == "ppgzip":
        gz, cs = corpus.gz_parallel(corpus.fastq(20000, fixed=150), 6, segment=1 << 20), 1000
    elif kind == "syncflush":
        gz, cs = corpus.gz_member(corpus.fastq(3000, fixed=150), 6, flush_every=50000), 100
    else:
        gz, cs = corpus.gz_member(corpus.fastq(8000, fixed=150), int(kind[-1])), 1000
    _assert_same_index(pp.Core.BuildDeflateIndex(gz, cs), O.OracleIndex.build(gz, cs))


def test_index_io_golden_bytes(tmp_path):
    """IndexIO.Deserialize/Serialize (Common/IndexIO.cs:7-53) on the golden file."""
    import parallelparsing_b200 as pp
    src = os.path.join(GOLD, "gen600.chunk50.gzi")
    ix = pp.IndexIO.Deserialize(src)
    assert ix.Count == G["index"]["points"]
    out = str(tmp_path / "o.gzi")
    pp.IndexIO.Serialize(ix, out)
    ref, got = open(src, "rb").read(), open(out, "rb").read()
    assert got[:4] == ref[:4] and got[8:] == ref[8:]  # ChunkMaxBytes is lost by Deserialize (quirk H7)
    # a freshly built index serialises to exactly the golden bytes
    gz = np.fromfile(os.path.join(GOLD, "gen600.fastq.gz"), np.uint8)
    pp.IndexIO.Serialize(pp.Core.BuildDeflateIndex(gz, G["index"]["chunksize"]), out)
    assert hashlib.md5(open(out, "rb").read()).hexdigest() == G["index"]["gzi_md5"]
    # and the oracle reads what the product wrote
    _assert_same_index(pp.IndexIO.Deserialize(out), O.OracleIndex.deserialize(out))


def test_index_io_version1_compact_windows(tmp_path):
    """Extension: IndexIO file version 1 (leading reserved int32 = 1, windows zlib-compressed).  Same
    points back, a much smaller file, version 0 files unchanged, damaged files rejected."""
    import parallelparsing_b200 as pp
    gz = corpus.gz_member(corpus.fastq(8000, fixed=150), 6)
    ix = pp.Core.BuildDeflateIndex(gz, 500)
    v0, v1 = str(tmp_path / "v0.gzi"), str(tmp_path / "v1.gzi")
    pp.IndexIO.Serialize(ix, v0)
    pp.IndexIO.Serialize(ix, v1, compact=True)
    b0, b1 = open(v0, "rb").read(), open(v1, "rb").read()
    assert b0[:4] == b"\0\0\0\0" and b1[:4] == b"\1\0\0\0" and b0[4:12] == b1[4:12]
    assert len(b1) * 2 < len(b0)
    a, b = pp.IndexIO.Deserialize(v0), pp.IndexIO.Deserialize(v1)
    assert a.Count == b.Count == ix.Count
    for i in range(ix.Count):
        p, q, r = ix[i], a[i], b[i]
        assert (p.Output, p.Input, p.Bits) == (q.Output, q.Input, q.Bits) == (r.Output, r.Input, r.Bits)
        assert np.array_equal(p.Window, q.Window) and np.array_equal(p.Window, r.Window)
        assert np.array_equal(p.offset, q.offset) and np.array_equal(p.offset, r.offset)
    _assert_same_index(a, O.OracleIndex.deserialize(v0))  # the oracle (version 0 only) still reads what we wrote
    # truncated / corrupted version-1 files are format errors, not crashes
    for bad in (b1[: len(b1) // 2], b1[:40] + bytes([b1[40] ^ 0xff]) + b1[41:]):
        pth = str(tmp_path / "bad.gzi")
        open(pth, "wb").write(bad)
        with pytest.raises(pp.ZException) as e:
            pp.IndexIO.Deserialize(pth)
        assert e.value.Code == -105


def test_index_errors():
    import parallelparsing_b200 as pp
    gz = corpus.gz_member(corpus.fastq(12, fixed=20000), 6)
    with pytest.raises(pp.ZException) as e:
        pp.Core.BuildDeflateIndex(gz, 2)  # Core.cs:93 IndexOutOfRangeException (quirk H2)
    assert e.value.Code == -104
    assert pp.Core.BuildDeflateIndex(gz, 2, lift_record_cap=True).Count >= 2
    with pytest.raises(pp.ZException) as e:
        pp.Core.BuildDeflateIndex(np.frombuffer(b"not a gzip file at all, definitely", np.uint8), 100)
    assert e.value.Code == -3  # Z_DATA_ERROR (Core.cs:74)
    with pytest.raises(pp.ZException):
        pp.IndexIO.Deserialize("/nonexistent/index.gzi")


def test_add_point_unrotates_window():
    """Index.AddPoint (Common/Index.cs:24-48): the circular 32 KB window is stored in stream order."""
    import parallelparsing_b200 as pp
    w = (np.arange(32768) % 251).astype(np.uint8)
    ix = pp.Index()
    left = 1000
    ix.AddPoint(3, 50, 70000, left, w, np.frombuffer(b"@abc", np.uint8))
    p = ix[0]
    assert np.array_equal(p.Window, np.concatenate([w[32768 - left:], w[:32768 - left]]))
    assert (p.Bits, p.Input, p.Output, bytes(p.offset)) == (3, 50, 70000, b"@abc")
    assert ix.ChunkMaxBytes == 70000


@pytest.mark.skipif(_have_gpu(), reason="checks the no-device behaviour")
def test_decode_fails_loudly_without_a_device():
    import parallelparsing_b200 as pp
    with pytest.raises(pp.ZException) as e:
        pp.Device(0)
    assert e.value.Code == -101  # PP_E_NO_DEVICE: no CPU fallback
    gz = np.fromfile(os.path.join(GOLD, "gen600.fastq.gz"), np.uint8)
    with pytest.raises(pp.ZException):
        pp.BatchedFASTQ(pp.IndexIO.Deserialize(os.path.join(GOLD, "gen600.chunk50.gzi")), gz).Count()


def test_product_never_references_the_oracle():
    for base, _, files in os.walk(os.path.join(ROOT, "parallelparsing_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or f == "Makefile":
                txt = open(os.path.join(base, f), errors="replace").read()
                assert "oracle_lib" not in txt and "pporacle" not in txt and "ora_" not in txt, f


def test_partition_chunks_properties():
    from parallelparsing_b200.shard import partition_chunks, record_bases
    rng = np.random.default_rng(0)
    for n_pts in (1, 2, 3, 9, 99, 1000):
        inputs = np.cumsum(rng.integers(1000, 2_000_000, n_pts)) + 10
        for world in (1, 2, 3, 4, 8):
            parts = partition_chunks(inputs, world)
            assert len(parts) == world
            nxt = 0
            for f, n in parts:
                assert f == nxt and n >= 0
                nxt = f + n
            assert nxt == max(n_pts - 1, 0)
            if n_pts - 1 >= 8 * world:
                sizes = [int(inputs[f + n] - inputs[f]) for f, n in parts]
                assert max(sizes) - min(sizes) <= 2 * int(np.diff(inputs).max())
    assert list(record_bases([5, 0, 7])) == [0, 5, 5]


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import parallelparsing_b200 as pp
    from parallelparsing_b200.shard import partition_chunks, record_bases
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gz = np.fromfile(os.path.join(GOLD, "gen600.fastq.gz"), np.uint8)
    ix = pp.IndexIO.Deserialize(os.path.join(GOLD, "gen600.chunk50.gzi"))
    first, n = partition_chunks(ix.scalars()[1], world)[rank]
    # the per-rank decode is the GPU job on a B200; here (no device) the oracle stands in as the checker
    ox = O.OracleIndex.deserialize(os.path.join(GOLD, "gen600.chunk50.gzi"))
    recs, nbytes = O.decompress_all_mt(gz, ox, first, n, threads=1) if n else (0, 0)
    t = torch.tensor([recs, nbytes, n], dtype=torch.int64)
    gathered = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    counts = [int(g[0]) for g in gathered]
    q.put((rank, first, n, counts, int(record_bases(counts)[rank]), sum(int(g[1]) for g in gathered),
           sum(int(g[2]) for g in gathered)))
    dist.destroy_process_group()


def test_two_rank_partition_over_gloo():
    """world_size 2 on CPU: contiguous chunk ranges, no data-path collective; only the per-rank
    record counts are exchanged to number the records globally."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=180) for _ in ps)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    (r0, f0, n0, c0, b0, tb0, tn0), (r1, f1, n1, c1, b1, tb1, tn1) = res
    assert f0 == 0 and f1 == n0 and n0 + n1 == len(G["chunks"]) == tn0
    assert c0 == c1 and sum(c0) == G["total_records"] and tb0 == G["index"]["end_output"]
    assert b0 == 0 and b1 == c0[0]
    assert c0[0] == sum(c["records"] for c in G["chunks"][:n0])


def test_c_abi_from_plain_c(tmp_path):
    """The boundary as a P/Invoke binding sees it: tests/c_abi/harness.c (C99, include/ppb200.h only)
    builds an index from a .gz file, writes and reads both index file versions, walks the points, and
    gets a loud PP_E_NO_DEVICE from pp_open on a machine without a GPU.  Its numbers must equal the
    Python mirror's."""
    import subprocess
    import parallelparsing_b200 as pp
    from parallelparsing_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "harness")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(root, "include"),
                           os.path.join(root, "tests", "c_abi", "harness.c"), "-o", exe,
                           "-L", libdir, "-lppb200", f"-Wl,-rpath,{libdir}"])
    gz = corpus.gz_member(corpus.fastq(6000, fixed=150), 6)
    gz_path = str(tmp_path / "r.fastq.gz")
    gz.tofile(gz_path)
    out = subprocess.run([exe, gz_path, "700", str(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    kv = dict(line.split(" ", 1) for line in out.stdout.strip().splitlines())
    ix = pp.Core.BuildDeflateIndex(gz, 700)
    assert int(kv["points"]) == ix.Count == int(kv["rebuilt_points"])
    assert int(kv["chunk_max_bytes"]) == ix.ChunkMaxBytes
    want = sum(ix[i].Output * 31 + ix[i].Input * 7 + ix[i].Bits for i in range(ix.Count)) & 0xFFFFFFFFFFFFFFFF
    assert int(kv["point_sum"]) == want
    assert [int(x) for x in kv["partition5"].split()] == [n for _, n in pp.partition_chunks(ix, 5)]
    import torch
    code = int(kv["open"].split()[0])
    assert code == (0 if torch.cuda.is_available() else -101), kv["open"]
    if not torch.cuda.is_available():   # no device: the multi-GPU and paired calls say so, they never fall back
        assert int(kv["multi_nodata"]) == -101 and int(kv["pair_nodata"]) == -101
        assert int(kv["create_gpu_noctx"]) == -102


def test_record_cap_verdict_matches_oracle_around_the_limit():
    """Core.cs:93 throws as soon as a partial record needs its 32769th byte — also when an '@' follows in
    the same inflate span.  CreateIndex and the oracle must give the same verdict for records just
    below, at and above the limit, wherever the span boundaries fall."""
    import parallelparsing_b200 as pp
    rng = np.random.default_rng(3)
    verdicts = set()
    for seq_len in (16300, 16350, 16370, 16376, 16377, 16380, 16400, 17000):
        recs = []
        for i in range(6):
            L = seq_len if i == 3 else int(rng.integers(50, 4000))
            seq = bytes(rng.choice(list(b"ACGT"), L).astype(np.uint8))
            recs.append(b"@r%d\n" % i + seq + b"\n+\n" + b"?" * L + b"\n")
        gz = corpus.gz_member(b"".join(recs), 6)
        try:
            O.OracleIndex.build(gz, 2)
            want = 0
        except RuntimeError:
            want = -104
        try:
            pp.Core.BuildDeflateIndex(gz, 2)
            got = 0
        except pp.ZException as e:
            got = e.Code
        assert got == want, (seq_len, got, want)
        verdicts.add(want)
    assert verdicts == {0, -104}


def test_cxx_partitioner_equals_python_partitioner():
    """pp_partition_chunks (what pp_decompress_all_multi and bench.py's ranks use) against shard.partition_chunks."""
    import parallelparsing_b200 as pp
    from parallelparsing_b200.shard import partition_chunks
    gz = corpus.gz_member(corpus.fastq(20000, fixed=150), 6, flush_every=50000)
    for chunk in (100, 700, 5000, 10**6):
        ix = pp.Core.BuildDeflateIndex(gz, chunk)
        inputs = ix.scalars()[1]
        for world in (1, 2, 3, 4, 7, 8, 100):
            got = pp.partition_chunks(ix, world)
            assert got == partition_chunks(inputs, world), (chunk, world)
            assert sum(n for _, n in got) == ix.Count - 1 and got[0][0] == 0
            for (f0, n0), (f1, _) in zip(got, got[1:]):
                assert f1 == f0 + n0
    empty = pp.Index()
    assert pp.partition_chunks(empty, 3) == [(0, 0)] * 3


def _block_stats(data: np.ndarray, outs, total):
    """What pp_ci_count_kernel reports per deflate block: '@' count, first, last, largest gap (numpy)."""
    st = np.zeros((len(outs), 4), np.uint32)
    ends = list(outs[1:]) + [total]
    for i, (a, b) in enumerate(zip(outs, ends)):
        at = np.flatnonzero(data[int(a): int(b)] == 64)
        if at.size:
            st[i] = (at.size, at[0], at[-1], int(np.diff(at).max()) if at.size > 1 else 0)
        else:
            st[i] = (0, 0xFFFFFFFF, 0xFFFFFFFF, 0)
    return st


def _plan_points(gz, chunksize, flags=0):
    import zlib
    from parallelparsing_b200._lib import lib
    bits, outs, kinds, end, tot = O.block_stops(gz)
    data = np.frombuffer(zlib.decompress(gz.tobytes(), 47), np.uint8)
    st = np.ascontiguousarray(_block_stats(data, outs, tot))
    b, o = np.ascontiguousarray(bits, np.uint64), np.ascontiguousarray(outs, np.uint64)
    plan = np.zeros((len(bits) + 2, 4), np.int64)
    n = C.c_int64()
    L = lib()
    L.pp_internal_plan_points.restype = C.c_int
    L.pp_internal_plan_points.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_uint64, C.c_uint64, C.c_uint32,
                                          C.c_uint32, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    rc = L.pp_internal_plan_points(b.ctypes.data, o.ctypes.data, st.ctypes.data, len(bits), tot, gz.size, chunksize, flags,
                                   plan.ctypes.data, plan.shape[0], C.byref(n))
    return rc, plan[: n.value], data


@pytest.mark.parametrize("kind", ["zlib6", "zlib1", "syncflush", "stored", "tiny"])
def test_gpu_create_index_host_half_chooses_the_oracles_points(kind):
    """index_plan_points — the host half of pp_index_create_gpu, Core.cs:98-125 over per-block '@' statistics —
    picks exactly the oracle's points (Input, Bits, Output) and offsets, given the statistics the device
    would report (computed here with numpy from zlib's output)."""
    if kind == "syncflush":
        gz, cs = corpus.gz_member(corpus.fastq(3000, fixed=150), 6, flush_every=50000), 100
    elif kind == "stored":
        gz, cs = corpus.gz_member(corpus.fastq(2000, fixed=150), 0), 50
    elif kind == "tiny":
        gz, cs = corpus.gz_member(corpus.fastq(3, fixed=150), 6), 1000
    else:
        gz, cs = corpus.gz_member(corpus.fastq(8000, fixed=150), int(kind[-1])), 1000
    ox = O.OracleIndex.build(gz, cs)
    rc, plan, data = _plan_points(gz, cs)
    assert rc == 0 and len(plan) == ox.count
    for i in range(ox.count):
        p = ox.point(i)
        assert (int(plan[i, 0]), int(plan[i, 1]), int(plan[i, 2])) == (p["bits"], p["input"], p["output"]), i
        assert np.array_equal(data[int(plan[i, 3]): int(plan[i, 2])], p["offset"]), i


def test_gpu_create_index_host_half_record_cap():
    """... and gives the oracle's verdict on records around the 32 768-byte limit (Core.cs:93)."""
    rng = np.random.default_rng(3)
    verdicts = set()
    for seq_len in (16300, 16370, 16376, 16377, 16380, 17000):
        recs = []
        for i in range(6):
            n = seq_len if i == 3 else int(rng.integers(50, 4000))
            seq = bytes(rng.choice(list(b"ACGT"), n).astype(np.uint8))
            recs.append(b"@r%d\n" % i + seq + b"\n+\n" + b"?" * n + b"\n")
        gz = corpus.gz_member(b"".join(recs), 6)
        try:
            O.OracleIndex.build(gz, 2)
            want = 0
        except RuntimeError:
            want = -104
        rc, plan, _ = _plan_points(gz, 2)
        assert rc == want, (seq_len, rc, want)
        assert _plan_points(gz, 2, 1)[0] == 0          # PP_INDEX_LIFT_RECORD_CAP
        verdicts.add(want)
    assert verdicts == {0, -104}


def test_host_walk_of_fixed_code_blocks_equals_zlib():
    """The block scan bridges the short fixed-codes blocks at sync-flush seams on the host
    (blockscan.cu host_walk_fixed): next block start and bytes produced equal zlib's Z_BLOCK stops."""
    import zlib
    from parallelparsing_b200._lib import lib
    L = lib()
    L.pp_internal_walk_fixed.restype = C.c_int
    L.pp_internal_walk_fixed.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    walked = 0
    for data, kw in ((corpus.fastq(300, fixed=150), dict(strategy=zlib.Z_FIXED, flush_every=7000)),
                     (corpus.fastq(50), dict(strategy=zlib.Z_FIXED)),
                     (bytes(np.random.default_rng(5).integers(0, 256, 5000, dtype=np.uint8)), dict(strategy=zlib.Z_FIXED, flush_every=999))):
        gz = corpus.gz_member(data, 6, **kw)
        bits, outs, kinds, end, tot = O.block_stops(gz)
        for i in range(len(bits)):
            hdr3 = (int.from_bytes(gz[int(bits[i]) // 8: int(bits[i]) // 8 + 2].tobytes(), "little") >> (int(bits[i]) % 8)) & 7
            if hdr3 >> 1 != 1:
                continue
            nxt, ob = C.c_uint64(), C.c_uint64()
            assert L.pp_internal_walk_fixed(gz.ctypes.data, gz.size, int(bits[i]), C.byref(nxt), C.byref(ob)) == 0
            want_next = int(bits[i + 1]) if i + 1 < len(bits) else None
            want_out = (int(outs[i + 1]) if i + 1 < len(bits) else int(tot)) - int(outs[i])
            assert ob.value == want_out, i
            if want_next is not None:
                assert nxt.value == want_next, i
            walked += 1
    assert walked > 10

"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs — bit-exact bytes per checkpoint, identical record counts, identical
per-record field offsets, identical chunk boundaries.  All integer/byte work: the bar
is equality, no tolerance."""
import zlib

import numpy as np
import pytest

import corpus
import oracle_lib as O

pytestmark = pytest.mark.gpu


def _check_job_against_oracle(pp, device, gz, chunksize, lift=False, strict=False):
    ox = O.OracleIndex.build(gz, chunksize, lift)
    ix = pp.Core.BuildDeflateIndex(gz, chunksize, lift_record_cap=lift)
    assert ix.Count == ox.count
    job = pp.Job(device, ix, gz.size)
    info = job.run(gz)
    assert info.status == 0
    assert info.n_chunks == ox.count - 1
    l0, l1, l2, l3 = job.line_starts()
    total = 0
    for k in range(info.n_chunks):
        n, recs, buf, _ = O.chunk(gz, ox, k)
        c = job.chunk(k)
        assert c.status == 0
        assert c.inflated == buf.size, f"chunk {k}: inflated {c.inflated} != {buf.size}"
        got = job.chunk_bytes(k)
        assert np.array_equal(got, buf), f"chunk {k}: inflated bytes differ"
        assert c.records == n, f"chunk {k}: records {c.records} != {n}"
        assert c.record_base == total
        s = slice(c.record_base, c.record_base + c.records)
        f = pp.fields_from_line_starts(l0[s], l1[s], l2[s], l3[s], c.parse_end)
        assert np.array_equal(f, recs), f"chunk {k}: record fields differ"
        total += n
    assert info.total_records == total
    job.free()
    return total


@pytest.mark.parametrize("mode", ["dynamic6", "dynamic1", "dynamic9", "fixed", "stored", "huffman", "rle", "syncflush"])
def test_decompress_all_block_types(device, mode):
    import parallelparsing_b200 as pp
    fq = corpus.fastq(12000, fixed=150)
    kw = dict(dynamic6=dict(level=6), dynamic1=dict(level=1), dynamic9=dict(level=9),
              fixed=dict(level=6, strategy=zlib.Z_FIXED), stored=dict(level=0),
              huffman=dict(level=6, strategy=zlib.Z_HUFFMAN_ONLY), rle=dict(level=6, strategy=zlib.Z_RLE),
              syncflush=dict(level=6, flush_every=70000))[mode]
    gz = corpus.gz_member(fq, **kw)
    total = _check_job_against_oracle(pp, device, gz, 1000)
    assert total >= 12000  # quirk H1 may add duplicates, never drops


def test_decompress_all_system_gzip_native_lengths(device):
    import parallelparsing_b200 as pp
    fq = corpus.fastq(20000)  # Generator's own U[128,512) lengths
    gz = corpus.gz_system(fq, 6)
    _check_job_against_oracle(pp, device, gz, 2000)


def test_decompress_all_ppgzip_segments(device):
    import parallelparsing_b200 as pp
    fq = corpus.fastq(30000, fixed=150)
    gz = corpus.gz_parallel(fq, 6, segment=1 << 20)
    _check_job_against_oracle(pp, device, gz, 1000)


def test_long_reads_lifted_cap(device):
    import parallelparsing_b200 as pp
    fq = corpus.fastq(300, lognormal=(10000, 0.5), seed=3)
    gz = corpus.gz_member(fq, 6)
    _check_job_against_oracle(pp, device, gz, 20, lift=True)


def test_extract_single_checkpoint(device):
    """Decompress(checkpoint): Core.ExtractDeflateIndex on the exact fileBuffer LazyFileReader reads."""
    import parallelparsing_b200 as pp
    fq = corpus.fastq(8000, fixed=150)
    gz = corpus.gz_member(fq, 6)
    ox = O.OracleIndex.build(gz, 1000)
    ix = pp.Core.BuildDeflateIndex(gz, 1000)
    outs, ins = ox.outputs(), ox.inputs()
    for k in range(ox.count - 1):
        fb = gz[ins[k] - 1: ins[k + 1]]
        buf = np.zeros(outs[k + 1] - outs[k], np.uint8)
        n = pp.Core.ExtractDeflateIndex(fb, ix, k, buf, device)
        ref = O.extract(gz, ox, k)
        assert n == ref.size and np.array_equal(buf[:n], ref)


def test_concat_chunks_equals_stream(device):
    import parallelparsing_b200 as pp
    fq = corpus.fastq(9000, fixed=150)
    gz = corpus.gz_member(fq, 6)
    ix = pp.Core.BuildDeflateIndex(gz, 1000)
    job = pp.Job(device, ix, gz.size)
    job.run(gz)
    assert job.all_bytes().tobytes() == fq
    job.free()


def _parse_both(pp, device, prepend: bytes, rest: bytes):
    pre = np.frombuffer(prepend, np.uint8)
    rs = np.frombuffer(rest, np.uint8)
    rent = np.zeros(max(O.lib().ora_rent_size(rs.size), rs.size + 1), np.uint8)  # zero tail as the pool gives
    rent[: rs.size] = rs
    n_ref, recs = O.parse(pre, rent)
    n, ls, pe = pp.Parsing.ParseRaw(pre, rs, device)
    assert n == n_ref, (n, n_ref)
    f = pp.fields_from_line_starts(ls[:, 0], ls[:, 1], ls[:, 2], ls[:, 3], pe)
    assert np.array_equal(f, recs)
    return n


@pytest.mark.parametrize("case", [
    (b"", b""),
    (b"", b"@a\nAC\n+\n??\n"),
    (b"@a\nAC", b"GT\n+a\n????\n@b\nA\n+\n?\n"),
    (b"", b"@a\nAC\n+\n??\n@b\nAC\n+\n?"),            # trailing partial record dropped
    (b"", b"@a\nAC\n+\n??\n\x00@b\nAC\n+\n??\n"),     # stops at the first NUL
    (b"", b"@a\n\n+\n\n@b\nA\n+\n?\n"),               # empty sequence/quality lines are ordinary
    (b"", b"\n\nAC\n+\n??\n@b\nA\n+\n?\n"),           # empty id line: first byte skipped unchecked (Parsing.cs:19)
    (b"", b"@a\nAC\n\n\n??\n@b\nA\n+\n?\n"),          # empty '+' line: first byte skipped unchecked (:30)
    (b"", b"@a\r\nAC\r\n+\r\n??\r\n"),                # \r is not stripped
    (b"@a\nAC\n+\n??\n", b"@b\nA\n+\n?\n"),           # quirk H1: complete record in the prefix
    (b"", b"xa\nAC\nya\n??\n"),                       # '@' / '+' are not checked
])
def test_parse_edge_cases(device, case):
    import parallelparsing_b200 as pp
    _parse_both(pp, device, *case)


def test_parse_large_random_lines(device):
    import parallelparsing_b200 as pp
    rng = np.random.default_rng(5)
    body = bytearray()
    for i in range(20000):
        L = int(rng.integers(0, 300))
        body += b"@r%d\n" % i + bytes(rng.choice(list(b"ACGT"), L).astype(np.uint8)) + b"\n+\n" + b"?" * L + b"\n"
    data = bytes(body)
    for cut in (0, 1, 17, 4095, 4096, 16383):
        _parse_both(pp, device, data[:cut], data[cut:])


def test_parse_short_records_many_newlines_per_tile(device):
    """Records of ~10 bytes: a 64 KB tile then holds >20 000 newlines, far beyond the 2048 positions the
    parse kernel keeps in shared memory per emission round (the multi-round path)."""
    import parallelparsing_b200 as pp
    rng = np.random.default_rng(9)
    body = bytearray()
    for i in range(40000):
        L = int(rng.integers(1, 4))
        body += b"@%d\n" % (i % 10) + bytes(rng.choice(list(b"ACGT"), L).astype(np.uint8)) + b"\n+\n" + b"?" * L + b"\n"
    data = bytes(body)
    for cut in (0, 7, 65536, 70001):
        n = _parse_both(pp, device, data[:cut], data[cut:])
        assert n == 40000


def test_strict_drops_h1_duplicate(device):
    """Extension: PP_JOB_STRICT removes the duplicate record quirk H1 creates."""
    import parallelparsing_b200 as pp
    fq = corpus.fastq(400, fixed=150)
    recs = fq.split(b"\n")
    co = zlib.compressobj(6, zlib.DEFLATED, 31)
    parts = []
    for i in range(0, len(recs) - 1, 4):  # block end forced onto every record end
        parts.append(co.compress(b"\n".join(recs[i:i + 4]) + b"\n"))
        parts.append(co.flush(zlib.Z_FULL_FLUSH))
    parts.append(co.flush())
    gz = np.frombuffer(b"".join(parts), np.uint8).copy()
    total = _check_job_against_oracle(pp, device, gz, 20)
    assert total > 400  # the reference emits duplicates here
    ix = pp.Core.BuildDeflateIndex(gz, 20)
    job = pp.Job(device, ix, gz.size, strict=True)
    info = job.run(gz)
    assert info.total_records == 400
    job.free()


def test_corrupt_stream_reports_data_error(device):
    import parallelparsing_b200 as pp
    fq = corpus.fastq(6000, fixed=150)
    gz = corpus.gz_member(fq, 6)
    ix = pp.Core.BuildDeflateIndex(gz, 1000)
    bad = gz.copy()
    ins = ix.scalars()[1]
    mid = int((ins[1] + ins[2]) // 2)
    bad[mid: mid + 64] = 0xFF
    job = pp.Job(device, ix, bad.size)
    info = job.run(bad)
    ox = O.OracleIndex.build(gz, 1000)
    try:
        _, _, ref, _ = O.chunk(bad, ox, 1)
    except RuntimeError:
        ref = None
    if ref is None:  # zlib reports Z_DATA_ERROR: the reference throws ZException(DATA_ERROR)
        assert job.chunk(1).status == -3 and info.status == -3
    else:            # zlib decodes the damaged bits to (garbage) bytes: same garbage expected
        assert job.chunk(1).status == 0 and np.array_equal(job.chunk_bytes(1), ref)
    assert job.chunk(0).status == 0
    job.free()


def test_golden_fixture(device):
    """Committed fixture (tests/golden): expected values do not depend on the oracle being rebuilt."""
    import hashlib
    import json
    import os
    import parallelparsing_b200 as pp
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    G = json.load(open(os.path.join(gold, "golden.json")))
    gz = np.fromfile(os.path.join(gold, "gen600.fastq.gz"), np.uint8)
    ix = pp.IndexIO.Deserialize(os.path.join(gold, "gen600.chunk50.gzi"))
    job = pp.Job(device, ix, gz.size)
    info = job.run(gz)
    assert info.status == 0 and info.total_records == G["total_records"] and info.n_chunks == len(G["chunks"])
    l0, l1, l2, l3 = job.line_starts()
    for k, c in enumerate(G["chunks"]):
        ci = job.chunk(k)
        assert (ci.inflated, ci.records) == (c["inflated"], c["records"])
        assert hashlib.md5(job.chunk_bytes(k).tobytes()).hexdigest() == c["bytes_md5"]
        s = slice(ci.record_base, ci.record_base + ci.records)
        f = pp.fields_from_line_starts(l0[s], l1[s], l2[s], l3[s], ci.parse_end)
        assert hashlib.md5(f.astype("<i8").tobytes()).hexdigest() == c["fields_md5"]
    job.free()


@pytest.mark.parametrize("world", [2, 3, 8])
def test_partitioned_ranges_equal_whole(device, world):
    """Multi-GPU sharding: contiguous chunk ranges decoded as separate jobs (here on one
    GPU) give exactly the whole-file result; global record ordinals come from the per-range counts."""
    import parallelparsing_b200 as pp
    from parallelparsing_b200.shard import partition_chunks, record_bases
    fq = corpus.fastq(30000, fixed=150)
    gz = corpus.gz_member(fq, 6)
    ix = pp.Core.BuildDeflateIndex(gz, 2000)
    whole = pp.Job(device, ix, gz.size)
    wi = whole.run(gz)
    parts = partition_chunks(ix.scalars()[1], world)
    counts, cat = [], []
    for first, n in parts:
        j = pp.Job(device, ix, gz.size, first, n)
        info = j.run(gz)
        assert info.status == 0 and info.n_chunks == n
        for k in range(n):
            a, b = j.chunk(k), whole.chunk(first + k)
            assert (a.inflated, a.records, a.parse_end) == (b.inflated, b.records, b.parse_end)
        counts.append(info.total_records)
        cat.append(j.all_bytes().tobytes())
        j.free()
    assert sum(counts) == wi.total_records and b"".join(cat) == fq
    bases = record_bases(counts)
    assert [whole.chunk(f).record_base if n else None for f, n in parts] == \
        [int(b) if n else None for b, (f, n) in zip(bases, parts)]
    whole.free()


def test_zero_copy_matches_staged(device):
    """PP_JOB_ZEROCOPY: kernels pull the compressed bytes and windows from pinned host memory."""
    import parallelparsing_b200 as pp
    fq = corpus.fastq(20000, fixed=150)
    gz_np = corpus.gz_parallel(fq, 6, segment=1 << 20)
    gz, ptr = pp.pinned_copy(gz_np)
    ix = pp.Core.BuildDeflateIndex(gz_np, 1000)
    a = pp.Job(device, ix, gz.size)
    b = pp.Job(device, ix, gz.size, zero_copy=True)
    ia = a.run(gz)
    b.upload(ptr); b.execute(); b.download()
    ib = b.info()
    assert (ia.status, ia.total_records, ia.total_bytes) == (ib.status, ib.total_records, ib.total_bytes) == (0, ia.total_records, len(fq))
    assert b.all_bytes().tobytes() == fq
    for x, y in zip(a.line_starts(), b.line_starts()):
        assert np.array_equal(x, y)
    a.free(); b.free()
    pp.lib().pp_host_free(ptr)


def test_paired_end_pairs_by_global_ordinal(device):
    """BASELINE config 3 (small): R1/R2 from two Generator streams with the same read count."""
    import parallelparsing_b200 as pp
    r1, r2 = corpus.fastq(6000, fixed=150, seed=0), corpus.fastq(6000, fixed=150, seed=1)
    g1, g2 = corpus.gz_member(r1, 6), corpus.gz_member(r2, 1)   # different block boundaries on purpose
    i1, i2 = pp.Core.BuildDeflateIndex(g1, 1000), pp.Core.BuildDeflateIndex(g2, 700)
    pe = pp.PairedFASTQ(i1, g1, i2, g2, device=device)
    assert pe.Count() == 6000
    l1, l2 = r1.split(b"\n"), r2.split(b"\n")
    n = 0
    for a, b in pe:
        assert a.Identifier.encode() == l1[4 * n][1:] and b.Identifier.encode() == l2[4 * n][1:]
        assert a.Sequence.encode() == l1[4 * n + 1] and b.Quality.encode() == l2[4 * n + 3]
        n += 1
    assert n == 6000
    pe.Dispose()


@pytest.mark.parametrize("chunk", [1000, 5000, 20000, 100000])
def test_chunk_size_sweep_properties(device, chunk):
    """BASELINE config 5 in small: the same file under a sweep of chunk sizes.  Size-independent
    properties only (no per-chunk oracle): every chunk inflates to exactly to.Output-from.Output
    bytes, the concatenation is the generator's output, every read is found once (plus the H1
    duplicates the oracle's index-only count predicts), record bases are the running sum."""
    import hashlib
    import parallelparsing_b200 as pp
    fq = corpus.fastq(150000, fixed=150)
    gz = corpus.gz_parallel(fq, 6, segment=4 << 20, threads=4)
    ix = pp.Core.BuildDeflateIndex(gz, chunk)
    outs = ix.scalars()[0]
    job = pp.Job(device, ix, gz.size)
    info = job.run(gz)
    assert info.status == 0 and info.total_bytes == len(fq) == int(outs[-1])
    base = 0
    dup = 0
    for k in range(info.n_chunks):
        c = job.chunk(k)
        assert c.status == 0 and c.inflated == int(outs[k + 1] - outs[k]) and c.record_base == base
        base += c.records
        p = ix[k]
        # a checkpoint exactly on a record boundary: Point.offset holds one complete record (quirk H1)
        dup += int(k > 0 and p.offset.size > 0 and p.offset[-1] == 10 and int((p.offset == 10).sum()) == 4)
    assert info.total_records == base == 150000 + dup
    assert hashlib.md5(job.all_bytes().tobytes()).hexdigest() == hashlib.md5(fq).hexdigest()
    job.free()


def test_on_device_base_histogram(device):
    """On-device consumer (Decompressor/Program.cs:51-52 counts records and 'A's): the histogram of
    all sequence lines computed on the GPU equals the host count over the generator's output."""
    import parallelparsing_b200 as pp
    fq = corpus.fastq(20000)  # native U[128,512) lengths
    gz = corpus.gz_member(fq, 6)
    ix = pp.Core.BuildDeflateIndex(gz, 3000)
    job = pp.Job(device, ix, gz.size, strict=True)  # strict: a duplicated record would be counted twice
    info = job.run(gz)
    assert info.status == 0 and info.total_records == 20000
    h = job.base_histogram()
    seq = b"".join(fq.split(b"\n")[1::4])
    want = np.bincount(np.frombuffer(seq, np.uint8), minlength=256).astype(np.uint64)
    assert np.array_equal(h, want)
    assert int(h.sum()) == len(seq) and int(h[ord("A")]) > 0
    job.free()


def test_on_device_pattern_count(device):
    """On-device consumer for the reference benchmark's pattern search (Benchmark/Naive.cs:167-180):
    records whose Sequence contains the pattern, counted on the GPU, equal Python's `in` over the
    generator's sequences — for short, long, absent, full-line and empty patterns."""
    import parallelparsing_b200 as pp
    fq = corpus.fastq(20000)  # native U[128,512) lengths
    gz = corpus.gz_member(fq, 6)
    ix = pp.Core.BuildDeflateIndex(gz, 3000)
    job = pp.Job(device, ix, gz.size, strict=True)
    info = job.run(gz)
    assert info.status == 0 and info.total_records == 20000
    seqs = fq.split(b"\n")[1::4]
    pats = [b"GTTATACACTGC", b"ACGT", b"A", b"GATTACA", seqs[17][:40], seqs[123][-25:], seqs[5], seqs[9] + b"A",
            b"N", b"", b"ACGTACGTACGTACGTACGTACGTACGT"]
    for pat in pats:
        want = sum(1 for q in seqs if pat in q)
        assert job.count_pattern(pat) == want, pat
    assert job.count_pattern(b"ACGT") > 0 and job.count_pattern(seqs[5]) >= 1
    job.free()


def test_empty_chunk_range_and_threaded_callers(device):
    """A rank that gets no chunks (more GPUs than chunks) and callers on thread-pool threads
    (BatchedFASTQ.cs:62 spawns tasks; README.md:50 asks for thread safety)."""
    import threading
    import parallelparsing_b200 as pp
    fq = corpus.fastq(9000, fixed=150)
    gz = corpus.gz_member(fq, 6)
    ix = pp.Core.BuildDeflateIndex(gz, 1000)
    nchunks = ix.Count - 1
    j = pp.Job(device, ix, gz.size, nchunks, 0)
    info = j.run(gz)
    assert (info.status, info.n_chunks, info.total_records, info.total_bytes) == (0, 0, 0, 0)
    j.free()
    results, errors = {}, []

    def work(first, n):
        try:
            jb = pp.Job(device, ix, gz.size, first, n)
            i = jb.run(gz)
            results[first] = (i.status, i.total_records, jb.all_bytes().tobytes())
            jb.free()
        except Exception as e:  # noqa: BLE001
            errors.append(e)
    half = nchunks // 2
    ts = [threading.Thread(target=work, args=a) for a in ((0, half), (half, nchunks - half))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors
    assert results[0][0] == 0 and results[half][0] == 0
    assert results[0][2] + results[half][2] == fq
    assert results[0][1] + results[half][1] >= 9000


def test_bit_flips_agree_with_zlib(device):
    """Random bit flips in a chunk's compressed bytes: wherever zlib (the oracle's restatement of
    Core.ExtractDeflateIndex) produces bytes the GPU produces the same bytes, wherever the reference
    would throw ZException(DATA_ERROR) — invalid data, or input exhausted before `len` bytes are out
    (Core.cs:174) — the GPU path raises it too."""
    import parallelparsing_b200 as pp
    rng = np.random.default_rng(7)
    fq = corpus.fastq(4000, fixed=150, seed=2)
    gz = corpus.gz_member(fq, 6)
    ox = O.OracleIndex.build(gz, 700)
    ix = pp.Core.BuildDeflateIndex(gz, 700)
    outs, ins = ox.outputs(), ox.inputs()
    n_err = n_ok = 0
    for _ in range(30):
        k = int(rng.integers(0, ox.count - 1))
        bad = gz.copy()
        for _ in range(int(rng.integers(1, 4))):
            pos = int(rng.integers(ins[k] + 1, ins[k + 1] - 1))
            bad[pos] ^= np.uint8(1 << int(rng.integers(0, 8)))
        try:
            ref = O.extract(bad, ox, k)
        except RuntimeError:
            ref = None
        fb = bad[ins[k] - 1: ins[k + 1]]
        buf = np.zeros(outs[k + 1] - outs[k], np.uint8)
        try:
            n = pp.Core.ExtractDeflateIndex(fb, ix, k, buf, device)
            got = buf[:n]
        except pp.ZException as e:
            assert e.Code == -3
            got = None
        if ref is None:
            assert got is None, f"chunk {k}: zlib reports a data error, the GPU path does not"
            n_err += 1
        else:
            assert got is not None and got.size == ref.size and np.array_equal(got, ref), f"chunk {k}"
            n_ok += 1
    assert n_err > 0 and n_err + n_ok == 30


def test_bit_flips_in_block_headers_agree_with_zlib(device):
    """Flips aimed at the dynamic block header every chunk starts with (code-length code, run-length
    coded lengths): the kernel decodes those speculatively in parallel, zlib serially; the verdicts
    (bytes, or DATA_ERROR for an over-subscribed / incomplete code, a bad repeat, a missing
    end-of-block) must agree."""
    import parallelparsing_b200 as pp
    rng = np.random.default_rng(11)
    fq = corpus.fastq(3000, fixed=150, seed=5)
    gz = corpus.gz_member(fq, 6)
    ox = O.OracleIndex.build(gz, 300)
    ix = pp.Core.BuildDeflateIndex(gz, 300)
    outs, ins = ox.outputs(), ox.inputs()
    n_err = n_ok = 0
    for _ in range(60):
        k = int(rng.integers(0, ox.count - 1))
        span = min(90, ins[k + 1] - ins[k] - 2)
        if span < 8:
            continue
        bad = gz.copy()
        for _ in range(int(rng.integers(1, 3))):
            bad[int(ins[k] + rng.integers(0, span))] ^= np.uint8(1 << int(rng.integers(0, 8)))
        try:
            ref = O.extract(bad, ox, k)
        except RuntimeError:
            ref = None
        buf = np.zeros(outs[k + 1] - outs[k], np.uint8)
        try:
            n = pp.Core.ExtractDeflateIndex(bad[ins[k] - 1: ins[k + 1]], ix, k, buf, device)
            got = buf[:n]
        except pp.ZException as e:
            assert e.Code == -3
            got = None
        if ref is None:
            assert got is None, f"chunk {k}: zlib reports a data error, the GPU path does not"
            n_err += 1
        else:
            assert got is not None and got.size == ref.size and np.array_equal(got, ref), f"chunk {k}"
            n_ok += 1
    assert n_err > 10


@pytest.mark.parametrize("seed", range(4))
def test_parse_fuzz(device, seed):
    """Random line structures (empty lines, NULs, CRs, missing final newline, arbitrary prepend split)
    through Parsing.Parse on the GPU against the oracle's literal restatement of Parsing.cs."""
    import parallelparsing_b200 as pp
    rng = np.random.default_rng(1000 + seed)
    alphabet = np.frombuffer(b"\n\n\n@+ACGT?\r", np.uint8)
    for trial in range(60):
        n = int(rng.integers(0, 4000))
        body = alphabet[rng.integers(0, alphabet.size, n)].copy()
        if n and rng.random() < 0.3:  # a NUL somewhere: the reference stops there
            body[int(rng.integers(0, n))] = 0
        if rng.random() < 0.5:        # mostly well-formed records with noise lines in between
            recs = []
            for i in range(int(rng.integers(1, 40))):
                L = int(rng.integers(0, 60))
                recs.append(b"@r%d\n" % i + bytes(rng.choice(list(b"ACGT"), L).astype(np.uint8)) + b"\n+\n" + b"?" * L + b"\n")
                if rng.random() < 0.1:
                    recs.append(bytes(body[: int(rng.integers(0, 8))]))
            body = np.frombuffer(b"".join(recs), np.uint8)
        data = body.tobytes()
        cut = int(rng.integers(0, len(data) + 1))
        _parse_both(pp, device, data[:cut], data[cut:])


def test_decompress_all_with_irregular_records(device):
    """DecompressAll over a stream the fast parser must NOT trust: NUL bytes inside chunks (the
    reference stops there, Parsing.cs:16,60), empty id / '+' lines (first byte skipped unchecked,
    :19,:30), CRLF line ends.  Such chunks are routed through the exact parser; everything is
    compared with the oracle chunk by chunk."""
    import parallelparsing_b200 as pp
    rng = np.random.default_rng(11)
    recs = []
    for i in range(6000):
        L = int(rng.integers(20, 200))
        seq = bytes(rng.choice(list(b"ACGT"), L).astype(np.uint8))
        r = b"@read%d\n" % i + seq + b"\n+\n" + b"?" * L + b"\n"
        x = rng.random()
        if x < 0.002:
            r = r.replace(b"?", b"\x00", 1)           # NUL inside a quality line
        elif x < 0.004:
            r = b"\n" + seq + b"\n+\n" + b"?" * L + b"\n"  # empty id line
        elif x < 0.006:
            r = b"@read%d\n" % i + seq + b"\n\n" + b"?" * L + b"\n"  # empty '+' line
        elif x < 0.02:
            r = r.replace(b"\n", b"\r\n")
        recs.append(r)
    data = b"".join(recs)
    gz = corpus.gz_member(data, 6, flush_every=90000)
    ox = O.OracleIndex.build(gz, 300)
    ix = pp.Core.BuildDeflateIndex(gz, 300)
    job = pp.Job(device, ix, gz.size)
    info = job.run(gz)
    assert info.status == 0 and info.n_chunks == ox.count - 1 and info.exact_chunks > 0
    l0, l1, l2, l3 = job.line_starts()
    total = 0
    for k in range(info.n_chunks):
        n, recs_o, buf, _ = O.chunk(gz, ox, k)
        c = job.chunk(k)
        assert np.array_equal(job.chunk_bytes(k), buf)
        assert c.records == n, f"chunk {k}: {c.records} != {n}"
        s = slice(c.record_base, c.record_base + c.records)
        assert np.array_equal(pp.fields_from_line_starts(l0[s], l1[s], l2[s], l3[s], c.parse_end), recs_o), f"chunk {k}"
        total += n
    assert total == info.total_records
    job.free()


# ----------------------------------------------------------------------------------------------
# Parity at the BASELINE.json configurations, at their stated sizes.  Every chunk is compared with
# the oracle through four integers (inflated length, record count, digest of the inflated bytes,
# digest of the nine per-record integers); the digests are computed on the GPU (pp_job_digests) and,
# independently, by the oracle over its own output — only K x 4 integers cross PCIe.
# ----------------------------------------------------------------------------------------------
import functools
import hashlib
import os
import subprocess
import tempfile


def _digest_parity(pp, device, gz, ix, ox, zero_copy=False, gz_ptr=None):
    """All chunks of (gz, ix) on the GPU against the oracle's (length, records, bytes digest, fields digest)."""
    assert ix.Count == ox.count
    job = pp.Job(device, ix, gz.size, zero_copy=zero_copy)
    if zero_copy:
        job.upload(gz_ptr); job.execute(); job.download()
        info = job.info()
    else:
        info = job.run(gz)
    assert info.status == 0 and info.n_chunks == ox.count - 1
    want = O.chunk_digests(gz, ox)
    bd, fd = job.digests()
    base = 0
    for k in range(info.n_chunks):
        c = job.chunk(k)
        assert (c.status, c.inflated, c.records, c.record_base) == (0, int(want[k, 0]), int(want[k, 1]), base), f"chunk {k}"
        assert int(bd[k]) == int(want[k, 2]), f"chunk {k}: inflated bytes differ (digest)"
        assert int(fd[k]) == int(want[k, 3]), f"chunk {k}: record fields differ (digest)"
        base += c.records
    assert info.total_records == base
    job.free()
    return info


@functools.lru_cache(maxsize=None)
def _generator_file(reads, fixed=150, seed=0, system_gzip=False, lognormal=None, cap=0):
    """Generator-exact FASTQ -> one gzip member on disk (cached for the session); returns the .gz path."""
    d = tempfile.mkdtemp(prefix="pp_cfg_")
    fq_path, gz_path = os.path.join(d, "reads.fastq"), os.path.join(d, "reads.fastq.gz")
    cmd = [corpus.PPGEN, str(reads), "--seed", str(seed)]
    if lognormal:
        cmd += ["--lognormal", str(lognormal[0]), str(lognormal[1])]
        if cap:
            cmd += ["--cap", str(cap)]
    elif fixed:
        cmd += ["--fixed", str(fixed)]
    if system_gzip:
        with open(fq_path, "wb") as f:
            subprocess.check_call(cmd, stdout=f)
        subprocess.check_call(["gzip", "-6", "-f", fq_path])     # the reference's inputs: `gzip -6`
    else:
        p1 = subprocess.Popen(cmd, stdout=subprocess.PIPE)
        p2 = subprocess.Popen([corpus.PPGZIP, "-l", "6", "-", gz_path], stdin=p1.stdout)
        p1.stdout.close()
        assert p2.wait() == 0 and p1.wait() == 0
    return gz_path


def test_baseline_config1_1M_reads_gzip6_chunk10000(device):
    """BASELINE config 1: Generator seed 0, 1 M reads x 150 bp, `gzip -6`, chunk 10 000 — the
    reference's own CPU-runnable case.  Must reproduce SURVEY.md §8c's known answers (381 111 160 B,
    md5 4e84..., 99 points / 98 chunks, 4 209-10 284 records per chunk) and equal the oracle on every
    chunk: chunk boundaries, inflated length, bytes, record count, the nine integers of every record."""
    import parallelparsing_b200 as pp
    gz = np.fromfile(_generator_file(1_000_000, system_gzip=True), np.uint8)
    ix = pp.Core.BuildDeflateIndex(gz, 10_000)
    ox = O.OracleIndex.build(gz, 10_000)
    assert ix.Count == ox.count == 99
    so, si, sb, sl = ix.scalars()
    assert list(so) == ox.outputs() and list(si) == ox.inputs()
    info = _digest_parity(pp, device, gz, ix, ox)
    assert info.total_bytes == 381_111_160 and info.n_chunks == 98
    job = pp.Job(device, ix, gz.size)
    job.run(gz)
    recs = [job.chunk(k).records for k in range(98)]
    assert (min(recs), max(recs)) == (4209, 10284) and sum(recs) == info.total_records >= 1_000_000
    assert hashlib.md5(job.all_bytes().tobytes()).hexdigest() == "4e840faad7d4a77e3a8bd25b0f47dd23"
    job.free()


def test_baseline_config2_10M_reads_chunk10000_every_chunk(device):
    """BASELINE config 2 at full size (10 M reads x 150 bp, 3.8 GB inflated, ~977 chunks): every chunk
    against the oracle (boundaries from the oracle's own CreateIndex), in staged and in pull mode."""
    import parallelparsing_b200 as pp
    gz_np = np.fromfile(_generator_file(10_000_000), np.uint8)
    gz, ptr = pp.pinned_copy(gz_np)
    ix = pp.Core.BuildDeflateIndex(gz_np, 10_000)
    ox = O.OracleIndex.build(gz_np, 10_000)
    assert list(ix.scalars()[1]) == ox.inputs()
    info = _digest_parity(pp, device, gz, ix, ox)
    assert info.total_records >= 10_000_000 and info.n_chunks > 900
    _digest_parity(pp, device, gz, ix, ox, zero_copy=True, gz_ptr=ptr)
    pp.lib().pp_host_free(ptr)


def test_baseline_config4_long_reads_capped_chunk1000(device):
    """BASELINE config 4 (reference-legal variant, quirk H2): lognormal lengths, mean 10 kbp, sigma 0.5,
    capped at 16 000 bp, chunk 1 000 — ~20 MB chunks, long matches, records of up to 32 KB."""
    import parallelparsing_b200 as pp
    gz = np.fromfile(_generator_file(30_000, lognormal=(10000, 0.5), cap=16000), np.uint8)
    ix = pp.Core.BuildDeflateIndex(gz, 1000)
    ox = O.OracleIndex.build(gz, 1000)
    info = _digest_parity(pp, device, gz, ix, ox)
    assert info.n_chunks >= 25 and info.total_bytes > 500_000_000


def test_baseline_config4_long_reads_uncapped_chunk1000(device):
    """BASELINE config 4 as stated (uncapped): records above 32 768 B make the reference's CreateIndex
    throw (quirk H2), so the cap is lifted on both sides — a documented extension."""
    import parallelparsing_b200 as pp
    gz = np.fromfile(_generator_file(20_000, lognormal=(10000, 0.5)), np.uint8)
    with pytest.raises(pp.ZException):
        pp.Core.BuildDeflateIndex(gz, 1000)
    ix = pp.Core.BuildDeflateIndex(gz, 1000, lift_record_cap=True)
    ox = O.OracleIndex.build(gz, 1000, True)
    _digest_parity(pp, device, gz, ix, ox)


@pytest.mark.parametrize("chunk", [1000, 10_000, 100_000])
def test_baseline_config5_chunk_sweep_every_chunk(device, chunk):
    """BASELINE config 5's chunk-size sweep on a 1 M-read file: per-chunk oracle comparison at every
    chunk size (781 / 99 / 11 points), not totals only."""
    import parallelparsing_b200 as pp
    gz = np.fromfile(_generator_file(1_000_000), np.uint8)
    ix = pp.Core.BuildDeflateIndex(gz, chunk)
    ox = O.OracleIndex.build(gz, chunk)
    info = _digest_parity(pp, device, gz, ix, ox)
    assert info.total_bytes == 381_111_160


def test_gpu_records_equal_naive_parser_field_by_field(device):
    """Second, independent pin: on well-formed input the GPU's records (PP_JOB_STRICT: without the H1
    duplicates) are, field by field, the records of the reference's OTHER parser
    (SimpleDecompressor/Parsing.cs:9-49, restated in the oracle) run over the whole stream."""
    import parallelparsing_b200 as pp
    fq = corpus.fastq(200_000)  # native U[128,512) lengths
    gz = corpus.gz_parallel(fq, 6, segment=4 << 20, threads=4)
    data = np.frombuffer(fq, np.uint8)
    n_naive, naive = O.naive_records(np.concatenate([data, np.zeros(65536, np.uint8)]))
    assert n_naive == 200_000
    ix = pp.Core.BuildDeflateIndex(gz, 3000)
    job = pp.Job(device, ix, gz.size, strict=True)
    info = job.run(gz)
    assert info.status == 0 and info.total_records == n_naive
    l0, l1, l2, l3 = [x.astype(np.int64) for x in job.line_starts()]
    outs, _, _, offl = ix.scalars()
    base = np.zeros(info.total_records, np.int64)   # stream offset of combined-memory index 0, per record
    pend = np.zeros(info.total_records, np.int64)
    for k in range(info.n_chunks):
        c = job.chunk(k)
        base[c.record_base: c.record_base + c.records] = outs[k] - offl[k]
        pend[c.record_base: c.record_base + c.records] = c.parse_end
    nxt = np.concatenate([l0[1:], [0]])
    last = np.ones(info.total_records, bool)
    for k in range(info.n_chunks):
        c = job.chunk(k)
        if c.records:
            last[c.record_base: c.record_base + c.records - 1] = False
    nxt = np.where(last, pend, nxt)
    got = np.stack([l0 + 1 + base, l1 - l0 - 2, l1 + base, l2 - l1 - 1, l2 + 1 + base, l3 - l2 - 2, l3 + base,
                    nxt - l3 - 1], axis=1)
    assert np.array_equal(got, naive)
    job.free()


def test_zero_copy_reads_nothing_past_a_registered_buffer(device):
    """PP_JOB_ZEROCOPY on a caller buffer pinned with pp_host_register that ENDS at the end of a
    mapping (the page behind it is inaccessible): the kernels may read exactly gz_len bytes."""
    import ctypes as C
    import mmap
    import parallelparsing_b200 as pp
    fq = corpus.fastq(20000, fixed=150)
    gz_np = corpus.gz_member(fq, 6)
    for trim in (0, 1, 5, 13):   # ragged tails: the buffer ends 0..15 bytes past a 16-byte boundary
        n = gz_np.size - trim
        pages = (n + mmap.PAGESIZE - 1) // mmap.PAGESIZE
        mm = mmap.mmap(-1, (pages + 1) * mmap.PAGESIZE)
        addr = C.addressof(C.c_char.from_buffer(mm))
        libc = C.CDLL(None, use_errno=True)
        assert libc.mprotect(C.c_void_p(addr + pages * mmap.PAGESIZE), C.c_size_t(mmap.PAGESIZE), 0) == 0  # PROT_NONE
        start = pages * mmap.PAGESIZE - n
        view = np.frombuffer(mm, np.uint8, n, start)
        view[:] = gz_np[:n]
        ix = pp.Core.BuildDeflateIndex(gz_np, 1000)
        pp.check(pp.lib().pp_host_register(C.c_void_p(addr + start), n))
        try:
            # all chunks but (for trimmed buffers) the last, whose input is cut short
            nch = ix.Count - 1 - (1 if trim else 0)
            job = pp.Job(device, ix, n, 0, nch, zero_copy=True)
            job.upload(C.c_void_p(addr + start)); job.execute(); job.download()
            info = job.info()
            assert info.status == 0
            outs = ix.scalars()[0]
            assert job.all_bytes().tobytes() == fq[: int(outs[nch])]
            job.free()
            if trim:  # the last chunk alone: its input ends early -> the reference's DATA_ERROR (Core.cs:174), no fault
                job = pp.Job(device, ix, n, ix.Count - 2, 1, zero_copy=True)
                job.upload(C.c_void_p(addr + start)); job.execute(); job.download()
                assert job.info().status in (0, -3)
                job.free()
        finally:
            pp.lib().pp_host_unregister(C.c_void_p(addr + start))
        del view
        mm.close()


# ----------------------------------------------------------------------------------------------
# Runtime modes: pipelined upload, compact windows, streamed download, multi-GPU call
# ----------------------------------------------------------------------------------------------

def _job_signature(job, info):
    """Everything a mode must reproduce: per-chunk (status, inflated, records, base, parse_end) + digests."""
    bd, fd = job.digests()
    rows = []
    for k in range(info.n_chunks):
        c = job.chunk(k)
        rows.append((c.status, c.inflated, c.records, c.record_base, c.parse_end, int(bd[k]), int(fd[k])))
    return rows


@pytest.mark.parametrize("chunk", [300, 2000])
def test_upload_modes_agree(device, chunk):
    """Plain staged, pipelined (PP_JOB_PIPELINE), pull (PP_JOB_ZEROCOPY), each with and without
    PP_JOB_COMPACT_WINDOWS, over several back-to-back steps: identical per-chunk results and digests,
    and the concatenation is the generator's output."""
    import parallelparsing_b200 as pp
    fq = corpus.fastq(40000, fixed=150)
    gz_np = corpus.gz_parallel(fq, 6, segment=2 << 20, threads=4)
    gz, ptr = pp.pinned_copy(gz_np)
    ix = pp.Core.BuildDeflateIndex(gz_np, chunk)
    ref_job = pp.Job(device, ix, gz.size)
    ref_info = ref_job.run(gz)
    assert ref_info.status == 0
    ref = _job_signature(ref_job, ref_info)
    h2d_plain = ref_info.h2d_bytes
    for kw in (dict(pipeline=True), dict(compact_windows=True), dict(pipeline=True, compact_windows=True),
               dict(zero_copy=True), dict(zero_copy=True, compact_windows=True),
               dict(pipeline=True, pageable=True), dict(pipeline=True, compact_windows=True, pageable=True)):
        # pinned source: the pipelined upload is hybrid (first wave pulled by the kernel); pageable: plain copies
        src = gz_np.ctypes.data_as(__import__("ctypes").c_void_p) if kw.pop("pageable", False) else ptr
        job = pp.Job(device, ix, gz.size, **kw)
        for step in range(3):   # back to back: the next upload must wait for the previous kernels
            job.upload(src); job.execute()
        job.download()
        info = job.info()
        assert info.status == 0, kw
        assert _job_signature(job, info) == ref, kw
        assert job.all_bytes().tobytes() == fq, kw
        if kw.get("compact_windows") and not kw.get("zero_copy"):
            assert info.h2d_bytes < h2d_plain - 20000 * (ix.Count - 2), kw   # windows really crossed compressed
        job.free()
    ref_job.free()
    pp.lib().pp_host_free(ptr)


def test_compact_windows_from_a_version1_index_file(device, tmp_path):
    """IndexIO version 1 (compact windows on disk) -> Deserialize -> PP_JOB_COMPACT_WINDOWS job == oracle."""
    import parallelparsing_b200 as pp
    fq = corpus.fastq(30000)
    gz = corpus.gz_member(fq, 6)
    p = str(tmp_path / "v1.gzi")
    pp.IndexIO.Serialize(pp.Core.BuildDeflateIndex(gz, 1000), p, compact=True)
    ix = pp.IndexIO.Deserialize(p)
    ox = O.OracleIndex.build(gz, 1000)
    job = pp.Job(device, ix, gz.size, compact_windows=True, pipeline=True)
    info = job.run(gz)
    assert info.status == 0 and info.n_chunks == ox.count - 1
    want = O.chunk_digests(gz, ox)
    bd, fd = job.digests()
    for k in range(info.n_chunks):
        c = job.chunk(k)
        assert (c.inflated, c.records, int(bd[k]), int(fd[k])) == tuple(int(x) for x in want[k]), f"chunk {k}"
    job.free()


def test_streamed_download_delivers_every_byte(device):
    """pp_job_execute_to_host: the inflated stream arrives in pinned host memory while the decode runs."""
    import ctypes as C
    import parallelparsing_b200 as pp
    fq = corpus.fastq(60000, fixed=150)
    gz_np = corpus.gz_parallel(fq, 6, segment=2 << 20, threads=4)
    gz, ptr = pp.pinned_copy(gz_np)
    ix = pp.Core.BuildDeflateIndex(gz_np, 1000)
    dst = C.c_void_p()
    pp.check(pp.lib().pp_host_alloc(len(fq), C.byref(dst)))
    view = np.ctypeslib.as_array(C.cast(dst, C.POINTER(C.c_uint8)), shape=(len(fq),))
    for kw in (dict(), dict(pipeline=True, compact_windows=True), dict(zero_copy=True)):
        job = pp.Job(device, ix, gz.size, **kw)
        for step in range(2):
            view[:] = 0
            job.upload(ptr)
            job.execute_to_host(dst, len(fq))
            assert view.tobytes() == fq, kw
        job.download()
        assert job.info().status == 0 and job.info().total_records >= 60000
        job.free()
    with pytest.raises(pp.ZException):
        job = pp.Job(device, ix, gz.size)
        job.upload(ptr)
        job.execute_to_host(dst, len(fq) - 1)   # too small: PP_BUF_ERROR, nothing written past the end
    pp.lib().pp_host_free(dst)
    pp.lib().pp_host_free(ptr)


def test_multi_gpu_call_equals_whole_file_job(device):
    """pp_decompress_all_multi over every GPU of the box (one on the single-GPU test box: the same code
    path with one part; the 2-GPU run exercises two contexts and host threads) against the whole-file job:
    parts are contiguous, cover every chunk, record bases are the global prefix sums."""
    import torch
    import parallelparsing_b200 as pp
    ngpu = torch.cuda.device_count()
    fq = corpus.fastq(50000, fixed=150)
    gz_np = corpus.gz_parallel(fq, 6, segment=2 << 20, threads=4)
    gz, ptr = pp.pinned_copy(gz_np)
    ix = pp.Core.BuildDeflateIndex(gz_np, 1500)
    whole = pp.Job(device, ix, gz.size)
    wi = whole.run(gz)
    wsig = _job_signature(whole, wi)
    for devices in ([0], list(range(ngpu)), [0, 0, 0]):   # [0,0,0]: three parts on one GPU (three contexts)
        for kw in (dict(), dict(zero_copy=True, compact_windows=True), dict(pipeline=True)):
            m = pp.MultiGpuDecompressAll(devices, ix, gz, **kw)
            assert m.status == 0, (devices, kw, [m.part(r)[0].info().status for r in range(len(devices))])
            mi = m.info()
            assert (mi.n_parts, mi.n_chunks, mi.total_records, mi.total_bytes) == (len(devices), wi.n_chunks, wi.total_records, wi.total_bytes)
            parts = pp.partition_chunks(ix, len(devices))
            cat = []
            for r, (first, n) in enumerate(parts):
                job, dev, base = m.part(r)
                info = job.info()
                assert (info.first_chunk, info.n_chunks, dev) == (first, n, devices[r])
                bd, fd = job.digests() if n else ([], [])
                for k in range(n):
                    c = job.chunk(k)
                    w = wsig[first + k]
                    assert (c.status, c.inflated, c.records, c.record_base + base, c.parse_end, int(bd[k]), int(fd[k])) == w
                cat.append(job.all_bytes().tobytes() if n else b"")
            assert b"".join(cat) == fq
            m.free()
    whole.free()
    pp.lib().pp_host_free(ptr)


# ----------------------------------------------------------------------------------------------
# GPU-assisted CreateIndex, first slice: deflate block boundaries (pp_scan_blocks) == zlib's Z_BLOCK stops
# ----------------------------------------------------------------------------------------------

@pytest.mark.parametrize("mode,segment", [("dynamic6", 0), ("dynamic6", 65536), ("dynamic1", 32768), ("syncflush", 65536),
                                           ("fixed", 200000), ("stored", 100000), ("huffman", 65536)])
def test_scan_blocks_equals_zlib_block_stops(device, mode, segment):
    """Every block's first bit (8*Input - Bits of a checkpoint taken there) and output offset, for streams
    of dynamic blocks (found speculatively, segments in parallel), and for stored / fixed blocks and
    sync-flush seams, which the search cannot see: those seams are re-walked until the chain closes."""
    import parallelparsing_b200 as pp
    kw = dict(dynamic6=dict(level=6), dynamic1=dict(level=1), fixed=dict(level=6, strategy=zlib.Z_FIXED), stored=dict(level=0),
              huffman=dict(level=6, strategy=zlib.Z_HUFFMAN_ONLY), syncflush=dict(level=6, flush_every=70000))[mode]
    gz = corpus.gz_member(corpus.fastq(40000, fixed=150), **kw)
    bits, outs, kinds, end, tot = O.block_stops(gz)
    gb, go, gend, gtot, ms, passes = pp.Core.ScanBlocks(gz, device, segment)
    assert gtot == tot and (gend + 7) // 8 * 8 + 64 == end
    assert np.array_equal(gb, bits) and np.array_equal(go, outs)
    if mode.startswith("dynamic"):
        assert passes == 1   # every seam closed at the first attempt


def test_scan_blocks_baseline_config1_file(device):
    """BASELINE config 1's file (1 M reads x 150 bp, gzip -6): 1 559 dynamic blocks found in parallel;
    the checkpoints CreateIndex chooses are a subset of them (Core.cs:98-109)."""
    import parallelparsing_b200 as pp
    gz = np.fromfile(_generator_file(1_000_000, system_gzip=True), np.uint8)
    bits, outs, kinds, end, tot = O.block_stops(gz)
    gb, go, gend, gtot, ms, passes = pp.Core.ScanBlocks(gz, device)
    assert len(bits) == 1559 and gtot == tot == 381_111_160
    assert np.array_equal(gb, bits) and np.array_equal(go, outs) and passes == 1
    ix = pp.Core.BuildDeflateIndex(gz, 10_000)
    so, si, sb, _ = ix.scalars()
    cps = set(zip((8 * si - sb).tolist(), so.tolist()))
    blocks = set(zip(gb.tolist(), go.tolist()))
    assert cps - {(int(8 * si[-1] - sb[-1]), int(so[-1]))} <= blocks   # every checkpoint but the end sentinel is a block start
    print(f"scan_blocks: {len(gb)} blocks of {gz.size/1e6:.1f} MB in {ms:.2f} ms kernel time = {tot/ms/1e6:.1f} GB/s inflated-equivalent")


# ----------------------------------------------------------------------------------------------
# Paired-end R1/R2 below the C ABI (BASELINE config 3)
# ----------------------------------------------------------------------------------------------

def _record_ids(job, index, first_chunk):
    """Identifier line (without '@') of every record of a job, from its bytes + line starts."""
    info = job.info()
    l0, l1, _, _ = job.line_starts()
    ids = []
    for k in range(info.n_chunks):
        c = job.chunk(k)
        mem = np.concatenate([index[first_chunk + k].offset, job.chunk_bytes(k)])
        for r in range(c.record_base, c.record_base + c.records):
            ids.append(bytes(mem[l0[r] + 1: l1[r] - 1]))
    return ids


@pytest.mark.parametrize("parts", [1, 2, 5])
def test_paired_end_mates_are_co_resident(device, parts):
    """R1/R2 from two Generator streams (same read count, different compression level and chunk size, so
    block boundaries and chunk record counts differ): every part's R1 records have their mates in the R2
    jobs of the same part, located by ordinal; ids are the generator's, pair by pair."""
    import parallelparsing_b200 as pp
    nreads = 24000
    r1, r2 = corpus.fastq(nreads, fixed=150, seed=0), corpus.fastq(nreads, fixed=150, seed=1)
    g1, g2 = corpus.gz_member(r1, 6), corpus.gz_member(r2, 1)
    i1, i2 = pp.Core.BuildDeflateIndex(g1, 1000), pp.Core.BuildDeflateIndex(g2, 700)
    pe = pp.PairedDecompressAll([0] * parts, i1, g1, i2, g2)
    info = pe.info()
    assert pe.status == 0 and (info.records_r1, info.records_r2, info.pairs, info.n_parts) == (nreads, nreads, nreads, parts)
    if parts > 1:
        assert info.topup_chunks > 0      # the seams of the two partitions never coincide here
    ids1, ids2 = [ln[1:] for ln in r1.split(b"\n")[0::4]], [ln[1:] for ln in r2.split(b"\n")[0::4]]
    p1 = pp.partition_chunks(i1, parts)
    seen = 0
    for g in range(parts):
        j1, base1, r2jobs = pe.part(g)
        n1 = j1.info().total_records
        got1 = _record_ids(j1, i1, p1[g][0])
        assert got1 == ids1[base1: base1 + n1]
        got2 = [(_record_ids(j, i2, j.info().first_chunk), b) for j, b in r2jobs]
        for b, (ids, base) in zip([b for _, b in r2jobs], got2):
            assert ids == ids2[base: base + len(ids)]      # every R2 job knows its true global ordinals
        for r in (0, n1 // 2, n1 - 1) if n1 else ():
            w, idx = pe.locate(g, base1 + r)
            assert got2[w][0][idx] == ids2[base1 + r]      # the mate of R1 record base1+r, on the same part
        # every ordinal of the part's R1 range is covered by the part's R2 jobs
        cover = set()
        for ids, base in got2:
            cover.update(range(base, base + len(ids)))
        assert set(range(base1, base1 + n1)) <= cover
        seen += n1
    assert seen == nreads
    pe.free()


def test_scan_blocks_on_damaged_and_foreign_input(device):
    """pp_scan_blocks must end cleanly, never hang: on a damaged stream, on a file that is not gzip at all
    (rejected by the header check) and on a stream cut short (no final block: an error)."""
    import parallelparsing_b200 as pp
    gz = corpus.gz_member(corpus.fastq(30000, fixed=150), 6)
    bits, outs, kinds, end, tot = O.block_stops(gz)
    bad = gz.copy()
    mid = int(bits[len(bits) // 2] // 8) + 2000
    bad[mid: mid + 4000] = np.random.default_rng(1).integers(0, 256, 4000, dtype=np.uint8)
    # damage in the middle: Huffman garbage is mostly decodable (zlib itself only notices through the CRC-32 of
    # the trailer, which a scan that produces no bytes cannot check), so the scan may end either way — but it
    # must END, with an error code or with a block list that agrees with the true one up to the damage
    try:
        gb, go, gend, gtot, ms, passes = pp.Core.ScanBlocks(bad, device, 65536)
        nb = int(np.searchsorted(bits, mid * 8))
        assert np.array_equal(gb[:nb], bits[:nb]) and np.array_equal(go[:nb], outs[:nb])
    except pp.ZException as e:
        assert e.Code in (-3, -5)
    for cut in range(gz.size // 2, gz.size // 2 + 17):       # cut short: no final block; every length mod 16
        with pytest.raises(pp.ZException) as e:
            pp.Core.ScanBlocks(gz[:cut], device, 65536)
        assert e.value.Code in (-3, -5), cut
    with pytest.raises(pp.ZException) as e:
        pp.Core.ScanBlocks(np.frombuffer(b"this is not a gzip file, not even close" * 100, np.uint8), device)
    assert e.value.Code == -3
    # and the undamaged stream still scans
    gb, go, gend, gtot, ms, passes = pp.Core.ScanBlocks(gz, device, 65536)
    assert np.array_equal(gb, bits) and gtot == tot


# ----------------------------------------------------------------------------------------------
# CreateIndex on the GPU (pp_index_create_gpu) == the oracle's Core.BuildDeflateIndex, byte for byte
# ----------------------------------------------------------------------------------------------

def _same_index_file(pp, ix, ox, tmp_path, tag):
    """Both indexes through the reference's own file format (IndexIO.Serialize): identical bytes."""
    a, b = str(tmp_path / f"{tag}_gpu.gzi"), str(tmp_path / f"{tag}_ora.gzi")
    pp.IndexIO.Serialize(ix, a)
    ox.serialize(b)
    fa, fb = np.fromfile(a, np.uint8), np.fromfile(b, np.uint8)
    assert fa.size == fb.size and np.array_equal(fa, fb), tag
    assert ix.ChunkMaxBytes == ox.chunk_max_bytes


def _mixed_block_stream(data: bytes) -> np.ndarray:
    """One gzip member whose deflate blocks alternate between stored, fixed-codes and dynamic every few KB
    (raw deflate pieces joined with full flushes; the last piece carries BFINAL)."""
    import struct
    rng = np.random.default_rng(7)
    out = [b"\x1f\x8b\x08\x00\x00\x00\x00\x00\x00\x03"]
    pos, k = 0, 0
    comp = None
    while pos < len(data):
        n = int(rng.integers(2000, 90000))
        piece = data[pos: pos + n]
        pos += n
        level, strategy = [(0, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_FIXED), (6, zlib.Z_DEFAULT_STRATEGY)][k % 3]
        k += 1
        # one compressor per piece would lose the history: use ONE compressor and switch its parameters is not in
        # Python's zlib, so pieces are independent streams with an empty history -> legal in one member as long as
        # no piece is final before the last: Z_FULL_FLUSH ends each on a byte boundary without BFINAL
        c = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
        body = c.compress(piece) + (c.flush(zlib.Z_FINISH) if pos >= len(data) else c.flush(zlib.Z_FULL_FLUSH))
        out.append(body)
    out.append(struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data) & 0xFFFFFFFF))
    return np.frombuffer(b"".join(out), np.uint8)


@pytest.mark.parametrize("mode,chunk", [("dynamic6", 1000), ("dynamic6", 100), ("dynamic1", 1000), ("dynamic9", 5000),
                                        ("syncflush", 100), ("fixed", 1000), ("stored", 200), ("huffman", 1000),
                                        ("native_lengths", 1000), ("tiny", 1000), ("chunk_lt8", 5), ("lognormal", 20),
                                        ("empty", 1000), ("one_byte", 1000), ("zeros_lift", 1000), ("repeat32k", 50),
                                        ("binary_at", 300), ("mixed_members_of_blocks", 40)])
def test_gpu_create_index_equals_oracle(device, tmp_path, mode, chunk):
    """Points (Input, Bits, Output), 32 KB windows, offsets and ChunkMaxBytes of the index built on the GPU
    equal the oracle's serial inflate(Z_BLOCK) pass: the serialized files are identical.  Streams of dynamic,
    fixed and stored blocks, sync-flush seams (empty stored blocks), files shorter than one window, chunksize
    < 8 (the uint wrap of Core.cs:105), and long lognormal reads with the record cap lifted."""
    import parallelparsing_b200 as pp
    kw = dict(dynamic6=dict(level=6), dynamic1=dict(level=1), dynamic9=dict(level=9), fixed=dict(level=6, strategy=zlib.Z_FIXED),
              stored=dict(level=0), huffman=dict(level=6, strategy=zlib.Z_HUFFMAN_ONLY), syncflush=dict(level=6, flush_every=70000),
              native_lengths=dict(level=6), tiny=dict(level=6), chunk_lt8=dict(level=6), lognormal=dict(level=6)).get(mode, dict(level=6))
    lift = mode in ("lognormal", "zeros_lift", "repeat32k", "binary_at", "mixed_members_of_blocks")
    rng = np.random.default_rng(11)
    if mode == "empty":
        data = b""
    elif mode == "one_byte":
        data = b"@"
    elif mode == "zeros_lift":                      # ratio ~1000:1: a few blocks of many megabytes, no '@' at all
        data = bytes(12_000_000)
    elif mode == "repeat32k":                       # matches at the far end of the window, across every segment seam
        unit = b"@" + bytes(rng.integers(65, 91, 32767, dtype=np.uint8))
        data = unit * 40 + corpus.fastq(500, fixed=150) + unit * 10
    elif mode == "binary_at":                       # incompressible bytes ('@' is 1 in 256 of them) between text
        data = corpus.fastq(3000, fixed=150) + bytes(rng.integers(0, 256, 3_000_000, dtype=np.uint8)) + corpus.fastq(3000, fixed=150)
    elif mode == "mixed_members_of_blocks":         # stored, fixed and dynamic blocks alternating every few KB
        data = corpus.fastq(20000, fixed=150)
    elif mode == "native_lengths":
        data = corpus.fastq(30000)
    elif mode == "tiny":
        data = corpus.fastq(20, fixed=150)
    elif mode == "lognormal":
        data = corpus.fastq(600, lognormal=(10000, 0.5))
    else:
        data = corpus.fastq(40000, fixed=150)
    if mode == "mixed_members_of_blocks":
        gz = _mixed_block_stream(data)
    else:
        gz = corpus.gz_member(data, **kw)
    ox = O.OracleIndex.build(gz, chunk, lift)
    ix, st = pp.Core.BuildDeflateIndexGpu(gz, chunk, device, lift_record_cap=lift, want_stats=True)
    assert st["total_out"] == len(data) and st["points"] == ox.count
    _same_index_file(pp, ix, ox, tmp_path, mode)


def test_gpu_create_index_then_decompress_all(device):
    """The index built on the GPU drives DecompressAll: every chunk's bytes and records equal the oracle's."""
    import parallelparsing_b200 as pp
    gz = corpus.gz_member(corpus.fastq(40000, fixed=150), 6)
    ix = pp.Core.BuildDeflateIndexGpu(gz, 1000, device)
    ox = O.OracleIndex.build(gz, 1000)
    _digest_parity(pp, device, gz, ix, ox)


def test_gpu_create_index_baseline_config1_file(device, tmp_path):
    """BASELINE config 1's file (1 M reads x 150 bp, system gzip -6, chunk 10 000): 99 points, identical file."""
    import parallelparsing_b200 as pp
    gz = np.fromfile(_generator_file(1_000_000, system_gzip=True), np.uint8)
    ox = O.OracleIndex.build(gz, 10_000)
    ix, st = pp.Core.BuildDeflateIndexGpu(gz, 10_000, device, want_stats=True)
    assert ox.count == 99 and st["blocks"] == 1559 and st["total_out"] == 381_111_160
    _same_index_file(pp, ix, ox, tmp_path, "c1")


def test_gpu_create_index_errors_like_zlib(device):
    """What zlib reports as Z_DATA_ERROR in the reference's pass is reported here too: a wrong CRC-32, a wrong
    ISIZE, a cut file, a distance that reaches in front of the stream.  An over-long record: -104 as the host
    CreateIndex.  A second member: PP_E_UNSUPPORTED (-106), the host path handles it."""
    import parallelparsing_b200 as pp
    data = corpus.fastq(20000, fixed=150)
    gz = corpus.gz_member(data, 6)

    def code(buf, chunk=1000, **kw):
        try:
            pp.Core.BuildDeflateIndexGpu(buf, chunk, device, **kw)
            return 0
        except pp.ZException as e:
            return e.Code

    assert code(gz) == 0
    bad = gz.copy(); bad[-6] ^= 1                        # CRC-32
    assert code(bad) == -3
    bad = gz.copy(); bad[-2] ^= 1                        # ISIZE
    assert code(bad) == -3
    bad = gz.copy(); bad[gz.size // 2] ^= 0x10           # one bit in the middle: decodes to other bytes or breaks the stream
    assert code(bad) in (-3, -5)
    assert code(gz[: gz.size // 2]) in (-3, -5)
    assert code(gz[:-1]) in (-3, -5)
    assert code(np.concatenate([gz, gz])) == -106
    assert code(np.concatenate([gz, np.zeros(3, np.uint8)])) == -106
    assert code(np.frombuffer(b"not a gzip file at all, just text" * 10, np.uint8)) == -3
    # a flush every 40 bytes: two blocks per ~35 compressed bytes, more than the scan's record areas hold -> declined
    tiny = corpus.gz_member(data[:300000], 6, flush_every=40)
    try:
        ix = pp.Core.BuildDeflateIndexGpu(tiny, 100, device)
        ox = O.OracleIndex.build(tiny, 100)
        assert ix.Count == ox.count and all(np.array_equal(ix[i].Window, ox.point(i)["window"]) for i in range(ox.count))
    except pp.ZException as e:
        assert e.Code == -106
    # a fixed-codes block whose first symbol is a match: distance 1 with nothing in front of it
    hdr = b"\x1f\x8b\x08\x00\x00\x00\x00\x00\x00\x03"
    # bits: BFINAL=1, BTYPE=01, length code 257 (len 3) = 0000001, distance code 0 = 00000, EOB = 0000000
    bits = "1" + "10" + "0000001" + "00000" + "0000000"
    val = 0
    for i, ch in enumerate(bits):
        val |= int(ch) << i
    body = val.to_bytes((len(bits) + 7) // 8, "little")
    import struct
    far = np.frombuffer(hdr + body + struct.pack("<II", 0, 3), np.uint8)
    with pytest.raises(zlib.error, match="too far back"):
        zlib.decompress(far.tobytes(), 47)
    assert code(far) == -3
    # the record cap
    long_rec = b"@r\n" + b"A" * 20000 + b"\n+\n" + b"?" * 20000 + b"\n"
    gzl = corpus.gz_member(data[:3000] + long_rec + data[:3000], 6)
    assert code(gzl, 10) == -104
    assert code(gzl, 10, lift_record_cap=True) == 0

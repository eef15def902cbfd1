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

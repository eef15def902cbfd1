"""CPU tests (no GPU): the oracle against the golden vectors and against zlib itself;
the quirks of the reference that the oracle must reproduce."""
import ctypes as C
import hashlib
import json
import os
import zlib

import numpy as np
import pytest

import corpus
import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
G = json.load(open(os.path.join(GOLD, "golden.json")))


def _tools():
    L = C.CDLL(os.path.join(corpus.ROOT, "tools", "_build", "libpptools.so"))
    L.ppgen_kat_next.restype = C.c_int32
    L.ppgen_kat_next_range.restype = C.c_int32
    L.ppgen_kat_next_range.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_double)]
    return L


def test_dotnet_random_known_answers():
    """System.Random(seed) (Generator/Generator.cs:8): SURVEY.md §8c known answers."""
    L = _tools()
    assert L.ppgen_kat_next(0) == 1559595546
    assert L.ppgen_kat_next(42) == 1434747710
    d = C.c_double(0)
    assert L.ppgen_kat_next_range(0, 128, 512, C.byref(d)) == G["dotnet_random"]["Random(0).Next(128,512)"] == 406
    assert d.value == G["dotnet_random"]["then NextDouble()"]


def test_generator_golden():
    fq = corpus.fastq(600)
    assert fq.split(b"\n", 1)[0].decode() == "@SRR18173253.1.1 1 length=406" == G["generator"]["first_line_seed0"]
    assert len(fq) == G["generator"]["bytes_600_native"]
    assert hashlib.md5(fq).hexdigest() == G["generator"]["md5_600_native"]
    assert hashlib.md5(corpus.fastq(20000, fixed=150)).hexdigest() == G["generator"]["md5_20000_fixed150"]
    # Generator.cs:14-18,53-56: four lines per record, ACGT only, quality alphabet ?*!
    lines = fq.split(b"\n")
    assert len(lines) == 600 * 4 + 1 and lines[-1] == b""
    assert set(lines[1]) <= set(b"ACGT") and set(lines[3]) <= set(b"?*!") and len(lines[1]) == len(lines[3]) == 406
    assert lines[2].startswith(b"+SRR") and lines[2].endswith(b".1.1 1 length=406")


def _golden_inputs():
    gz = np.fromfile(os.path.join(GOLD, "gen600.fastq.gz"), np.uint8)
    assert hashlib.md5(gz.tobytes()).hexdigest() == G["index"]["gz_md5"]
    return gz


def test_oracle_index_matches_golden():
    gz = _golden_inputs()
    ox = O.OracleIndex.build(gz, G["index"]["chunksize"])
    assert ox.count == G["index"]["points"] and ox.chunk_max_bytes == G["index"]["chunk_max_bytes"]
    end = ox.point(ox.count - 1)
    assert (end["output"], end["input"]) == (G["index"]["end_output"], G["index"]["end_input"])
    for k, c in enumerate(G["chunks"]):
        p = ox.point(k)
        assert (p["output"], p["input"], p["bits"], p["offset"].size) == (c["output"], c["input"], c["bits"], c["offset_len"])


def test_oracle_serialize_matches_golden_bytes(tmp_path):
    """IndexIO.Serialize (Common/IndexIO.cs:7-27): byte-identical file; Deserialize round trip."""
    gz = _golden_inputs()
    ox = O.OracleIndex.build(gz, G["index"]["chunksize"])
    p = str(tmp_path / "a.gzi")
    ox.serialize(p)
    ref = open(os.path.join(GOLD, "gen600.chunk50.gzi"), "rb").read()
    assert open(p, "rb").read() == ref and hashlib.md5(ref).hexdigest() == G["index"]["gzi_md5"]
    ox2 = O.OracleIndex.deserialize(p)
    p2 = str(tmp_path / "b.gzi")
    ox2.serialize(p2)
    got = open(p2, "rb").read()
    # Deserialize discards ChunkMaxBytes (IndexIO.cs:35,52: quirk H7): bytes 4..8 differ, the rest is identical
    assert got[:4] == ref[:4] and got[8:] == ref[8:]
    # format: int32 0, int32 ChunkMaxBytes, int32 Count, then per point i64 i64 i32 i32(32768) window i32 offset
    hdr = np.frombuffer(ref[:12], "<i4")
    assert hdr[0] == 0 and hdr[1] == G["index"]["chunk_max_bytes"] and hdr[2] == G["index"]["points"]
    assert np.frombuffer(ref[12 + 20:12 + 24], "<i4")[0] == 32768


def test_oracle_chunks_match_golden_and_zlib():
    gz = _golden_inputs()
    ox = O.OracleIndex.deserialize(os.path.join(GOLD, "gen600.chunk50.gzi"))
    cat, total = [], 0
    for k, c in enumerate(G["chunks"]):
        n, recs, buf, _ = O.chunk(gz, ox, k)
        assert n == c["records"] and buf.size == c["inflated"]
        assert hashlib.md5(buf.tobytes()).hexdigest() == c["bytes_md5"]
        assert hashlib.md5(recs.astype("<i8").tobytes()).hexdigest() == c["fields_md5"]
        cat.append(buf.tobytes())
        total += n
    assert total == G["total_records"] == 600
    assert b"".join(cat) == zlib.decompress(gz.tobytes(), 31) == corpus.fastq(600)


@pytest.mark.parametrize("level,chunk", [(1, 300), (6, 1000), (9, 2000)])
def test_concat_chunks_equals_zcat(level, chunk):
    fq = corpus.fastq(6000, fixed=150)
    gz = corpus.gz_member(fq, level)
    ox = O.OracleIndex.build(gz, chunk)
    cat = b"".join(O.extract(gz, ox, k).tobytes() for k in range(ox.count - 1))
    assert cat == fq
    n, b = O.decompress_all_mt(gz, ox, threads=2)
    assert (n, b) == (6000, len(fq))
    assert O.naive_count(gz) == (6000, len(fq))


def test_quirk_h1_duplicate_record_on_boundary():
    """A checkpoint exactly on a record boundary makes the next chunk re-emit the last record
    (Core.cs:86-94,107 + BatchedFASTQ.cs:68)."""
    fq = corpus.fastq(400, fixed=150)
    lines = fq.split(b"\n")[:-1]
    co = zlib.compressobj(6, zlib.DEFLATED, 31)
    parts = []
    for i in range(400):
        parts.append(co.compress(b"\n".join(lines[4 * i:4 * i + 4]) + b"\n"))
        parts.append(co.flush(zlib.Z_SYNC_FLUSH))
    parts.append(co.flush())
    gz = np.frombuffer(b"".join(parts), np.uint8).copy()
    ox = O.OracleIndex.build(gz, 20)
    total = sum(O.chunk(gz, ox, k)[0] for k in range(ox.count - 1))
    assert total == 400 + (ox.count - 2)  # one extra per interior checkpoint


def test_quirk_h2_record_longer_than_window():
    fq = corpus.fastq(12, fixed=20000)
    gz = corpus.gz_member(fq, 6)
    with pytest.raises(RuntimeError):
        O.OracleIndex.build(gz, 2)
    ox = O.OracleIndex.build(gz, 2, lift_cap=True)  # documented extension
    assert sum(O.chunk(gz, ox, k)[0] for k in range(ox.count - 1)) >= 12


def test_quirk_chunksize_below_8_has_no_interior_points():
    """chunksize-8 is a uint (Core.cs:105): it wraps, so only the first point and the end sentinel exist."""
    gz = corpus.gz_member(corpus.fastq(3000, fixed=150), 6)
    assert O.OracleIndex.build(gz, 4).count == 2


def test_quirk_h5_multi_member_is_truncated():
    fq = corpus.fastq(400, fixed=150)
    one = corpus.gz_member(fq, 6)
    gz = np.concatenate([one, one])
    ox = O.OracleIndex.build(gz, 100000)
    total = sum(O.chunk(gz, ox, k)[0] for k in range(ox.count - 1))
    assert total == 400  # raw inflate stops at the first member's final block (Core.cs:185)


def test_parse_stop_rules():
    """Parsing.cs:16,24,28,34,38: stop at NUL / incomplete line; '@' and '+' skipped unchecked."""
    z = np.zeros(64, np.uint8)

    def run(pre, rest):
        r = np.concatenate([np.frombuffer(rest, np.uint8), z])
        return O.parse(np.frombuffer(pre, np.uint8), r)[0]
    assert run(b"", b"@a\nAC\n+\n??\n") == 1
    assert run(b"@a\nAC\n", b"+\n??\n@b\nGG\n+\n!!\n") == 2
    assert run(b"", b"@a\nAC\n+\n??") == 0          # last line has no '\n'
    assert run(b"", b"@a\nAC\n+\n??\n@b\nGG\n") == 1  # trailing partial record dropped
    assert run(b"", b"xa\nAC\nx\n??\n") == 1          # first bytes of lines 0 and 2 are not checked
    assert run(b"", b"") == 0


def test_hot_path_parser_equals_naive_parser_on_well_formed_input():
    """Independent second pin of the PARSE half: the restatement of the hot-path parser
    (Decompressor/Parsing.cs:11-69, chunk by chunk over CombinedMemory) and the restatement of the
    Naive parser (SimpleDecompressor/Parsing.cs:9-49, one pass over the whole stream) are two different
    algorithms of the reference; on well-formed FASTQ they must cut the same four fields out of the
    same bytes, record for record.  Records are compared in stream coordinates (chunk k's combined
    memory starts |offset_k| bytes before Output_k); the H1 duplicate (a complete record in
    Point.offset) is the one documented difference and is dropped before comparing."""
    fq = corpus.fastq(30000)  # Generator's native U[128,512) lengths
    gz = corpus.gz_member(fq, 6)
    data = np.frombuffer(fq, np.uint8)
    n_naive, naive = O.naive_records(np.concatenate([data, np.zeros(65536, np.uint8)]))
    assert n_naive == 30000
    ox = O.OracleIndex.build(gz, 1000)
    outs = ox.outputs()
    got = []
    for k in range(ox.count - 1):
        n, recs, buf, _ = O.chunk(gz, ox, k)
        off = ox.point(k)["offset"]
        base = outs[k] - off.size
        r = recs.copy()
        if k > 0 and off.size and off[-1] == 10 and int((off == 10).sum()) == 4:
            r = r[1:]  # quirk H1: the previous chunk already yielded this record
        # idnFrom, idnLen, seqFrom, seqLen, plsFrom, plsLen, qltFrom, qltLen in stream coordinates
        g = r[:, 1:9].copy()
        g[:, 0::2] += base
        got.append(g)
    got = np.concatenate(got)
    assert got.shape == naive.shape and np.array_equal(got, naive)
    # and the fields really are the generator's lines
    lines = fq.split(b"\n")
    for r in (0, 1, 12345, 29999):
        f = naive[r]
        assert fq[f[0]:f[0] + f[1]] == lines[4 * r][1:] and fq[f[2]:f[2] + f[3]] == lines[4 * r + 1]
        assert fq[f[4]:f[4] + f[5]] == lines[4 * r + 2][1:] and fq[f[6]:f[6] + f[7]] == lines[4 * r + 3]


def test_digests_are_order_and_position_sensitive_and_match_numpy():
    """The oracle's digests (the functions the GPU library restates, pp_job_digests) against a direct
    numpy evaluation of their definition; swaps and shifts must change them."""
    def mix(x):
        x = (x + np.uint64(0x9E3779B97F4A7C15))
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))
    rng = np.random.default_rng(0)
    with np.errstate(over="ignore"):
        for n in (0, 1, 7, 8, 9, 15, 16, 17, 1000, 4099):
            a = rng.integers(0, 256, n, dtype=np.uint8)
            pad = np.concatenate([a, np.zeros((-n) % 8, np.uint8)]).view("<u8")
            want = int(mix(np.array([n], np.uint64))[0] + (mix(np.arange(pad.size, dtype=np.uint64)) * (pad + np.uint64(1))).sum(dtype=np.uint64))
            assert O.digest_bytes(a) == want, n
        recs = rng.integers(-5, 1 << 31, (300, 9)).astype(np.int64)
        want = int((mix(np.arange(2700, dtype=np.uint64)) * (recs.reshape(-1).view(np.uint64) + np.uint64(1))).sum(dtype=np.uint64))
        assert O.digest_fields(recs) == want
    a = rng.integers(0, 256, 100, dtype=np.uint8)
    b = a.copy(); b[[3, 4]] = b[[4, 3]]
    assert a[3] == a[4] or O.digest_bytes(a) != O.digest_bytes(b)
    assert O.digest_bytes(a) != O.digest_bytes(a[:-1]) and O.digest_bytes(np.zeros(8, np.uint8)) != O.digest_bytes(np.zeros(9, np.uint8))
    r2 = recs.copy(); r2[[0, 1]] = r2[[1, 0]]
    assert O.digest_fields(recs) != O.digest_fields(r2)


def test_chunk_digests_mt_equals_single_chunk_calls():
    gz = corpus.gz_member(corpus.fastq(9000, fixed=150), 6)
    ox = O.OracleIndex.build(gz, 1000)
    d = O.chunk_digests(gz, ox, threads=4)
    assert d.shape[0] == ox.count - 1
    for k in range(ox.count - 1):
        n, recs, buf, _ = O.chunk(gz, ox, k)
        assert (int(d[k, 0]), int(d[k, 1])) == (buf.size, n)
        assert int(d[k, 2]) == O.digest_bytes(buf) and int(d[k, 3]) == O.digest_fields(recs)

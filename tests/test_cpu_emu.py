"""CPU tests (no GPU) of the inflate kernel's LOGIC: parallelparsing_b200/csrc/inflate_core.cuh
compiled in PP_HOST_EMU mode (the CTA's threads run one after another, phase by phase) against
zlib through the oracle.  Covers every block type, several CTA sizes and sub-sequence widths,
unaligned checkpoints (Bits 1-7), long reads, the mid-block stop (output full) and corrupt
input.  The emulation is test scaffolding only; the product runs the same source on the GPU."""
import zlib

import numpy as np
import pytest

import corpus
import emu_lib as E
import oracle_lib as O


def _check(gz, chunksize, T, subw=31, lift=False):
    ox = O.OracleIndex.build(gz, chunksize, lift)
    outs, ins = ox.outputs(), ox.inputs()
    bits_seen = set()
    for k in range(ox.count - 1):
        p = ox.point(k)
        ref = O.extract(gz, ox, k)
        st, got, nl, mb, _ = E.inflate_chunk(gz, p["input"], p["bits"], ins[k + 1], p["window"], outs[k + 1] - outs[k], T, subw)
        assert st == 0, (k, st)
        assert got.size == ref.size and np.array_equal(got, ref), f"chunk {k}"
        assert nl == int((ref == 10).sum()) and mb == (0 if (ref == 0).any() else 1)
        bits_seen.add(p["bits"])
    return ox.count - 1, bits_seen


MODES = dict(dynamic6=dict(level=6), dynamic1=dict(level=1), dynamic9=dict(level=9),
             fixed=dict(level=6, strategy=zlib.Z_FIXED), stored=dict(level=0),
             huffman=dict(level=6, strategy=zlib.Z_HUFFMAN_ONLY), rle=dict(level=6, strategy=zlib.Z_RLE),
             syncflush=dict(level=6, flush_every=70000))


@pytest.mark.parametrize("mode", sorted(MODES))
def test_emulated_inflate_block_types(mode):
    gz = corpus.gz_member(corpus.fastq(6000, fixed=150), **MODES[mode])
    n, _ = _check(gz, 1000, 64)
    assert n >= 2


@pytest.mark.parametrize("T,subw", [(32, 31), (64, 31), (256, 31), (512, 31), (1024, 31), (128, 27), (64, 63)])  # 512 / 1024: the shipped CTA sizes
def test_emulated_inflate_geometries(T, subw):
    gz = corpus.gz_member(corpus.fastq(12000, fixed=150), 6)
    n, bits = _check(gz, 2000, T, subw)
    assert n >= 4 and len(bits) >= 2  # unaligned checkpoints are the norm


def test_emulated_inflate_other_writers():
    _check(corpus.gz_system(corpus.fastq(8000), 6), 2000, 128)
    _check(corpus.gz_parallel(corpus.fastq(20000, fixed=150), 6, segment=1 << 20), 1000, 64)
    _check(corpus.gz_member(corpus.fastq(200, lognormal=(10000, 0.5), seed=3), 6), 20, 128, lift=True)


def test_emulated_inflate_incompressible_and_runs():
    rng = np.random.default_rng(1)
    raw = rng.integers(0, 256, 300000, dtype=np.uint8).tobytes()          # stored blocks with real payload
    runs = (b"@r\n" + b"A" * 5000 + b"\n+\n" + b"?" * 5000 + b"\n") * 40     # dist-1 overlapping runs of length 258
    for data, T in ((raw, 64), (runs, 64), (raw[:70000] + runs + raw[:5000], 128)):
        gz = corpus.gz_member(data, 6)
        ox = O.OracleIndex.build(gz, 100000, True)
        p = ox.point(0)
        ref = np.frombuffer(data, np.uint8)
        st, got, nl, mb, _ = E.inflate_chunk(gz, p["input"], p["bits"], ox.inputs()[1], p["window"], len(data), T)
        assert st == 0 and np.array_equal(got, ref)
        assert mb == (0 if (ref == 0).any() else 1)


def test_emulated_inflate_stops_when_output_is_full():
    """Core.cs:187: the loop ends once `len` bytes are out, wherever that is inside a block."""
    fq = corpus.fastq(3000, fixed=150)
    gz = corpus.gz_member(fq, 6)
    ox = O.OracleIndex.build(gz, 100000)
    p = ox.point(0)
    for want in (1, 17, 4097, 100001, len(fq) - 3):
        st, got, _, _, _ = E.inflate_chunk(gz, p["input"], p["bits"], ox.inputs()[1], p["window"], want, 64)
        assert st == 0 and got.tobytes() == fq[:want]


def test_emulated_inflate_rejects_corrupt_input():
    fq = corpus.fastq(3000, fixed=150)
    gz = corpus.gz_member(fq, 6)
    ox = O.OracleIndex.build(gz, 100000)
    p = ox.point(0)
    bad = gz.copy()
    bad[p["input"]] |= 0x06  # block type 3: invalid
    st, _, _, _, _ = E.inflate_chunk(bad, p["input"], p["bits"], ox.inputs()[1], p["window"], len(fq), 64)
    assert st == -3
    # truncated input: the chunk may not read past in_limit (Core.cs:174)
    st, got, _, _, _ = E.inflate_chunk(gz, p["input"], p["bits"], p["input"] + 2000, p["window"], len(fq), 64)
    assert st == -3
    # a stream that ends (final block) before `len` bytes: Z_STREAM_END, short count (Core.cs:185,191)
    st, got, _, _, _ = E.inflate_chunk(gz, p["input"], p["bits"], ox.inputs()[1], p["window"], len(fq) + 500, 64)
    assert st == 0 and got.tobytes() == fq


# ---- block scanner (GPU-assisted CreateIndex, first slice): logic on the CPU ----------------------

@pytest.mark.parametrize("mode", ["dynamic6", "dynamic1", "dynamic9", "huffman", "rle", "syncflush"])
def test_emulated_probe_finds_every_dynamic_block_start_and_nothing_else(mode):
    """The speculative header probe (one thread per bit position on the GPU): true at every dynamic
    block's first bit as zlib's Z_BLOCK pass reports them, false at random other positions."""
    gz = corpus.gz_member(corpus.fastq(9000, fixed=150), **MODES[mode])
    bits, outs, kinds, end, tot = O.block_stops(gz)
    dyn = [int(b) for b, k in zip(bits, kinds) if k == 2]
    assert len(dyn) >= 3
    assert all(E.probe_dynamic_header(gz, b) for b in dyn)
    starts = set(int(b) for b in bits)
    rng = np.random.default_rng(3)
    others = [int(p) for p in rng.integers(int(bits[0]), int(bits[-1]), 4000) if int(p) not in starts]
    assert sum(E.probe_dynamic_header(gz, p) for p in others) == 0


@pytest.mark.parametrize("mode", sorted(MODES))
def test_emulated_block_walk_equals_zlib_block_stops(mode):
    """Walking the blocks with the Huffman passes alone (no history, no output) from the first block
    gives zlib's stops: every block's first bit and output offset, the stream length, the final block."""
    gz = corpus.gz_member(corpus.fastq(6000, fixed=150), **MODES[mode])
    bits, outs, kinds, end, tot = O.block_stops(gz)
    for T in (32, 128):
        st, first, land, ob, recs = E.scan_segment(gz, int(bits[0]), gz.size * 8, False, T=T, rec_cap=1 << 15)
        assert st == 1 and first == int(bits[0]) and ob == tot
        assert np.array_equal(recs[:, 0], bits) and np.array_equal(recs[:, 1], outs)
        assert (land + 7) // 8 * 8 + 64 == end   # the gzip trailer (CRC-32, ISIZE) follows the final block


def test_emulated_block_search_lands_on_a_true_block_start():
    """SEARCH + WALK from arbitrary bit positions: the first block found is the next true block start
    (dynamic blocks), and a walk bounded by end_bit lands on the first block start at or past it."""
    gz = corpus.gz_member(corpus.fastq(20000), 6)
    bits, outs, kinds, end, tot = O.block_stops(gz)
    rng = np.random.default_rng(8)
    for _ in range(6):
        a = int(rng.integers(int(bits[0]) + 1, int(bits[-1])))
        b = int(min(a + rng.integers(100_000, 3_000_000), gz.size * 8))
        st, first, land, ob, recs = E.scan_segment(gz, a, b, True)
        i = int(np.searchsorted(bits, a))
        if bits[i] >= b:
            assert first == 2 ** 64 - 1 and len(recs) == 0     # no block starts inside the segment
            continue
        assert first == int(bits[i]) and st in (0, 1)
        j = int(np.searchsorted(bits, b))
        want_land = int(bits[j]) if j < len(bits) else None
        if want_land is not None:
            assert land == want_land and np.array_equal(recs[:, 0], bits[i:j])
            assert ob == int(outs[j] - outs[i]) and np.array_equal(recs[:, 1], outs[i:j] - outs[i])


# ----------------------------------------------------------------------------------------------
# GPU CreateIndex, DECODE step: one Huffman decode resolved against two position-coded dictionaries
# ----------------------------------------------------------------------------------------------

def _coded_windows():
    i = np.arange(32768, dtype=np.uint32)
    return (i & 0xFF).astype(np.uint8), (((i >> 8) + 1 + (i & 0xFF)) & 0xFF).astype(np.uint8)


@pytest.mark.parametrize("kind,T", [("dynamic", 64), ("dynamic", 512), ("syncflush", 128), ("fixed", 64), ("stored", 64)])
def test_dual_inflate_with_position_coded_windows_reconstructs_the_stream(kind, T):
    """createindex.cu's method on the emulated kernel: segments of consecutive blocks are inflated with
    two dictionaries that encode their own positions (inflate_chunk<DUAL>: one decode, two resolves); equal
    bytes are final, differing bytes name the window position they copy; chaining the windows through the
    segments gives back zlib's output exactly.  Also: the dual call equals two single calls."""
    kw = dict(dynamic=dict(level=6), syncflush=dict(level=6, flush_every=30000), fixed=dict(level=6, strategy=zlib.Z_FIXED),
              stored=dict(level=0))[kind]
    data = corpus.fastq(2500, fixed=150, seed=4)
    gz = corpus.gz_member(data, **kw)
    bits, outs, kinds, end, tot = O.block_stops(gz)
    ref = np.frombuffer(zlib.decompress(gz.tobytes(), 47), np.uint8)
    wa, wb = _coded_windows()
    step = max(1, len(bits) // 5)
    firsts = list(range(0, len(bits), step))
    got = np.zeros(tot, np.uint8)
    win = np.zeros(32768, np.uint8)
    dependent = 0
    for si, b0 in enumerate(firsts):
        b1 = firsts[si + 1] if si + 1 < len(firsts) else len(bits)
        sb = int(bits[b0])
        o0 = int(outs[b0])
        o1 = int(outs[b1]) if b1 < len(bits) else int(tot)
        in_byte, nb = (sb + 7) // 8, (8 - sb % 8) % 8
        st, a, b = E.inflate_chunk_dual(gz, in_byte, nb, gz.size, wa, wb, o1 - o0, T, 27)
        assert st == 0 and a.size == o1 - o0
        if si == 1:   # the dual call is two single calls
            ra = E.inflate_chunk(gz, in_byte, nb, gz.size, wa, o1 - o0, T, 27)
            rb = E.inflate_chunk(gz, in_byte, nb, gz.size, wb, o1 - o0, T, 27)
            assert np.array_equal(ra[1], a) and np.array_equal(rb[1], b)
        a32, b32 = a.astype(np.uint32), b.astype(np.uint32)
        dep = a32 != b32
        pos = a32 + 256 * ((b32 - a32 - 1) & 0xFF)
        assert (pos[dep] < 32768).all()
        if si == 0:
            assert not dep.any()
        seg = np.where(dep, win[np.minimum(pos, 32767)], a)
        got[o0:o1] = seg
        win = np.concatenate([win, seg])[-32768:]
        dependent += int(dep.sum())
    assert np.array_equal(got, ref)
    if kind == "dynamic":
        assert dependent > 0


def test_window_chain_as_a_parallel_scan_of_maps():
    """createindex.cu CHAIN: the window behind a segment as a function of the window in front of it (per byte:
    a final value, or 0x8000 | position in the window in front), composed over all segments by a Hillis-Steele
    scan — the same windows as walking the segments in order.  Includes segments shorter than a window (the
    old window slides) and the first segment (nothing in front of it)."""
    data = corpus.fastq(2500, fixed=150, seed=9)
    gz = corpus.gz_member(data, 6, flush_every=9000)       # many short blocks: segments shorter than 32 KB occur
    bits, outs, kinds, end, tot = O.block_stops(gz)
    ref = np.frombuffer(zlib.decompress(gz.tobytes(), 47), np.uint8)
    wa, wb = _coded_windows()
    rng = np.random.default_rng(2)
    firsts = [0]
    while firsts[-1] + 1 < len(bits):
        firsts.append(min(len(bits) - 1, firsts[-1] + int(rng.integers(1, 9))))
    firsts = sorted(set(firsts))
    maps, lens = [], []
    for si, b0 in enumerate(firsts):
        b1 = firsts[si + 1] if si + 1 < len(firsts) else len(bits)
        sb, o0 = int(bits[b0]), int(outs[b0])
        o1 = int(outs[b1]) if b1 < len(bits) else int(tot)
        st, a, b = E.inflate_chunk_dual(gz, (sb + 7) // 8, (8 - sb % 8) % 8, gz.size, wa, wb, o1 - o0, 64, 27)
        assert st == 0
        n = min(o1 - o0, 32768)
        a32, b32 = a[a.size - n:].astype(np.uint32), b[b.size - n:].astype(np.uint32)
        tail = np.where(a32 == b32, a32, 0x8000 | (a32 + 256 * ((b32 - a32 - 1) & 0xFF)))
        slide = 0x8000 | (np.arange(32768 - n, dtype=np.uint32) + n)      # window byte j <- old window byte j + n
        maps.append(np.concatenate([slide, tail]).astype(np.uint32))
        lens.append(o1 - o0)
    assert min(lens) < 32768 < max(lens)
    # in order, as a serial walk would do it
    serial, w = [], np.zeros(32768, np.uint32)
    for m in maps:
        w = np.where(m & 0x8000, w[m & 0x7FFF], m & 0xFF)
        serial.append(w)
    # Hillis-Steele: P[s] = P[s] o P[s-d], d = 1, 2, 4, ...
    P, d = [m.copy() for m in maps], 1
    while d < len(P):
        Q = [p.copy() for p in P]
        for s in range(d, len(P)):
            dep = (P[s] & 0x8000) != 0
            Q[s] = np.where(dep, P[s - d][P[s] & 0x7FFF], P[s])
        P, d = Q, 2 * d
    ends = np.cumsum(lens)
    for s in range(len(P)):
        got = np.where(P[s] & 0x8000, 0, P[s] & 0xFF)                # what still points in front of the stream: zero
        assert np.array_equal(got, serial[s]), s
        e = int(ends[s])
        want = np.concatenate([np.zeros(max(0, 32768 - e), np.uint8), ref[max(0, e - 32768): e]])
        assert np.array_equal(got.astype(np.uint8), want), s


@pytest.mark.parametrize("kind,T,subw", [("dynamic", 64, 27), ("dynamic", 512, 27), ("dynamic1", 128, 31), ("syncflush", 128, 27),
                                         ("fixed", 64, 31), ("stored", 64, 31), ("mixed", 64, 27)])
def test_pull_mode_reuse_of_staged_bytes_equals_plain_decode(kind, T, subw):
    """inflate_chunk<PULL> (what PP_JOB_ZEROCOPY jobs run): a window that starts behind RESOLVE's scratch inside
    the bytes staged for the window before it moves them down in shared memory and fetches only the rest —
    the same bytes, end bit and status as the plain path, with stored blocks in between (the bit cursor
    jumps: nothing to re-use) and when the buffer ends exactly where the data ends."""
    kw = dict(dynamic=dict(level=6), dynamic1=dict(level=1), syncflush=dict(level=6, flush_every=30000),
              fixed=dict(level=6, strategy=zlib.Z_FIXED), stored=dict(level=0), mixed=dict(level=6))[kind]
    data = corpus.fastq(4000, fixed=150, seed=6)
    if kind == "mixed":
        data = data[:200000] + bytes(np.random.default_rng(1).integers(0, 256, 150000, dtype=np.uint8)) + data[200000:400000]
    gz = corpus.gz_member(data, **kw)
    ox = O.OracleIndex.build(gz, 1000 if kind != "mixed" else 400, True)
    outs, ins = ox.outputs(), ox.inputs()
    for k in range(ox.count - 1):
        p = ox.point(k)
        ref = O.extract(gz, ox, k)
        a = E.inflate_chunk(gz, p["input"], p["bits"], ins[k + 1], p["window"], outs[k + 1] - outs[k], T, subw)
        for exact in (False, True):
            b = E.inflate_chunk_pull(gz, p["input"], p["bits"], ins[k + 1], p["window"], outs[k + 1] - outs[k], T, subw, exact)
            assert b[0] == a[0] == 0 and np.array_equal(b[1], ref) and b[2:] == a[2:], (k, exact)

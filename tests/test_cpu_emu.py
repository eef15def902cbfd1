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

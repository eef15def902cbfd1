"""CPU fuzz of the inflate kernel's logic (host emulation of inflate_core.cuh) against zlib:
random data, random deflate parameters and flush patterns, random stop points, and random bit flips
in the compressed stream — wherever zlib (through the oracle's restatement of
Core.ExtractDeflateIndex) produces bytes the kernel logic must produce the same bytes, and
wherever zlib reports a data error the kernel logic must report one too (-3)."""
import zlib

import numpy as np
import pytest

import corpus
import emu_lib as E
import oracle_lib as O


def _emu_chunk(gz, ox, k, T=64, out_len=None):
    p = ox.point(k)
    outs, ins = ox.outputs(), ox.inputs()
    n = outs[k + 1] - outs[k] if out_len is None else out_len
    return E.inflate_chunk(gz, p["input"], p["bits"], ins[k + 1], p["window"], n, T)


def _random_payload(rng, n):
    kind = rng.integers(0, 4)
    if kind == 0:   # text-like with long repeats
        words = [bytes(rng.integers(97, 123, rng.integers(1, 9), dtype=np.uint8)) for _ in range(50)]
        return b" ".join(words[i] for i in rng.integers(0, 50, n // 5))[:n]
    if kind == 1:   # low-entropy runs
        return bytes(np.repeat(rng.integers(0, 4, n // 8 + 1, dtype=np.uint8), rng.integers(1, 40, n // 8 + 1))[:n])
    if kind == 2:   # incompressible
        return rng.integers(0, 256, n, dtype=np.uint8).tobytes()
    return corpus.fastq(max(n // 400, 10), seed=int(rng.integers(0, 1000)))[:n]


@pytest.mark.parametrize("seed", range(6))
def test_fuzz_random_streams(seed):
    rng = np.random.default_rng(seed)
    data = b"".join(_random_payload(rng, int(rng.integers(20000, 150000))) for _ in range(3))
    level = int(rng.integers(1, 10))
    strategy = int(rng.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED]))
    gz = corpus.gz_member(data, level, strategy, flush_every=int(rng.choice([0, 7001, 65536])),
                          mem_level=int(rng.integers(1, 10)))
    # the index cuts at deflate block ends after enough '@' bytes; chunksize 9 makes every block end a candidate
    ox = O.OracleIndex.build(gz, 9, True)
    ref_all = np.frombuffer(data, np.uint8)
    outs = ox.outputs()
    for k in range(ox.count - 1):
        T = int(rng.choice([32, 64, 128, 512]))
        st, got, nl, mb, eb = _emu_chunk(gz, ox, k, T=T)
        ref = ref_all[outs[k]: outs[k + 1]]
        assert st == 0 and np.array_equal(got, ref), (seed, k)
        assert nl == int((ref == 10).sum())
        # the pull-mode variant (staged bytes re-used between windows): the same in every respect
        p = ox.point(k)
        pst, pgot, pnl, pmb, peb = E.inflate_chunk_pull(gz, p["input"], p["bits"], ox.inputs()[k + 1], p["window"],
                                                        outs[k + 1] - outs[k], T, 31, exact_extent=bool(k & 1))
        assert pst == 0 and np.array_equal(pgot, ref) and (pnl, pmb, peb) == (nl, mb, eb), (seed, k)
    # stop points inside blocks (Core.cs:187)
    k = int(rng.integers(0, ox.count - 1))
    full = outs[k + 1] - outs[k]
    for want in sorted(set(int(x) for x in rng.integers(1, max(full, 2), 4))):
        st, got, _, _, _ = _emu_chunk(gz, ox, k, out_len=want)
        assert st == 0 and np.array_equal(got, ref_all[outs[k]: outs[k] + want]), (seed, k, want)


@pytest.mark.parametrize("seed", range(6))
def test_fuzz_bit_flips(seed):
    rng = np.random.default_rng(100 + seed)
    fq = corpus.fastq(4000, fixed=150, seed=seed)
    # seeds 4 and 5: fixed-Huffman blocks (invalid symbols 286/287 and distance codes 30/31 exist
    # there) and sync-flushed streams (stored blocks whose LEN/NLEN can be hit)
    kw = dict(strategy=zlib.Z_FIXED) if seed == 4 else dict(flush_every=30000) if seed == 5 else {}
    gz = corpus.gz_member(fq, int(rng.choice([1, 6, 9])), **kw)
    ox = O.OracleIndex.build(gz, 700)
    ins = ox.inputs()
    agree_err = agree_ok = 0
    for _ in range(40):
        k = int(rng.integers(0, ox.count - 1))
        bad = gz.copy()
        for _ in range(int(rng.integers(1, 4))):
            pos = int(rng.integers(ins[k] + 1, ins[k + 1] - 1))
            bad[pos] ^= np.uint8(1 << int(rng.integers(0, 8)))
        try:
            ref = O.extract(bad, ox, k)
        except RuntimeError:
            ref = None
        st, got, _, _, _ = _emu_chunk(bad, ox, k)
        if ref is None:
            assert st == -3, (seed, k, "zlib reports a data error, the kernel logic does not")
            agree_err += 1
        else:
            assert st == 0 and got.size == ref.size and np.array_equal(got, ref), (seed, k)
            agree_ok += 1
    assert agree_err + agree_ok == 40 and agree_err > 0


@pytest.mark.parametrize("seed", range(4))
def test_fuzz_bit_flips_in_block_headers(seed):
    """Flips aimed at the dynamic block header every chunk starts with (HLIT/HDIST/HCLEN, the
    code-length code, the run-length coded lengths): over-subscribed or incomplete codes, repeats
    with no previous length, repeats past the end, a missing end-of-block code.  zlib's verdict
    (inflate_table / 'invalid bit length repeat' / 'missing end-of-block') is the reference's."""
    rng = np.random.default_rng(900 + seed)
    fq = corpus.fastq(3000, fixed=int(rng.choice([50, 150])), seed=seed)
    gz = corpus.gz_member(fq, int(rng.choice([1, 6, 9])), strategy=zlib.Z_DEFAULT_STRATEGY if seed < 3 else zlib.Z_RLE)
    ox = O.OracleIndex.build(gz, 300)
    ins = ox.inputs()
    agree_err = agree_ok = 0
    for _ in range(80):
        k = int(rng.integers(0, ox.count - 1))
        bad = gz.copy()
        # a chunk starts at a block boundary, so its first ~90 bytes are a block header
        span = min(90, ins[k + 1] - ins[k] - 2)
        if span < 8:
            continue
        for _ in range(int(rng.integers(1, 3))):
            pos = int(ins[k] + rng.integers(0, span))
            bad[pos] ^= np.uint8(1 << int(rng.integers(0, 8)))
        try:
            ref = O.extract(bad, ox, k)
        except RuntimeError:
            ref = None
        st, got, _, _, _ = _emu_chunk(bad, ox, k, T=int(rng.choice([32, 64, 128])))
        if ref is None:
            assert st == -3, (seed, k, "zlib reports a data error, the kernel logic does not")
            agree_err += 1
        else:
            assert st == 0 and got.size == ref.size and np.array_equal(got, ref), (seed, k)
            agree_ok += 1
    assert agree_err > 10

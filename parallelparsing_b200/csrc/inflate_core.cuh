// Checkpoint inflate — the DEFLATE (RFC 1951) decoder that replaces zlib's
// inflate() on the Decompress(checkpoint) path (Core.ExtractDeflateIndex,
// Decompressor/Core.cs:133-192; zlib reached through Interop/PlatformInterop.cs:9-34).
//
// Execution model: ONE WARP decodes ONE index chunk.  All 32 lanes run the
// serial Huffman decode redundantly on identical registers (the table lookups
// are shared-memory broadcasts), so every lane knows every (length, distance)
// token without any queue, shuffle or barrier; the lanes then split only where
// the work is data parallel:
//   * LZ77 copies: lane j moves byte j of the match (coalesced byte loads from
//     the history, coalesced byte stores to the output);
//   * Huffman table fills: lane j writes replica j of a code;
//   * the 32 KB checkpoint window is copied into place with 16-byte vectors.
// History is addressed directly in global memory: the chunk's output slot is
// laid out as [window (>= 32 KB)][output], so a back-reference is out[pos-dist]
// with no wrap test and stays in L1/L2.
//
// The same source compiles in two modes:
//   * device (default): used by inflate.cu, the only mode shipped in libppb200.so;
//   * PP_HOST_EMU: lanes are run one after another by a plain loop.  Compiled
//     ONLY by tests/emu/ to check the decoder logic against zlib on machines
//     without a GPU.  It is test scaffolding, not a fallback: nothing in the
//     product links it.
// Lane-parallel sections never read a byte written in the same section, which
// is what makes the sequential emulation equivalent to lockstep execution.
#pragma once
#include <stdint.h>

#ifdef PP_HOST_EMU
#include <string.h>
#define PP_DEV static inline
#define PP_LANES_BEGIN for (int lane = 0; lane < 32; ++lane) {
#define PP_LANES_END }
#define PP_LANE0_BEGIN {
#define PP_LANE0_END }
#define PP_LV(T, name) T name[32]
#define PP_L(name) name[lane]
#define PP_L0(name) name[0]
#define PP_SHARED
#else
#define PP_DEV __device__ __forceinline__
#define PP_LANES_BEGIN { const int lane = (int)(threadIdx.x & 31u);
#define PP_LANES_END } __syncwarp();
#define PP_LANE0_BEGIN if ((threadIdx.x & 31u) == 0) {
#define PP_LANE0_END } __syncwarp();
#define PP_LV(T, name) T name
#define PP_L(name) name
#define PP_L0(name) name
#define PP_SHARED
#endif

namespace ppinf {

// ---- geometry -------------------------------------------------------------
constexpr int kRootL = 9;              // primary bits, literal/length table
constexpr int kRootD = 9;              // primary bits, distance table
constexpr int kLitCap = 852 + 4;       // zlib's proven bound for (286 syms, root 9, max 15)
constexpr int kDistCap = 512 + 288;    // 30 syms: every 2^k sub-table holds >= k+1 symbols
constexpr int kTileBytes = 2048;       // one TMA bulk copy
constexpr int kTileWords = kTileBytes / 4;
constexpr int kStages = 4;             // ring depth
constexpr int kRingWords = kTileWords * kStages;

// ---- table entry ------------------------------------------------------------
// [4:0] bits to consume (code + extra)   [7:5] kind   [11:8] code length
// [15:12] sub-table index bits (kind SUB) [31:16] literal / base value / sub-table start
enum : uint32_t { K_LIT = 0, K_BASE = 1, K_SUB = 2, K_EOB = 3, K_BAD = 4 };
PP_DEV uint32_t mk_entry(uint32_t kind, uint32_t tot, uint32_t cl, uint32_t val)
{
    return tot | (kind << 5) | (cl << 8) | (val << 16);
}
PP_DEV uint32_t e_kind(uint32_t e) { return (e >> 5) & 7u; }
PP_DEV uint32_t e_tot(uint32_t e) { return e & 31u; }
PP_DEV uint32_t e_cl(uint32_t e) { return (e >> 8) & 15u; }
PP_DEV uint32_t e_sub(uint32_t e) { return (e >> 12) & 15u; }
PP_DEV uint32_t e_val(uint32_t e) { return e >> 16; }

// ---- per-chunk descriptor / result (shared with the host runtime) -----------
struct ChunkDesc {
    uint64_t in_bit;      // first bit of the chunk, relative to the compressed buffer (8*Input-Bits)
    uint64_t in_limit;    // bytes of compressed buffer the chunk may touch (relative; to.Input-ish)
    uint64_t slot_off;    // byte offset of the chunk's slot in the slots buffer (128 B aligned)
    uint64_t lead_src;    // byte offset of the chunk's window bytes in the lead staging buffer
    uint32_t lead_len;    // history bytes placed before the output (>= 32768, multiple of 16)
    uint32_t out_len;     // to.Output - from.Output
    uint32_t prefix_len;  // |from.offset| (used by the parse stage)
    uint32_t prefix_nl;   // '\n' count inside from.offset (host counted)
};
struct ChunkResult {
    int32_t status;     // 0 or negative ZResult
    uint32_t produced;  // bytes written (Core.cs:191)
    uint32_t newlines;  // '\n' bytes among them (by-product for the parse stage)
    uint32_t min_byte;  // smallest byte written (0 => a NUL is present => exact parser)
    uint64_t end_bit;   // bit position after the last consumed bit
};

struct alignas(128) Smem {
    uint32_t ring[kRingWords];  // first: TMA destinations must be 16 B aligned
    uint32_t lit[kLitCap];
    uint32_t dist[kDistCap];
    uint16_t count[16];
    uint16_t next[16];
    uint8_t lens[320];
#ifndef PP_HOST_EMU
    unsigned long long bar[kStages];
#endif
};

// RFC 1951 3.2.5 length / distance bases and extra-bit counts
#ifdef PP_HOST_EMU
#define PP_CONST static const
#else
#define PP_CONST __device__ const
#endif
PP_CONST uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
PP_CONST uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
PP_CONST uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
PP_CONST uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
PP_CONST uint8_t kClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

PP_DEV uint32_t bitrev(uint32_t v, int n)
{
#ifdef PP_HOST_EMU
    uint32_t r = 0;
    for (int i = 0; i < n; i++) r |= ((v >> i) & 1u) << (n - 1 - i);
    return r;
#else
    return __brev(v) >> (32 - n);
#endif
}

// ---- compressed-input reader ------------------------------------------------
// A 64-bit bit buffer fed 32 bits at a time from a shared-memory ring that TMA
// bulk copies (cp.async.bulk + mbarrier complete_tx) keep kStages tiles ahead.
struct Reader {
    const uint8_t *comp;   // compressed buffer (global, 16 B aligned)
    uint64_t comp_tiles;   // tiles available in the buffer
    uint64_t bitbuf;
    int bitcnt;            // valid bits in bitbuf
    uint64_t wnext;        // word index (from comp) of the next word to append
    uint32_t nw;           // preloaded word at wnext
    uint32_t rpos;         // ring position of wnext
    uint32_t s_cur;        // sequence number of the tile holding wnext
    uint32_t s_issued;     // tiles issued so far
    int64_t tile_bias;     // tile index = tile_bias + sequence number
    int exhausted;         // reader ran past the buffer
};

#ifndef PP_HOST_EMU
PP_DEV uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
PP_DEV void mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
PP_DEV void mbar_expect_tx(unsigned long long *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a transfer that never lands (bad pointer, driver fault) must not hang the GPU.
PP_DEV bool mbar_wait(unsigned long long *bar, uint32_t parity)
{
    const long long t0 = clock64();
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return true;
        if (clock64() - t0 > 4000000000LL) return false;  // ~2 s
    }
}
PP_DEV void tma_load_tile(void *dst_smem, const void *src_gmem, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"((uint32_t)kTileBytes), "r"(smem_u32(bar))
                 : "memory");
}
#endif

PP_DEV void rd_issue(Reader &r, Smem &sm)
{
    // issue the tile with sequence number s_issued (if the buffer still has it)
    const int64_t tile = r.tile_bias + (int64_t)r.s_issued;
    const uint32_t stage = r.s_issued % kStages;
    if (tile >= 0 && (uint64_t)tile < r.comp_tiles) {
#ifdef PP_HOST_EMU
        memcpy(&sm.ring[stage * kTileWords], r.comp + (uint64_t)tile * kTileBytes, kTileBytes);
#else
        __syncwarp();  // every lane is done reading the stage being overwritten
        if ((threadIdx.x & 31u) == 0) {
            mbar_expect_tx(&sm.bar[stage], kTileBytes);
            tma_load_tile(&sm.ring[stage * kTileWords], r.comp + (uint64_t)tile * kTileBytes, &sm.bar[stage]);
        }
#endif
    } else {
#ifndef PP_HOST_EMU
        // nothing to load: complete the phase by hand so waiters do not hang
        if ((threadIdx.x & 31u) == 0)
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&sm.bar[stage])) : "memory");
#endif
    }
    r.s_issued++;
}

PP_DEV void rd_wait(Reader &r, Smem &sm, uint32_t seq)
{
#ifndef PP_HOST_EMU
    if (!mbar_wait(&sm.bar[seq % kStages], (seq / kStages) & 1u)) r.exhausted = 2;
#else
    (void)sm; (void)seq;
#endif
    const int64_t tile = r.tile_bias + (int64_t)seq;
    if (tile < 0 || (uint64_t)tile >= r.comp_tiles) r.exhausted |= 1;
}

// Position the reader at absolute bit `bit` of the compressed buffer.
PP_DEV void rd_seek(Reader &r, Smem &sm, uint64_t bit, bool first)
{
    if (!first) {
        // drain tiles still in flight so the ring can be re-targeted
        for (uint32_t s = r.s_cur + 1; s < r.s_issued; s++) rd_wait(r, sm, s);
    }
    const uint64_t word = bit >> 5;
    const uint64_t tile = word / kTileWords;
    r.tile_bias = (int64_t)tile - (int64_t)r.s_issued;
    r.s_cur = r.s_issued;
    r.exhausted &= 2;  // a timed-out transfer stays fatal
    for (int i = 0; i < kStages; i++) rd_issue(r, sm);
    rd_wait(r, sm, r.s_cur);
    r.wnext = word;
    r.rpos = (r.s_cur % kStages) * kTileWords + (uint32_t)(word % kTileWords);
    r.nw = sm.ring[r.rpos];
    // load the first word and drop the bits in front of `bit`
    r.bitbuf = (uint64_t)r.nw >> (bit & 31ull);
    r.bitcnt = 32 - (int)(bit & 31ull);
    // (rd_advance is defined below; the body is repeated here to keep this self-contained)
    r.wnext++;
    r.rpos++;
    if ((r.rpos & (kTileWords - 1)) == 0) {
        if (r.rpos == kRingWords) r.rpos = 0;
        r.s_cur++;
        rd_wait(r, sm, r.s_cur);
        rd_issue(r, sm);
    }
    r.nw = sm.ring[r.rpos];
}

// advance the preloaded word by one (called after nw has been appended)
PP_DEV void rd_advance(Reader &r, Smem &sm)
{
    r.wnext++;
    r.rpos++;
    if ((r.rpos & (kTileWords - 1)) == 0) {
        // entering the next tile: its stage must have landed; the stage we leave is free again
        if (r.rpos == kRingWords) r.rpos = 0;
        r.s_cur++;
        rd_wait(r, sm, r.s_cur);
        rd_issue(r, sm);
    }
    r.nw = sm.ring[r.rpos];
}

// guarantee at least 33 valid bits
PP_DEV void rd_refill(Reader &r, Smem &sm)
{
    if (r.bitcnt <= 32) {
        r.bitbuf |= (uint64_t)r.nw << r.bitcnt;
        r.bitcnt += 32;
        rd_advance(r, sm);
    }
}
PP_DEV uint32_t rd_lo(const Reader &r) { return (uint32_t)r.bitbuf; }
PP_DEV void rd_consume(Reader &r, uint32_t n)
{
    r.bitbuf >>= n;
    r.bitcnt -= (int)n;
}
PP_DEV uint32_t rd_bits(Reader &r, Smem &sm, uint32_t n)  // n <= 16
{
    rd_refill(r, sm);
    const uint32_t v = rd_lo(r) & ((1u << n) - 1u);
    rd_consume(r, n);
    return v;
}
PP_DEV uint64_t rd_bitpos(const Reader &r) { return r.wnext * 32ull - (uint64_t)r.bitcnt; }

// ---- Huffman table construction ---------------------------------------------
// Canonical codes (RFC 1951 3.2.2) into a two-level lookup table: `root` primary
// bits, sub-tables sized by the longest code under each primary prefix.
// Validity rules follow zlib's inflate_table: over-subscribed sets are rejected,
// incomplete sets are rejected unless the set is a single 1-bit code (or, for
// distances, empty).  Returns 0 or -3 (Z_DATA_ERROR).
// is_dist selects the symbol -> (base, extra) mapping.
PP_DEV int build_table(Smem &sm, uint32_t *tbl, int root, int cap, int nsym, int lens_off, bool is_dist)
{
    const uint8_t *lens = sm.lens + lens_off;
    // 1. histogram of code lengths
    PP_LANE0_BEGIN
    for (int i = 0; i < 16; i++) sm.count[i] = 0;
    for (int s = 0; s < nsym; s++) sm.count[lens[s]]++;
    PP_LANE0_END
    int maxlen = 0;
    for (int l = 15; l >= 1; l--)
        if (sm.count[l]) { maxlen = l; break; }
    // 2. validity + first code of each length
    int left = 1;
    for (int l = 1; l <= 15; l++) {
        left <<= 1;
        left -= (int)sm.count[l];
        if (left < 0) return -3;  // over-subscribed
    }
    if (left > 0 && maxlen != 1 && !(is_dist && maxlen == 0)) return -3;  // incomplete set
    if (left > 0 && maxlen == 1 && sm.count[1] != 1) return -3;
    PP_LANE0_BEGIN
    {
        uint32_t code = 0;
        sm.next[0] = 0;
        for (int l = 1; l <= 15; l++) {
            code = (code + sm.count[l - 1] * (l > 1 ? 1u : 0u)) << 1;
            sm.next[l] = (uint16_t)code;
        }
    }
    PP_LANE0_END
    // 3. primary table: start from "invalid code" everywhere (incomplete sets leave holes)
    const int nprim = 1 << root;
    PP_LANES_BEGIN
    for (int i = lane; i < nprim; i += 32) tbl[i] = mk_entry(K_BAD, 1, 1, 0);
    PP_LANES_END
    // 4. sub-table geometry for prefixes that own codes longer than `root`
    int used = nprim;
    if (maxlen > root) {
        // first[l]: first canonical code of length l (before assignment consumed sm.next)
        // prefix p (root bits, MSB-first code order) owns length-l codes in
        // [p << (l-root), (p+1) << (l-root)); it needs a sub-table when that range meets
        // [first[l], first[l]+count[l]) for some l > root.  Prefixes are visited in code
        // order so sub-tables are laid out in canonical order.
        int pmin = nprim;
        for (int l = root + 1; l <= maxlen; l++)
            if (sm.count[l]) { pmin = (int)(sm.next[l] >> (l - root)); break; }
        for (int p = pmin; p < nprim; p++) {
            int sub = 0;
            for (int l = maxlen; l > root; l--) {
                const uint32_t lo = (uint32_t)p << (l - root), hi = ((uint32_t)p + 1u) << (l - root);
                const uint32_t f = sm.next[l], e = f + sm.count[l];
                if (sm.count[l] && lo < e && hi > f) { sub = l - root; break; }
            }
            if (sub) {
                if (used + (1 << sub) > cap) return -3;  // cannot happen for valid sets (see kLitCap/kDistCap)
                const uint32_t idx = bitrev((uint32_t)p, root);
                PP_LANE0_BEGIN
                tbl[idx] = mk_entry(K_SUB, (uint32_t)root, (uint32_t)root, (uint32_t)used) | ((uint32_t)sub << 12);
                PP_LANE0_END
                // pre-fill the sub-table with "invalid"
                PP_LANES_BEGIN
                for (int i = lane; i < (1 << sub); i += 32) tbl[used + i] = mk_entry(K_BAD, 1, 1, 0);
                PP_LANES_END
                used += 1 << sub;
            }
        }
    }
    // 5. assign codes in symbol order and replicate entries
    for (int s = 0; s < nsym; s++) {
        const int l = lens[s];
        if (l == 0) continue;
        const uint32_t code = sm.next[l];
        PP_LANE0_BEGIN
        sm.next[l] = (uint16_t)(code + 1);
        PP_LANE0_END
        uint32_t kind, tot, val;
        if (!is_dist) {
            if (s < 256) { kind = K_LIT; tot = (uint32_t)l; val = (uint32_t)s; }
            else if (s == 256) { kind = K_EOB; tot = (uint32_t)l; val = 0; }
            else if (s < 286) { kind = K_BASE; tot = (uint32_t)l + kLenExtra[s - 257]; val = kLenBase[s - 257]; }
            else { kind = K_BAD; tot = (uint32_t)l; val = 0; }
        } else {
            if (s < 30) { kind = K_BASE; tot = (uint32_t)l + kDistExtra[s]; val = kDistBase[s]; }
            else { kind = K_BAD; tot = (uint32_t)l; val = 0; }
        }
        const uint32_t ent = mk_entry(kind, tot, (uint32_t)l, val);
        if (l <= root) {
            const uint32_t base = bitrev(code, l);
            const int reps = 1 << (root - l);
            PP_LANES_BEGIN
            for (int j = lane; j < reps; j += 32) tbl[base | ((uint32_t)j << l)] = ent;
            PP_LANES_END
        } else {
            const uint32_t p = code >> (l - root);
            const uint32_t pe = tbl[bitrev(p, root)];
            const int sub = (int)e_sub(pe);
            const uint32_t start = e_val(pe);
            const int sl = l - root;  // bits of this code inside the sub-table
            const uint32_t base = bitrev(code & ((1u << sl) - 1u), sl);
            const int reps = 1 << (sub - sl);
            PP_LANES_BEGIN
            for (int j = lane; j < reps; j += 32) tbl[start + (base | ((uint32_t)j << sl))] = ent;
            PP_LANES_END
        }
    }
    return 0;
}

// ---- output side --------------------------------------------------------------
struct Out {
    uint8_t *base;     // &slot[lead_len]: output byte 0; history is at negative offsets
    uint32_t pos;      // bytes produced
    uint32_t len;      // bytes wanted
    // one deferred match: its loads are issued, its stores wait for the next token so the
    // load latency overlaps the next Huffman decode
    uint32_t p_len;    // 0 = nothing pending
    uint32_t p_dst;
    PP_LV(uint32_t, p_val);
    PP_LV(uint32_t, nl);   // '\n' bytes stored by this lane
    PP_LV(uint32_t, mn);   // min byte stored by this lane
};

PP_DEV void out_flush(Out &o)
{
    if (o.p_len) {
        PP_LANES_BEGIN
        if ((uint32_t)lane < o.p_len) {
            const uint32_t v = PP_L(o.p_val);
            o.base[o.p_dst + (uint32_t)lane] = (uint8_t)v;
            PP_L(o.nl) += (v == 10u);
            PP_L(o.mn) = v < PP_L(o.mn) ? v : PP_L(o.mn);
        }
        PP_LANES_END
        o.p_len = 0;
    }
}

PP_DEV void out_literal(Out &o, uint32_t v)
{
    PP_LANE0_BEGIN
    o.base[o.pos] = (uint8_t)v;
    PP_L0(o.nl) += (v == 10u);
    PP_L0(o.mn) = v < PP_L0(o.mn) ? v : PP_L0(o.mn);
    PP_LANE0_END
    o.pos++;
}

// LZ77 copy of `len` bytes from `dist` back.  len is clamped to the space left.
PP_DEV void out_match(Out &o, uint32_t len, uint32_t dist)
{
    const uint32_t room = o.len - o.pos;
    if (len > room) len = room;
    if (len == 0) return;
    // the deferred token's bytes are not in memory yet: store them first if this match reads them
    if (o.p_len && dist < (o.pos - o.p_dst) + len) out_flush(o);
    const uint8_t *src = o.base + o.pos - dist;  // may point into the window (negative offset)
    if (len <= 32u && dist >= len) {
        PP_LV(uint32_t, v);
        PP_LANES_BEGIN
        PP_L(v) = (uint32_t)lane < len ? src[lane] : 0u;
        PP_LANES_END
        out_flush(o);
        PP_LANES_BEGIN
        PP_L(o.p_val) = PP_L(v);
        PP_LANES_END
        o.p_len = len;
        o.p_dst = o.pos;
    } else {
        out_flush(o);
        uint8_t *dst = o.base + o.pos;
        if (dist >= 32u || dist >= len) {
            // every 32-byte step reads bytes written by earlier steps (or earlier tokens) only
            for (uint32_t b = 0; b < len; b += 32u) {
                PP_LANES_BEGIN
                const uint32_t j = b + (uint32_t)lane;
                if (j < len) {
                    const uint32_t v = src[j];
                    dst[j] = (uint8_t)v;
                    PP_L(o.nl) += (v == 10u);
                    PP_L(o.mn) = v < PP_L(o.mn) ? v : PP_L(o.mn);
                }
                PP_LANES_END
            }
        } else {
            // overlapping run (dist < len, dist < 32): the output is the last `dist` bytes repeated
            for (uint32_t b = 0; b < len; b += 32u) {
                PP_LANES_BEGIN
                const uint32_t j = b + (uint32_t)lane;
                if (j < len) {
                    const uint32_t v = src[dist == 1u ? 0u : j % dist];
                    dst[j] = (uint8_t)v;
                    PP_L(o.nl) += (v == 10u);
                    PP_L(o.mn) = v < PP_L(o.mn) ? v : PP_L(o.mn);
                }
                PP_LANES_END
            }
        }
    }
    o.pos += len;
}

// ---- block decoders -------------------------------------------------------------
PP_DEV int stored_block(Reader &r, Smem &sm, Out &o)
{
    rd_consume(r, (uint32_t)r.bitcnt & 7u);  // to the byte boundary
    rd_refill(r, sm);
    const uint32_t v = rd_lo(r);
    rd_consume(r, 32);
    const uint32_t len = v & 0xffffu, nlen = v >> 16;
    if ((len ^ 0xffffu) != nlen) return -3;  // invalid stored block lengths
    const uint64_t byte0 = rd_bitpos(r) >> 3;
    out_flush(o);
    uint32_t n = len;
    if (n > o.len - o.pos) n = o.len - o.pos;
    if (byte0 + len > r.comp_tiles * (uint64_t)kTileBytes) return -3;  // input exhausted
    const uint8_t *src = r.comp + byte0;
    uint8_t *dst = o.base + o.pos;
    for (uint32_t b = 0; b < n; b += 32u) {
        PP_LANES_BEGIN
        const uint32_t j = b + (uint32_t)lane;
        if (j < n) {
            const uint32_t c = src[j];
            dst[j] = (uint8_t)c;
            PP_L(o.nl) += (c == 10u);
            PP_L(o.mn) = c < PP_L(o.mn) ? c : PP_L(o.mn);
        }
        PP_LANES_END
    }
    o.pos += n;
    rd_seek(r, sm, (byte0 + len) * 8ull, false);
    return 0;
}

PP_DEV int fixed_tables(Smem &sm)
{
    PP_LANES_BEGIN
    for (int s = lane; s < 288; s += 32) sm.lens[s] = (uint8_t)(s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8);
    PP_LANES_END
    int rc = build_table(sm, sm.lit, kRootL, kLitCap, 288, 0, false);
    if (rc) return rc;
    PP_LANES_BEGIN
    if (lane < 32) sm.lens[lane] = 5;
    PP_LANES_END
    // zlib's fixed distance table is the 5-bit complete code over 32 symbols (30/31 invalid)
    return build_table(sm, sm.dist, kRootD, kDistCap, 32, 0, true);
}

PP_DEV int dynamic_tables(Reader &r, Smem &sm)
{
    const uint32_t nlen = rd_bits(r, sm, 5) + 257u;
    const uint32_t ndist = rd_bits(r, sm, 5) + 1u;
    const uint32_t ncode = rd_bits(r, sm, 4) + 4u;
    if (nlen > 286u || ndist > 30u) return -3;  // too many length or distance symbols
    // code-length code: 19 symbols, 3-bit lengths, stored at lens[288..307)
    PP_LANES_BEGIN
    if (lane < 19) sm.lens[288 + lane] = 0;
    PP_LANES_END
    for (uint32_t i = 0; i < ncode; i++) {
        const uint32_t l = rd_bits(r, sm, 3);
        PP_LANE0_BEGIN
        sm.lens[288 + kClOrder[i]] = (uint8_t)l;
        PP_LANE0_END
    }
    // zlib builds this table with root 7 and rejects incomplete sets outright (type CODES)
    {
        int left = 1, any = 0;
        uint32_t cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int s = 0; s < 19; s++) cnt[sm.lens[288 + s]]++;
        for (int l = 1; l <= 7; l++) {
            left <<= 1;
            left -= (int)cnt[l];
            any |= (int)cnt[l];
            if (left < 0) return -3;
        }
        if (left > 0) return -3;  // invalid code lengths set (also covers the empty set)
        (void)any;
    }
    // reuse the distance table storage for the 7-bit code-length table
    int rc = build_table(sm, sm.dist, 7, kDistCap, 19, 288, false /*symbols 0..18 act as literals*/);
    if (rc) return rc;
    uint32_t have = 0;
    const uint32_t total = nlen + ndist;
    while (have < total) {
        rd_refill(r, sm);
        const uint32_t e = sm.dist[rd_lo(r) & 127u];
        if (e_kind(e) != K_LIT) return -3;
        rd_consume(r, e_tot(e));
        const uint32_t sym = e_val(e);
        if (sym < 16u) {
            PP_LANE0_BEGIN
            sm.lens[have] = (uint8_t)sym;
            PP_LANE0_END
            have++;
        } else {
            uint32_t rep, val = 0;
            if (sym == 16u) {
                if (have == 0) return -3;  // invalid bit length repeat
                val = sm.lens[have - 1];
                rep = 3u + rd_bits(r, sm, 2);
            } else if (sym == 17u) {
                rep = 3u + rd_bits(r, sm, 3);
            } else {
                rep = 11u + rd_bits(r, sm, 7);
            }
            if (have + rep > total) return -3;  // invalid bit length repeat
            PP_LANES_BEGIN
            for (uint32_t j = (uint32_t)lane; j < rep; j += 32u) sm.lens[have + j] = (uint8_t)val;
            PP_LANES_END
            have += rep;
        }
    }
    if (sm.lens[256] == 0) return -3;  // invalid code -- missing end-of-block
    // distance lengths follow the literal/length lengths: move them out of the way first
    {
        PP_LV(uint8_t, t);
        PP_LANES_BEGIN
        PP_L(t) = (uint32_t)lane < ndist ? sm.lens[nlen + lane] : (uint8_t)0;
        PP_LANES_END
        PP_LANES_BEGIN
        if ((uint32_t)lane < ndist) sm.lens[288 + lane] = PP_L(t);
        PP_LANES_END
    }
    rc = build_table(sm, sm.lit, kRootL, kLitCap, (int)nlen, 0, false);
    if (rc) return rc;
    return build_table(sm, sm.dist, kRootD, kDistCap, (int)ndist, 288, true);
}

// Decode symbols of one Huffman block until end-of-block or the output is full.
PP_DEV int huffman_block(Reader &r, Smem &sm, Out &o)
{
    for (;;) {
        if (o.pos >= o.len) return 0;
        if (r.exhausted) return -3;
        rd_refill(r, sm);
        uint32_t lo = rd_lo(r);
        uint32_t e = sm.lit[lo & ((1u << kRootL) - 1u)];
        if (e_kind(e) == K_SUB) e = sm.lit[e_val(e) + ((lo >> kRootL) & ((1u << e_sub(e)) - 1u))];
        const uint32_t kind = e_kind(e);
        const uint32_t tot = e_tot(e);
        if (kind == K_LIT) {
            rd_consume(r, tot);
            out_literal(o, e_val(e));
            continue;
        }
        if (kind == K_BASE) {
            const uint32_t len = e_val(e) + ((lo & ~(0xffffffffu << tot)) >> e_cl(e));
            rd_consume(r, tot);
            rd_refill(r, sm);
            lo = rd_lo(r);
            uint32_t d = sm.dist[lo & ((1u << kRootD) - 1u)];
            if (e_kind(d) == K_SUB) d = sm.dist[e_val(d) + ((lo >> kRootD) & ((1u << e_sub(d)) - 1u))];
            if (e_kind(d) != K_BASE) return -3;  // invalid distance code
            const uint32_t dtot = e_tot(d);
            const uint32_t dist = e_val(d) + ((lo & ~(0xffffffffu << dtot)) >> e_cl(d));
            rd_consume(r, dtot);
            // dist <= 32768 <= lead_len always, so "distance too far back" cannot occur:
            // the reference primes a full 32 KB dictionary (Core.cs:158)
            out_match(o, len, dist);
            continue;
        }
        if (kind == K_EOB) {
            rd_consume(r, tot);
            return 1;
        }
        return -3;  // invalid literal/length code
    }
}

// Whole chunk: Core.ExtractDeflateIndex for one (from, to) pair.
PP_DEV void inflate_chunk(const ChunkDesc &d, const uint8_t *comp, uint64_t comp_bytes, uint8_t *slots,
                          const uint8_t *lead_src, Smem &sm, ChunkResult &res)
{
    uint8_t *slot = slots + d.slot_off;
    // 1. history: copy the checkpoint window (Core.cs:158 inflateSetDictionary) in front of the output
    {
        const uint4 *s4 = reinterpret_cast<const uint4 *>(lead_src + d.lead_src);
        uint4 *d4 = reinterpret_cast<uint4 *>(slot);
        const uint32_t n4 = d.lead_len / 16u;
        PP_LANES_BEGIN
        for (uint32_t i = (uint32_t)lane; i < n4; i += 32u) d4[i] = s4[i];
        PP_LANES_END
    }
    Out o;
    o.base = slot + d.lead_len;
    o.pos = 0;
    o.len = d.out_len;
    o.p_len = 0;
    o.p_dst = 0;
    PP_LANES_BEGIN
    PP_L(o.p_val) = 0;
    PP_L(o.nl) = 0;
    PP_L(o.mn) = 255u;
    PP_LANES_END

    Reader r;
    r.comp = comp;
    r.comp_tiles = comp_bytes / kTileBytes;
    r.s_cur = 0;
    r.s_issued = 0;
    r.tile_bias = 0;
    r.exhausted = 0;
    // 2. bit cursor: 8*Input - Bits (Core.cs:151-157 inflatePrime semantics)
    rd_seek(r, sm, d.in_bit, true);

    int status = 0;
    while (o.pos < o.len) {
        if (r.exhausted || (rd_bitpos(r) >> 3) > d.in_limit) { status = -3; break; }  // Core.cs:174
        rd_refill(r, sm);
        const uint32_t hdr = rd_lo(r) & 7u;
        rd_consume(r, 3);
        const uint32_t last = hdr & 1u, type = hdr >> 1;
        int rc;
        if (type == 0u) {
            rc = stored_block(r, sm, o);
        } else if (type == 1u || type == 2u) {
            rc = type == 1u ? fixed_tables(sm) : dynamic_tables(r, sm);
            if (rc == 0) rc = huffman_block(r, sm, o);
            if (rc == 1) rc = 0;
        } else {
            rc = -3;  // invalid block type
        }
        if (rc < 0) { status = rc; break; }
        if (last) break;  // Z_STREAM_END (Core.cs:185)
    }
    out_flush(o);
    // 3. NUL terminator / clean tail for the parse stage
    {
        const uint32_t from = o.pos;
        const uint32_t to = ((d.lead_len + d.out_len + 1u + 127u) & ~127u) - d.lead_len;
        PP_LANES_BEGIN
        for (uint32_t i = from + (uint32_t)lane; i < to; i += 32u) o.base[i] = 0;
        PP_LANES_END
    }
    // 4. results
    uint32_t nl = 0, mn = 255u;
#ifdef PP_HOST_EMU
    for (int lane = 0; lane < 32; lane++) { nl += o.nl[lane]; mn = o.mn[lane] < mn ? o.mn[lane] : mn; }
#else
    nl = o.nl;
    mn = o.mn;
    for (int s = 16; s > 0; s >>= 1) {
        nl += __shfl_xor_sync(0xffffffffu, nl, s);
        const uint32_t m2 = __shfl_xor_sync(0xffffffffu, mn, s);
        mn = m2 < mn ? m2 : mn;
    }
#endif
    PP_LANE0_BEGIN
    res.status = status;
    res.produced = o.pos;
    res.newlines = nl;
    res.min_byte = mn;
    res.end_bit = rd_bitpos(r);
    PP_LANE0_END
}

}  // namespace ppinf

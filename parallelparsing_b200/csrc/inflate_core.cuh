// Checkpoint inflate — the DEFLATE (RFC 1951) decoder that replaces zlib's
// inflate() on the Decompress(checkpoint) path (Core.ExtractDeflateIndex,
// Decompressor/Core.cs:133-192; zlib reached through Interop/PlatformInterop.cs:9-34).
//
// Execution model: ONE WARP decodes ONE index chunk, alternating two phases.
//
//  DECODE  All 32 lanes run the serial Huffman decode redundantly on identical
//          registers (table lookups are shared-memory broadcasts), so every lane
//          knows every token without a queue, shuffle or barrier.  Nothing touches
//          global memory here: each token is written into a shared-memory SOURCE MAP
//          — one u16 per output byte saying "literal v" or "copy from d bytes back" —
//          with lane j writing byte j of the token.  The decoder also records where
//          a token reads bytes produced since the last such point ("rounds").
//  COPY    About 1 KB of output at a time is resolved with all lanes busy:
//          (1) every byte whose source lies before the batch is gathered from the
//              history in global memory — all loads independent and in flight together,
//              so the L2 latency is paid once per batch, not once per match;
//          (2) bytes whose source lies inside the batch are resolved from shared memory,
//              round by round;
//          (3) the finished bytes leave with aligned 16-byte vector stores, and the same
//              pass counts '\n' and looks for NUL bytes for the parse stage.
//
// History is addressed directly in global memory: the chunk's output slot is laid out
// as [window (>= 32 KB)][output], so a back-reference is out[pos-dist] with no wrap test.
// Compressed bytes stream through a shared-memory ring filled by TMA bulk copies
// (cp.async.bulk + mbarrier complete_tx), kStages tiles ahead of the bit reader.
//
// The same source compiles in two modes:
//   * device (default): used by inflate.cu, the only mode shipped in libppb200.so;
//   * PP_HOST_EMU: lanes are run one after another by a plain loop.  Compiled ONLY by
//     tests/emu/ to check the decoder logic against zlib on machines without a GPU.
//     It is test scaffolding, not a fallback: nothing in the product links it.
// Lane-parallel sections never read a byte written in the same section, which is what
// makes the sequential emulation equivalent to lockstep execution.
#pragma once
#include <stdint.h>

#ifdef PP_HOST_EMU
#include <string.h>
#define PP_DEV static inline
#define PP_LANES_BEGIN for (int lane = 0; lane < 32; ++lane) {
#define PP_LANES_END }
#define PP_LANES_END_NOSYNC }
#define PP_WARP_SYNC()
#define PP_LANE0_BEGIN {
#define PP_LANE0_END }
#define PP_LV(T, name) T name[32]
#define PP_L(name) name[lane]
#define PP_CONST static const
#else
#define PP_DEV __device__ __forceinline__
#define PP_LANES_BEGIN { const int lane = (int)(threadIdx.x & 31u);
#define PP_LANES_END } __syncwarp();
#define PP_LANES_END_NOSYNC }
#define PP_WARP_SYNC() __syncwarp()
#define PP_LANE0_BEGIN if ((threadIdx.x & 31u) == 0) {
#define PP_LANE0_END } __syncwarp();
#define PP_LV(T, name) T name
#define PP_L(name) name
#define PP_CONST __device__ const
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
constexpr int kBatch = 1024;           // output bytes resolved per copy phase (multiple of 16)
constexpr int kMaxMatch = 258;
constexpr int kMaxRounds = 192;

// ---- table entry ------------------------------------------------------------
// [4:0] bits to consume (code + extra)  [7:5] kind  [12:8] code length
// [15:13] sub-table index bits (kind SUB)
// [31:16] literal: 0x8000|byte (a ready-made source-map entry) / base value / sub-table start
enum : uint32_t { K_LIT = 0, K_BASE = 1, K_SUB = 2, K_EOB = 3, K_BAD = 4 };
PP_DEV uint32_t mk_entry(uint32_t kind, uint32_t tot, uint32_t cl, uint32_t val)
{
    return tot | (kind << 5) | (cl << 8) | (val << 16);
}
PP_DEV uint32_t e_kind(uint32_t e) { return (e >> 5) & 7u; }
PP_DEV uint32_t e_tot(uint32_t e) { return e & 31u; }
PP_DEV uint32_t e_cl(uint32_t e) { return (e >> 8) & 31u; }
PP_DEV uint32_t e_sub(uint32_t e) { return (e >> 13) & 7u; }
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
    uint32_t min_byte;  // 0 when a NUL byte was written (=> exact parser), else 1
    uint64_t end_bit;   // bit position after the last consumed bit
};

struct alignas(128) Smem {
    uint32_t ring[kRingWords];        // first: TMA destinations must be 16 B aligned
    uint32_t lit[kLitCap];
    uint32_t dist[kDistCap];
    alignas(16) uint8_t obuf[kBatch + kMaxMatch + 32];   // resolved output bytes of the batch
    uint16_t map[kBatch + kMaxMatch + 32];                // source map of the batch
    uint16_t rounds[kMaxRounds + 2];                      // batch-relative start of every round
    uint16_t count[16];
    uint16_t next[16];
    uint8_t lens[320];
#ifndef PP_HOST_EMU
    unsigned long long bar[kStages];
#endif
};

#ifdef PP_HOST_EMU
static Smem g_sm;
#else
__shared__ Smem g_sm;
#endif

// RFC 1951 3.2.5 length / distance bases and extra-bit counts
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
PP_DEV uint32_t fsr(uint32_t lo, uint32_t hi, uint32_t n)  // low 32 bits of (hi:lo) >> (n & 31)
{
#ifdef PP_HOST_EMU
    n &= 31u;
    return n ? (lo >> n) | (hi << (32u - n)) : lo;
#else
    return __funnelshift_r(lo, hi, n);
#endif
}
PP_DEV uint32_t fsl_hi(uint32_t lo, uint32_t n)  // high 32 bits of (0:lo) << (n & 31)
{
#ifdef PP_HOST_EMU
    n &= 31u;
    return n ? lo >> (32u - n) : 0u;
#else
    return __funnelshift_l(lo, 0u, n);
#endif
}

// ---- compressed-input reader ------------------------------------------------
// (hi:lo) holds the next `bitcnt` bits of the stream, lo always fully valid
// (bitcnt >= 32 between tokens); words come from the ring one at a time and the
// word after the current one is always preloaded in `nw`.
struct Reader {
    const uint8_t *comp;   // compressed buffer (global, 16 B aligned)
    uint64_t comp_tiles;   // tiles available in the buffer
    uint32_t lo, hi;
    int bitcnt;            // valid bits in hi:lo
    uint64_t wnext;        // word index (from comp) of the word held in nw
    uint32_t nw;           // preloaded word
    uint32_t rpos;         // ring position of wnext
    uint32_t s_cur;        // sequence number of the tile holding wnext
    uint32_t s_issued;     // tiles issued so far
    int64_t tile_bias;     // tile index = tile_bias + sequence number
    int exhausted;         // 1: ran past the buffer, 2: a transfer never landed
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

PP_DEV void rd_issue(Reader &r)
{
    // issue the tile with sequence number s_issued (if the buffer still has it)
    const int64_t tile = r.tile_bias + (int64_t)r.s_issued;
    const uint32_t stage = r.s_issued % kStages;
    if (tile >= 0 && (uint64_t)tile < r.comp_tiles) {
#ifdef PP_HOST_EMU
        memcpy(&g_sm.ring[stage * kTileWords], r.comp + (uint64_t)tile * kTileBytes, kTileBytes);
#else
        __syncwarp();  // every lane is done reading the stage being overwritten
        if ((threadIdx.x & 31u) == 0) {
            mbar_expect_tx(&g_sm.bar[stage], kTileBytes);
            tma_load_tile(&g_sm.ring[stage * kTileWords], r.comp + (uint64_t)tile * kTileBytes, &g_sm.bar[stage]);
        }
#endif
    } else {
#ifndef PP_HOST_EMU
        // nothing to load: complete the phase by hand so waiters do not hang
        if ((threadIdx.x & 31u) == 0)
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&g_sm.bar[stage])) : "memory");
#endif
    }
    r.s_issued++;
}

PP_DEV void rd_wait(Reader &r, uint32_t seq)
{
#ifndef PP_HOST_EMU
    if (!mbar_wait(&g_sm.bar[seq % kStages], (seq / kStages) & 1u)) r.exhausted = 2;
#endif
    const int64_t tile = r.tile_bias + (int64_t)seq;
    if (tile < 0 || (uint64_t)tile >= r.comp_tiles) r.exhausted |= 1;
}

// move the preloaded word one further (the tile crossing is the rare, out-of-line part)
PP_DEV void rd_cross_tile(Reader &r)
{
    if (r.rpos == kRingWords) r.rpos = 0;
    r.s_cur++;
    rd_wait(r, r.s_cur);   // the stage we enter must have landed
    rd_issue(r);           // the stage we leave is free again
}
PP_DEV void rd_advance(Reader &r)
{
    r.wnext++;
    r.rpos++;
    if ((r.rpos & (kTileWords - 1)) == 0) rd_cross_tile(r);
    r.nw = g_sm.ring[r.rpos];
}

// Position the reader at absolute bit `bit` of the compressed buffer.
PP_DEV void rd_seek(Reader &r, uint64_t bit, bool first)
{
    if (!first) {
        // drain tiles still in flight so the ring can be re-targeted
        for (uint32_t s = r.s_cur + 1; s < r.s_issued; s++) rd_wait(r, s);
    }
    const uint64_t word = bit >> 5;
    const uint64_t tile = word / kTileWords;
    r.tile_bias = (int64_t)tile - (int64_t)r.s_issued;
    r.s_cur = r.s_issued;
    r.exhausted &= 2;  // a timed-out transfer stays fatal
    for (int i = 0; i < kStages; i++) rd_issue(r);
    rd_wait(r, r.s_cur);
    r.wnext = word;
    r.rpos = (r.s_cur % kStages) * kTileWords + (uint32_t)(word % kTileWords);
    r.nw = g_sm.ring[r.rpos];
    // first word, minus the bits in front of `bit`; then top up to >= 32 valid bits
    const uint32_t sh = (uint32_t)(bit & 31ull);
    r.lo = r.nw >> sh;
    r.hi = 0;
    r.bitcnt = 32 - (int)sh;
    rd_advance(r);
    if (r.bitcnt < 32) {
        r.lo |= r.nw << r.bitcnt;
        r.hi = fsl_hi(r.nw, (uint32_t)r.bitcnt);
        r.bitcnt += 32;
        rd_advance(r);
    }
}

// restore the invariant bitcnt >= 32 (lo fully valid)
PP_DEV void rd_refill(Reader &r)
{
    if (r.bitcnt < 32) {
        r.lo |= r.nw << r.bitcnt;
        r.hi = fsl_hi(r.nw, (uint32_t)r.bitcnt);
        r.bitcnt += 32;
        rd_advance(r);
    }
}
PP_DEV void rd_consume(Reader &r, uint32_t n)  // n < 32
{
    r.lo = fsr(r.lo, r.hi, n);
    r.hi >>= n;
    r.bitcnt -= (int)n;
}
PP_DEV uint32_t rd_bits(Reader &r, uint32_t n)  // n <= 16
{
    const uint32_t v = r.lo & ((1u << n) - 1u);
    rd_consume(r, n);
    rd_refill(r);
    return v;
}
PP_DEV uint64_t rd_bitpos(const Reader &r) { return r.wnext * 32ull - (uint64_t)r.bitcnt; }

// ---- Huffman table construction ---------------------------------------------
// Canonical codes (RFC 1951 3.2.2) into a two-level lookup table: `root` primary
// bits, sub-tables sized by the longest code under each primary prefix.
// Validity rules follow zlib's inflate_table: over-subscribed sets are rejected,
// incomplete sets are rejected unless the set is a single 1-bit code (or, for
// distances, empty).  Returns 0 or -3 (Z_DATA_ERROR).
// mode 0: literal/length alphabet, 1: distance alphabet, 2: code-length alphabet.
PP_DEV int build_table(uint32_t *tbl, int root, int cap, int nsym, int lens_off, int mode)
{
    Smem &sm = g_sm;
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
    if (left > 0 && maxlen != 1 && !(mode == 1 && maxlen == 0)) return -3;  // incomplete set
    if (left > 0 && mode == 2) return -3;                                   // zlib: CODES must be complete
    PP_LANE0_BEGIN
    {
        uint32_t code = 0;
        sm.next[0] = 0;
        for (int l = 1; l <= 15; l++) {
            code = (code + (l > 1 ? (uint32_t)sm.count[l - 1] : 0u)) << 1;
            sm.next[l] = (uint16_t)code;
        }
    }
    PP_LANE0_END
    // 3. primary table: start from "invalid code" everywhere (incomplete sets leave holes)
    const int nprim = 1 << root;
    PP_LANES_BEGIN
    for (int i = lane; i < nprim; i += 32) tbl[i] = mk_entry(K_BAD, 1, 1, 0);
    PP_LANES_END
    // 4. sub-table geometry for prefixes that own codes longer than `root`.
    // Prefix p (root bits, MSB-first code order) owns the length-l codes in
    // [p << (l-root), (p+1) << (l-root)); it needs a sub-table when that range meets
    // [first[l], first[l]+count[l]) for some l > root.  Long codes sit at the top of the
    // code space, so only prefixes from the first long code's prefix upward are visited,
    // in code order (sub-tables come out in canonical order).
    int used = nprim;
    if (maxlen > root) {
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
                tbl[idx] = mk_entry(K_SUB, (uint32_t)root, (uint32_t)root, (uint32_t)used) | ((uint32_t)sub << 13);
                PP_LANE0_END
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
        if (mode == 0) {
            if (s < 256) { kind = K_LIT; tot = (uint32_t)l; val = 0x8000u | (uint32_t)s; }
            else if (s == 256) { kind = K_EOB; tot = (uint32_t)l; val = 0; }
            else if (s < 286) { kind = K_BASE; tot = (uint32_t)l + kLenExtra[s - 257]; val = kLenBase[s - 257]; }
            else { kind = K_BAD; tot = (uint32_t)l; val = 0; }
        } else if (mode == 1) {
            if (s < 30) { kind = K_BASE; tot = (uint32_t)l + kDistExtra[s]; val = kDistBase[s]; }
            else { kind = K_BAD; tot = (uint32_t)l; val = 0; }
        } else {
            kind = K_LIT; tot = (uint32_t)l; val = (uint32_t)s;
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
// A batch covers output bytes [abase, abase + bpos): abase is 16-byte aligned, the first
// `carry` bytes are already resolved (the unaligned tail of the previous batch).
struct Out {
    uint8_t *base;       // &slot[lead_len]: output byte 0; history is at negative offsets
    uint32_t len;        // bytes wanted
    uint32_t abase;      // output offset of batch byte 0 (multiple of 16)
    uint32_t bpos;       // bytes of the batch described so far (carry + decoded)
    uint32_t carry;      // resolved bytes at the front of the batch
    uint32_t rstart;     // batch offset where the current round began
    uint32_t nrounds;
    PP_LV(uint32_t, nl);   // '\n' bytes flushed by this lane
    PP_LV(uint32_t, nul);  // non-zero when this lane flushed a NUL byte
};

PP_DEV uint32_t out_pos(const Out &o) { return o.abase + o.bpos; }

// 0x80 in every byte of w that equals c (c replicated in all four bytes of c4)
PP_DEV uint32_t eq_bytes(uint32_t w, uint32_t c4)
{
    const uint32_t x = w ^ c4;
    return ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x | 0x7f7f7f7fu);
}
PP_DEV uint32_t popc32(uint32_t v)
{
#ifdef PP_HOST_EMU
    return (uint32_t)__builtin_popcount(v);
#else
    return (uint32_t)__popc(v);
#endif
}

// COPY phase: resolve the batch and move its whole 16-byte vectors to global memory.
// final: also write the unaligned tail (end of chunk / before a stored block).
PP_DEV void out_resolve(Out &o, bool final)
{
    Smem &sm = g_sm;
    const uint32_t n = o.bpos;
    const int32_t carry = (int32_t)o.carry;
    PP_WARP_SYNC();  // the source map written during DECODE is read by other lanes now
    // (1) literals and bytes whose source is already resolved (before the batch, or in the
    //     carried head): independent gathers
    {
        const uint8_t *hist = o.base + o.abase;  // batch byte i lives at hist[i]
        for (uint32_t b = o.carry; b < n; b += 128u) {
            PP_LANES_BEGIN
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t i = b + (uint32_t)u * 32u + (uint32_t)lane;
                if (i < n) {
                    const uint32_t e = sm.map[i];
                    const int32_t sidx = (int32_t)i - (int32_t)e - 1;  // batch offset of the source
                    if (e & 0x8000u) sm.obuf[i] = (uint8_t)e;
                    else if (sidx < 0) sm.obuf[i] = hist[sidx];
                    else if (sidx < carry) sm.obuf[i] = sm.obuf[sidx];
                }
            }
            PP_LANES_END
        }
    }
    // (2) bytes whose source is inside the batch: round by round out of shared memory
    for (uint32_t r = 0; r < o.nrounds; r++) {
        const uint32_t rs = sm.rounds[r], re = r + 1 < o.nrounds ? sm.rounds[r + 1] : n;
        for (uint32_t b = rs; b < re; b += 32u) {
            PP_LV(uint32_t, val);
            PP_LV(uint32_t, has);
            PP_LANES_BEGIN
            const uint32_t i = b + (uint32_t)lane;
            PP_L(has) = 0;
            PP_L(val) = 0;
            if (i < re) {
                const uint32_t e = sm.map[i];
                const int32_t sidx = (int32_t)i - (int32_t)e - 1;
                if (!(e & 0x8000u) && sidx >= carry) { PP_L(val) = sm.obuf[sidx]; PP_L(has) = 1; }
            }
            PP_LANES_END
            PP_LANES_BEGIN
            if (PP_L(has)) sm.obuf[b + (uint32_t)lane] = (uint8_t)PP_L(val);
            PP_LANES_END
        }
    }
    // (3) whole 16-byte vectors leave; '\n' and NUL are counted on the way
    const uint32_t nvec = n / 16u;
    {
        uint4 *dst = reinterpret_cast<uint4 *>(o.base + o.abase);
        const uint4 *src = reinterpret_cast<const uint4 *>(sm.obuf);
        PP_LANES_BEGIN
        for (uint32_t i = (uint32_t)lane; i < nvec; i += 32u) {
            const uint4 w = src[i];
            dst[i] = w;
            PP_L(o.nl) += popc32(eq_bytes(w.x, 0x0a0a0a0au)) + popc32(eq_bytes(w.y, 0x0a0a0a0au)) +
                          popc32(eq_bytes(w.z, 0x0a0a0a0au)) + popc32(eq_bytes(w.w, 0x0a0a0a0au));
            PP_L(o.nul) |= eq_bytes(w.x, 0u) | eq_bytes(w.y, 0u) | eq_bytes(w.z, 0u) | eq_bytes(w.w, 0u);
        }
        PP_LANES_END
    }
    const uint32_t done = nvec * 16u, rem = n - done;
    if (final) {
        PP_LANES_BEGIN
        if ((uint32_t)lane < rem) {
            const uint32_t c = sm.obuf[done + (uint32_t)lane];
            o.base[o.abase + done + (uint32_t)lane] = (uint8_t)c;
            PP_L(o.nl) += (c == 10u);
            PP_L(o.nul) |= (c == 0u);
        }
        PP_LANES_END
        o.abase += n;  // may be unaligned now: only a stored block or the end of the chunk follows
        o.carry = 0;
    } else {
        // the unaligned tail stays in shared memory as the head of the next batch
        PP_LV(uint32_t, t);
        PP_LANES_BEGIN
        PP_L(t) = (uint32_t)lane < rem ? sm.obuf[done + (uint32_t)lane] : 0u;
        PP_LANES_END
        PP_LANES_BEGIN
        if ((uint32_t)lane < rem) sm.obuf[lane] = (uint8_t)PP_L(t);
        PP_LANES_END
        o.abase += done;
        o.carry = rem;
    }
    o.bpos = o.carry;
    o.rstart = o.carry;
    o.nrounds = 0;
}

// ---- block decoders -------------------------------------------------------------
// Stored block: bytes go straight from the compressed buffer to the output.
PP_DEV int stored_block(Reader &r, Out &o)
{
    rd_consume(r, (uint32_t)r.bitcnt & 7u);  // to the byte boundary
    rd_refill(r);
    const uint32_t v = r.lo;
    const uint32_t len = v & 0xffffu, nlen = v >> 16;
    if ((len ^ 0xffffu) != nlen) return -3;  // invalid stored block lengths
    const uint64_t byte0 = (rd_bitpos(r) >> 3) + 4u;
    out_resolve(o, true);
    uint32_t n = len;
    if (n > o.len - o.abase) n = o.len - o.abase;
    if (byte0 + len > r.comp_tiles * (uint64_t)kTileBytes) return -3;  // input exhausted
    const uint8_t *src = r.comp + byte0;
    uint8_t *dst = o.base + o.abase;
    const uint32_t end = o.abase + n;
    const uint32_t al = end & ~15u;  // the next batch must start on a 16-byte boundary
    for (uint32_t b = 0; b < n; b += 32u) {
        PP_LANES_BEGIN
        const uint32_t j = b + (uint32_t)lane;
        if (j < n) {
            const uint32_t c = src[j];
            dst[j] = (uint8_t)c;
            // bytes past the last boundary are counted when the next batch flushes them
            if (o.abase + j < al) {
                PP_L(o.nl) += (c == 10u);
                PP_L(o.nul) |= (c == 0u);
            }
        }
        PP_LANES_END
    }
    // pull the bytes past the boundary back into shared memory as the resolved head of the
    // next batch (fewer than 16)
    {
        PP_LV(uint32_t, t);
        PP_LANES_BEGIN
        PP_L(t) = (uint32_t)lane < end - al ? o.base[al + (uint32_t)lane] : 0u;
        PP_LANES_END
        PP_LANES_BEGIN
        if ((uint32_t)lane < end - al) g_sm.obuf[lane] = (uint8_t)PP_L(t);
        PP_LANES_END
    }
    if (al < o.abase) {
        // a short block: the boundary lies before it.  Bytes [al, abase) were counted by the
        // final flush above and will be counted again with the next vector: take them out once.
        PP_LANES_BEGIN
        if ((uint32_t)lane < o.abase - al) PP_L(o.nl) -= (g_sm.obuf[lane] == 10u);
        PP_LANES_END
    }
    o.carry = end - al;
    o.abase = al;
    o.bpos = o.carry;
    o.rstart = o.carry;
    o.nrounds = 0;
    rd_seek(r, (byte0 + len) * 8ull, false);
    return 0;
}

PP_DEV int fixed_tables()
{
    Smem &sm = g_sm;
    PP_LANES_BEGIN
    for (int s = lane; s < 288; s += 32) sm.lens[s] = (uint8_t)(s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8);
    PP_LANES_END
    int rc = build_table(sm.lit, kRootL, kLitCap, 288, 0, 0);
    if (rc) return rc;
    PP_LANES_BEGIN
    sm.lens[lane] = 5;
    PP_LANES_END
    // zlib's fixed distance table is the 5-bit complete code over 32 symbols (30/31 invalid)
    return build_table(sm.dist, kRootD, kDistCap, 32, 0, 1);
}

PP_DEV int dynamic_tables(Reader &r)
{
    Smem &sm = g_sm;
    const uint32_t nlen = rd_bits(r, 5) + 257u;
    const uint32_t ndist = rd_bits(r, 5) + 1u;
    const uint32_t ncode = rd_bits(r, 4) + 4u;
    if (nlen > 286u || ndist > 30u) return -3;  // too many length or distance symbols
    // code-length code: 19 symbols, 3-bit lengths, stored at lens[288..307)
    PP_LANES_BEGIN
    if (lane < 19) sm.lens[288 + lane] = 0;
    PP_LANES_END
    for (uint32_t i = 0; i < ncode; i++) {
        const uint32_t l = rd_bits(r, 3);
        PP_LANE0_BEGIN
        sm.lens[288 + kClOrder[i]] = (uint8_t)l;
        PP_LANE0_END
    }
    // zlib builds this table with root 7 and rejects incomplete sets outright (type CODES);
    // the distance table's storage is free at this point
    int rc = build_table(sm.dist, 7, kDistCap, 19, 288, 2);
    if (rc) return rc;
    uint32_t have = 0;
    const uint32_t total = nlen + ndist;
    while (have < total) {
        const uint32_t e = sm.dist[r.lo & 127u];
        if (e_kind(e) != K_LIT) return -3;
        rd_consume(r, e_tot(e));
        rd_refill(r);
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
                rep = 3u + rd_bits(r, 2);
            } else if (sym == 17u) {
                rep = 3u + rd_bits(r, 3);
            } else {
                rep = 11u + rd_bits(r, 7);
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
    rc = build_table(sm.lit, kRootL, kLitCap, (int)nlen, 0, 0);
    if (rc) return rc;
    return build_table(sm.dist, kRootD, kDistCap, (int)ndist, 288, 1);
}

// DECODE phase for one Huffman block: symbols -> source map, until end-of-block (returns 1),
// the output is complete (returns 0) or an error (-3).  Runs the copy phase whenever a
// batch fills up.
PP_DEV int huffman_block(Reader &r, Out &o)
{
    Smem &sm = g_sm;
    for (;;) {
        if (o.bpos >= (uint32_t)kBatch || o.nrounds >= (uint32_t)kMaxRounds) out_resolve(o, false);
        if (out_pos(o) >= o.len) return 0;
        if (r.exhausted) return -3;
        uint32_t lo = r.lo;
        uint32_t e = sm.lit[lo & ((1u << kRootL) - 1u)];
        if (e_kind(e) == K_SUB) e = sm.lit[e_val(e) + ((lo >> kRootL) & ((1u << e_sub(e)) - 1u))];
        const uint32_t kind = e_kind(e);
        const uint32_t tot = e_tot(e);
        if (kind == K_LIT) {
            rd_consume(r, tot);
            rd_refill(r);
            sm.map[o.bpos] = (uint16_t)e_val(e);  // every lane stores the same value: no branch, no sync
            o.bpos++;
            continue;
        }
        if (kind == K_BASE) {
            uint32_t len = e_val(e) + ((lo & ~(0xffffffffu << tot)) >> e_cl(e));
            rd_consume(r, tot);
            rd_refill(r);
            lo = r.lo;
            uint32_t d = sm.dist[lo & ((1u << kRootD) - 1u)];
            if (e_kind(d) == K_SUB) d = sm.dist[e_val(d) + ((lo >> kRootD) & ((1u << e_sub(d)) - 1u))];
            if (e_kind(d) != K_BASE) return -3;  // invalid distance code
            const uint32_t dtot = e_tot(d);
            const uint32_t dist = e_val(d) + ((lo & ~(0xffffffffu << dtot)) >> e_cl(d));
            rd_consume(r, dtot);
            rd_refill(r);
            // dist <= 32768 <= lead_len always, so "distance too far back" cannot occur:
            // the reference primes a full 32 KB dictionary (Core.cs:158)
            const uint32_t room = o.len - out_pos(o);
            if (len > room) len = room;
            // does the match read bytes produced since the current round began?
            const uint32_t span = len < dist ? len : dist;
            if ((int32_t)(o.bpos - dist + span) > (int32_t)o.rstart) {
                sm.rounds[o.nrounds] = (uint16_t)o.bpos;  // uniform store
                o.nrounds++;
                o.rstart = o.bpos;
            }
            if (dist >= len) {
                PP_LANES_BEGIN
                for (uint32_t j = (uint32_t)lane; j < len; j += 32u) sm.map[o.bpos + j] = (uint16_t)(dist - 1u);
                PP_LANES_END_NOSYNC
            } else {
                // overlapping run: byte j repeats the `dist` bytes before the match, so its
                // source is dist*(j/dist+1) back — always in front of the match itself
                PP_LANES_BEGIN
                for (uint32_t j = (uint32_t)lane; j < len; j += 32u)
                    sm.map[o.bpos + j] = (uint16_t)(dist * (j / dist + 1u) - 1u);
                PP_LANES_END_NOSYNC
            }
            o.bpos += len;
            continue;
        }
        if (kind == K_EOB) {
            rd_consume(r, tot);
            rd_refill(r);
            return 1;
        }
        return -3;  // invalid literal/length code
    }
}

// Whole chunk: Core.ExtractDeflateIndex for one (from, to) pair.
PP_DEV void inflate_chunk(const ChunkDesc &d, const uint8_t *comp, uint64_t comp_bytes, uint8_t *slots,
                          const uint8_t *lead_src, ChunkResult &res)
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
    o.len = d.out_len;
    o.abase = 0;
    o.bpos = 0;
    o.carry = 0;
    o.rstart = 0;
    o.nrounds = 0;
    PP_LANES_BEGIN
    PP_L(o.nl) = 0;
    PP_L(o.nul) = 0;
    PP_LANES_END

    Reader r;
    r.comp = comp;
    r.comp_tiles = comp_bytes / kTileBytes;
    r.s_cur = 0;
    r.s_issued = 0;
    r.tile_bias = 0;
    r.exhausted = 0;
    // 2. bit cursor: 8*Input - Bits (Core.cs:151-157 inflatePrime semantics)
    rd_seek(r, d.in_bit, true);

    int status = 0;
    while (out_pos(o) < o.len) {
        if (r.exhausted || (rd_bitpos(r) >> 3) > d.in_limit) { status = -3; break; }  // Core.cs:174
        const uint32_t hdr = r.lo & 7u;
        rd_consume(r, 3);
        rd_refill(r);
        const uint32_t last = hdr & 1u, type = hdr >> 1;
        int rc;
        if (type == 0u) {
            rc = stored_block(r, o);
        } else if (type == 1u || type == 2u) {
            rc = type == 1u ? fixed_tables() : dynamic_tables(r);
            if (rc == 0) rc = huffman_block(r, o);
            if (rc == 1) rc = 0;
        } else {
            rc = -3;  // invalid block type
        }
        if (rc < 0) { status = rc; break; }
        if (last) break;  // Z_STREAM_END (Core.cs:185)
    }
    out_resolve(o, true);
    const uint32_t produced = o.abase;
    // 3. NUL terminator / clean tail for the parse stage
    {
        const uint32_t to = ((d.lead_len + d.out_len + 1u + 127u) & ~127u) - d.lead_len;
        PP_LANES_BEGIN
        for (uint32_t i = produced + (uint32_t)lane; i < to; i += 32u) o.base[i] = 0;
        PP_LANES_END
    }
    // 4. results
    uint32_t nl = 0, nul = 0;
#ifdef PP_HOST_EMU
    for (int lane = 0; lane < 32; lane++) { nl += o.nl[lane]; nul |= o.nul[lane]; }
#else
    nl = o.nl;
    nul = o.nul;
    for (int s = 16; s > 0; s >>= 1) {
        nl += __shfl_xor_sync(0xffffffffu, nl, s);
        nul |= __shfl_xor_sync(0xffffffffu, nul, s);
    }
#endif
    PP_LANE0_BEGIN
    res.status = status;
    res.produced = produced;
    res.newlines = nl;
    res.min_byte = nul ? 0u : 1u;
    res.end_bit = rd_bitpos(r);
    PP_LANE0_END
}

}  // namespace ppinf

// Checkpoint inflate — the DEFLATE (RFC 1951) decoder that replaces zlib's
// inflate() on the Decompress(checkpoint) path (Core.ExtractDeflateIndex,
// Decompressor/Core.cs:133-192; zlib reached through Interop/PlatformInterop.cs:9-34).
//
// Execution model: ONE CTA decodes ONE index chunk, and the threads of the CTA share
// the work INSIDE every deflate block.  A serial Huffman decoder is a dependent chain
// of table lookups (hundreds of cycles per symbol on a GPU); instead the block's bit
// stream is cut into fixed-size sub-sequences, one per thread:
//
//  STAGE    the compressed window (T sub-sequences) is brought into shared memory with
//           TMA bulk copies (cp.async.bulk + mbarrier complete_tx);
//  HEADER   block header; dynamic code lengths; both lookup tables are built by all
//           threads (canonical-code search per table slot, no serial fill);
//  GUESS    every thread decodes from the first bit of its sub-sequence.  Only thread 0
//           starts on a real symbol boundary, but a Huffman decoder that starts at a
//           wrong bit falls back onto true boundaries after a few symbols
//           (self-synchronisation), so most threads END on a true boundary;
//  SYNC     every thread restarts where its predecessor ended, until nothing moves.
//           Thread 0 is right by construction, so after k rounds threads 0..k are
//           exact; with self-synchronisation two or three rounds settle all of them;
//  SCAN     block-wide prefix sum of the bytes each sub-sequence produces;
//  EMIT     every thread decodes its (now exact) sub-sequence once more and writes one 32-bit
//           TOKEN per symbol (literal byte, or match length + distance) into its row of the
//           CTA's global scratch, plus a group index entry per 16 output bytes;
//  RESOLVE  the CTA walks the window's output in tiles of 16 bytes per thread: sources
//           in front of the tile are gathered from global memory (all loads in flight
//           together), sources inside the tile are settled in shared memory by pointer
//           jumping, and the tile leaves with 16-byte stores ('\n' / NUL counted on the
//           way for the parse stage).
//
// History is addressed directly in global memory: the chunk's output slot is laid out
// as [window (>= 32 KB)][output], so a back-reference is out[pos-dist] with no wrap.
//
// The same source compiles in two modes:
//   * device (default): used by inflate.cu, the only mode shipped in libppb200.so;
//   * PP_HOST_EMU: threads are run one after another by plain loops, phase by phase.
//     Compiled ONLY by tests/emu/ to check the decoder logic against zlib on machines
//     without a GPU.  It is test scaffolding, not a fallback: nothing in the product
//     links it.  Phases never read what another thread writes in the same phase
//     (except the monotone pointer-jumping of RESOLVE, which is order-independent),
//     which is what makes the sequential emulation equivalent.
#pragma once
#include <stdint.h>

#ifdef PP_HOST_EMU
#include <string.h>
namespace ppinf { static int g_T = 64; static uint64_t g_stat[8]; }  // stat: windows, sync rounds, max rounds, blocks
#define PP_DEV static inline
#define PP_HD static inline
#define PP_NT (ppinf::g_T)
#define PP_FOR_T(t) for (int t = 0; t < PP_NT; ++t) {
#define PP_END_T }
#define PP_SYNC()
#define PP_SYNC_OR(v) (v)
#define PP_T0_BEGIN {
#define PP_T0_END }
#define PP_CONST static const
#define PP_ATOMIC_ADD(p, v) (*(p) += (v))
#define PP_ATOMIC_MIN(p, v) (*(p) = *(p) < (v) ? *(p) : (v))
// per-thread state that lives across a barrier: registers on the device, one row per thread here
// two warp-synchronous steps: every lane of a warp finishes the first before any starts the second
#define PP_FOR_W(t) for (int w_ = 0; w_ < PP_NT; w_ += 32) { for (int t = w_; t < w_ + 32 && t < PP_NT; ++t) {
#define PP_WARP_SPLIT(t) } for (int t = w_; t < w_ + 32 && t < PP_NT; ++t) {
#define PP_END_W } }
#define PP_TLS_DECL(type, name, n) static type name##_tls[1024][n]
#define PP_TLS(name) name##_tls[t]
#else
#define PP_DEV __device__ __forceinline__
#define PP_HD __host__ __device__ __forceinline__
#define PP_NT ((int)blockDim.x)
#define PP_FOR_T(t) { const int t = (int)threadIdx.x;
#define PP_END_T }
#define PP_SYNC() __syncthreads()
#define PP_SYNC_OR(v) __syncthreads_or(v)
#define PP_T0_BEGIN if (threadIdx.x == 0) {
#define PP_T0_END }
#define PP_CONST __device__ const
#define PP_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#define PP_ATOMIC_MIN(p, v) atomicMin((p), (v))
#define PP_FOR_W(t) { { const int t = (int)threadIdx.x;
#define PP_WARP_SPLIT(t) } __syncwarp(); { const int t = (int)threadIdx.x;
#define PP_END_W } }
#define PP_TLS_DECL(type, name, n) type name[n]
#define PP_TLS(name) name
#endif

namespace ppinf {

// ---- phase timers (thread 0's clock, summed over all CTAs; read with pp_internal_phase_cycles) ----
enum { PH_STAGE = 0, PH_HEADER, PH_GUESS, PH_SYNC, PH_SCAN, PH_EMIT, PH_RESOLVE, PH_STORED, PH_OTHER, PH_R_EXPAND, PH_R_GATHER, PH_R_CHASE, PH_H_PARSE, PH_H_LIT, PH_WAIT, PH_COUNT };  // PH_R_EXPAND is folded into PH_R_GATHER
#if defined(PP_HOST_EMU)
#define PP_PHASE(ph)
#else
__device__ unsigned long long g_phase_cycles[PH_COUNT];
#define PP_PHASE(ph)                                                               \
    do {                                                                           \
        if (threadIdx.x == 0) {                                                    \
            const long long now_ = clock64();                                      \
            atomicAdd(&g_phase_cycles[ph], (unsigned long long)(now_ - sm.u[24] - ((long long)sm.u[25] << 32))); \
            sm.u[24] = (uint32_t)now_;                                             \
            sm.u[25] = (uint32_t)((unsigned long long)now_ >> 32);                 \
        }                                                                          \
    } while (0)
#endif

// ---- geometry -------------------------------------------------------------
constexpr int kRootL = 10;              // primary bits, literal/length table
constexpr int kRootD = 8;               // primary bits, distance table
constexpr int kLitCap = 1024 + 512;     // zlib `enough 288 10 15` = 1334
constexpr int kDistCap = 256 + 256;     // zlib `enough 32 8 15` = 402
#ifndef PP_SUBW
#define PP_SUBW 31                      // words per sub-sequence; odd => the threads' word reads spread over all banks
#endif
constexpr int kSubW = PP_SUBW;
constexpr int kSubBits = kSubW * 32;
constexpr int kHdrWords = 160;          // a dynamic block header is < 4600 bits
constexpr int kSlackWords = 16;         // last symbol overrun (<= 48 bits) + reader look-ahead
constexpr int kMaxThreads = 1024;
constexpr int kTileB = 16;              // output bytes per thread per resolve tile

PP_HD uint32_t cw_words_for(int T) { return (uint32_t)(T * kSubW + kHdrWords + kSlackWords + 3) & ~3u; }
// Per-CTA scratch in global memory (u32 words): the tokens of a window — token k of
// sub-sequence s at tok[s * kTokRows + k] (a row is read by one warp, 32 tokens = 128 bytes
// per load).  A sub-sequence holds at most kSubBits symbols; a match that crosses a resolve-tile
// boundary is written twice (once per tile), at most once every 31 tokens.
constexpr int kTokRows = kSubBits + kSubBits / 16 + 16;
PP_HD uint32_t tok_words_for(int T) { return (uint32_t)kTokRows * (uint32_t)T; }
PP_HD size_t scratch_words_for(int T) { return (size_t)tok_words_for(T); }

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

enum : uint32_t { F_NONE = 0, F_EOB = 1, F_BAD = 2 };

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
constexpr uint64_t kLeadInPlace = ~0ull;
struct ChunkResult {
    int32_t status;     // 0 or negative ZResult
    uint32_t produced;  // bytes written (Core.cs:191)
    uint32_t newlines;  // '\n' bytes among them (by-product for the parse stage)
    uint32_t min_byte;  // 0 when a NUL byte was written (=> exact parser), else 1
    uint64_t end_bit;   // bit position after the last consumed bit
};

// Shared-memory carve-up (pointers into dynamic shared memory / emulation arrays).
struct Sm {
    uint32_t *cw;       // staged compressed words                         [cw_words_for(T)]
    uint32_t *lit;      // literal/length lookup table                     [kLitCap]
    uint32_t *dist;     // distance lookup table                           [kDistCap]
    uint32_t *start;    // per thread: first bit of its segment (window relative)   [T]
    uint32_t *end;      // per thread: bit after its last symbol                    [T]
    uint32_t *outc;     // per thread: bytes its segment produces; then exclusive sums [T]
    uint32_t *flag;     // per thread: F_*                                          [T]
    uint32_t *nl;       // per thread: '\n' bytes stored                            [T]
    uint32_t *nul;      // per thread: non-zero when a NUL byte was stored          [T]
    uint32_t *ns;       // per thread: its predecessor's end, latched for a SYNC round [T]
    uint32_t *ntok;     // per thread: tokens its segment emitted                     [T]
    uint16_t *res;      // resolve tile: 0x8000|byte or tile-relative source index  [kTileB*T]
    uint16_t *sorted;   // symbols in canonical order                               [320]
    uint16_t *codes;    // canonical code of every symbol                           [320]
    uint8_t *lens;      // code lengths: lit/len at 0, distance at 288              [320]
    uint32_t *count;    // [16] symbols per code length
    uint32_t *first;    // [16] first code of each length (MSB-first value)
    uint32_t *offs;     // [16] index of each length's first symbol in `sorted`
    uint32_t *u;        // [32] block-uniform scratch (broadcast slots)
    uint32_t *wsum;     // [32] per-warp partial sums
#ifndef PP_HOST_EMU
    unsigned long long *bar;  // one mbarrier for the staging copies
#endif
};

PP_HD uint32_t sm_bytes_for(int T)
{
    uint32_t b = 0;
    b += cw_words_for(T) * 4u;
    b += (kLitCap + kDistCap) * 4u;
    b += (uint32_t)T * 4u * 8u;
    b += ((uint32_t)T * kTileB + (uint32_t)T * kTileB / 32u + 2u + 7u) / 8u * 16u;
    b += 320u * 2u * 2u + 320u;
    b += (16u * 3u + 32u * 2u) * 4u;
    b += 16u;  // mbarrier
    return (b + 127u) & ~127u;
}
// `raw` must be 128-byte aligned (TMA destinations need 16)
PP_DEV void sm_carve(Sm &s, uint8_t *raw, int T)
{
    uint8_t *p = raw;
    s.cw = (uint32_t *)p; p += cw_words_for(T) * 4u;
    s.lit = (uint32_t *)p; p += kLitCap * 4u;
    s.dist = (uint32_t *)p; p += kDistCap * 4u;
    s.start = (uint32_t *)p; p += (uint32_t)T * 4u;
    s.end = (uint32_t *)p; p += (uint32_t)T * 4u;
    s.outc = (uint32_t *)p; p += (uint32_t)T * 4u;
    s.flag = (uint32_t *)p; p += (uint32_t)T * 4u;
    s.nl = (uint32_t *)p; p += (uint32_t)T * 4u;
    s.nul = (uint32_t *)p; p += (uint32_t)T * 4u;
    s.ns = (uint32_t *)p; p += (uint32_t)T * 4u;
    s.ntok = (uint32_t *)p; p += (uint32_t)T * 4u;
    s.res = (uint16_t *)p; p += ((uint32_t)T * kTileB + (uint32_t)T * kTileB / 32u + 2u + 7u) / 8u * 16u;
    s.count = (uint32_t *)p; p += 16u * 4u;
    s.first = (uint32_t *)p; p += 16u * 4u;
    s.offs = (uint32_t *)p; p += 16u * 4u;
    s.u = (uint32_t *)p; p += 32u * 4u;
    s.wsum = (uint32_t *)p; p += 32u * 4u;
#ifndef PP_HOST_EMU
    s.bar = (unsigned long long *)p;
#endif
    p += 16u;
    s.sorted = (uint16_t *)p; p += 320u * 2u;
    s.codes = (uint16_t *)p; p += 320u * 2u;
    s.lens = p;
}

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
    return n ? __brev(v) >> (32 - n) : 0u;
#endif
}
PP_DEV uint32_t popc32(uint32_t v)
{
#ifdef PP_HOST_EMU
    return (uint32_t)__builtin_popcount(v);
#else
    return (uint32_t)__popc(v);
#endif
}
// 0x80 in every byte of w that equals the byte replicated in c4
PP_DEV uint32_t eq_bytes(uint32_t w, uint32_t c4)
{
    const uint32_t x = w ^ c4;
    return ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x | 0x7f7f7f7fu);
}
// n bits (n <= 25) at window-relative bit position pos
PP_DEV uint32_t peek_bits(const uint32_t *cw, uint32_t pos, uint32_t n)
{
    const uint32_t w = pos >> 5, s = pos & 31u;
    const uint64_t v = (uint64_t)cw[w] | ((uint64_t)cw[w + 1] << 32);
    return (uint32_t)(v >> s) & ((1u << n) - 1u);
}

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
PP_DEV void tma_load(void *dst_smem, const void *src_gmem, uint32_t bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
#endif

// ---- STAGE -------------------------------------------------------------------
// Bring compressed bytes [byte_off, byte_off + 4*words) into sm.cw (zero filled past the
// end of the buffer).  byte_off is a multiple of 16.  Returns false when a transfer never
// landed.  `phase` counts the uses of the mbarrier.
PP_DEV bool stage_window(const Sm &sm, const uint8_t *comp, uint64_t comp_bytes, uint64_t byte_off, uint32_t words,
                         uint32_t &phase)
{
    // comp_bytes is the TRUE readable extent of the buffer (it may be a caller's pinned host
    // buffer with nothing behind it): whole 16-byte units go through TMA, a ragged tail of
    // fewer than 16 bytes is fetched with byte loads, everything past the end reads as zero
    const uint64_t avail = byte_off < comp_bytes ? comp_bytes - byte_off : 0;
    const uint32_t want = words * 4u;
    const uint32_t have = avail < want ? (uint32_t)avail : want;   // bytes that exist
    const uint32_t nbytes = have & ~15u;                            // bulk-copied part
    const uint32_t tail = have - nbytes;                            // < 16
    PP_SYNC();  // every thread is done with the previous contents of cw
#ifdef PP_HOST_EMU
    if (have) memcpy(sm.cw, comp + byte_off, have);
    memset((uint8_t *)sm.cw + have, 0, want - have);
    (void)phase;
    (void)tail;
    return true;
#else
    if (threadIdx.x == 0 && nbytes) {
        mbar_expect_tx(sm.bar, nbytes);
        for (uint32_t o = 0; o < nbytes; o += 32768u) {
            const uint32_t n = nbytes - o < 32768u ? nbytes - o : 32768u;
            tma_load((uint8_t *)sm.cw + o, comp + byte_off + o, n, sm.bar);
        }
    }
    if (nbytes < want) {
        // the 16-byte unit holding the ragged tail (threads 0..15, one byte each), zeros behind it
        if (threadIdx.x < 16u)
            ((uint8_t *)sm.cw)[nbytes + threadIdx.x] = threadIdx.x < tail ? comp[byte_off + nbytes + threadIdx.x] : (uint8_t)0;
        for (uint32_t i = nbytes / 4u + 4u + threadIdx.x; i < words; i += blockDim.x) sm.cw[i] = 0;
    }
    bool ok = true;
    if (nbytes) {
        ok = mbar_wait(sm.bar, phase & 1u);
        phase++;
    }
    return __syncthreads_and(ok ? 1 : 0) != 0;
#endif
}

// Pull mode (the compressed input is pinned HOST memory, every staged byte crosses PCIe and the link is the limit:
// 51.5 GB/s to SM reads, profiles/src/pull_probe.cu).  A window is staged whole (~63 KB) but usually ends with its
// block (~57 KB), and the next window starts where it ended: the last few KB of what is staged are the first few
// KB of the next window.  RESOLVE's scratch overwrites only the front of cw (resolve_scratch_bytes), so when the
// next window starts behind that, its head is moved down inside shared memory and only the rest is fetched:
// `delta` = new base - old base (multiple of 16, >= resolve_scratch_bytes(T), < 4 * words).
PP_DEV bool stage_window_reuse(const Sm &sm, const uint8_t *comp, uint64_t comp_bytes, uint64_t byte_off, uint32_t words,
                               uint32_t &phase, uint32_t delta)
{
    const int T = PP_NT;
    const uint32_t want = words * 4u, keep = want - delta;  // keep < delta: source and destination do not overlap
    PP_SYNC();  // every thread is done with the previous contents of cw
    PP_FOR_T(t)
    for (uint32_t i = (uint32_t)t; i < keep / 16u; i += (uint32_t)T)
        reinterpret_cast<uint4 *>(sm.cw)[i] = reinterpret_cast<const uint4 *>(sm.cw)[delta / 16u + i];
    PP_END_T
    PP_SYNC();  // the old bytes are read before the new ones land on them
    const uint64_t off = byte_off + keep;
    const uint64_t avail = off < comp_bytes ? comp_bytes - off : 0;
    const uint32_t need = delta;                                    // bytes [keep, want) of the window
    const uint32_t have = avail < need ? (uint32_t)avail : need;
    const uint32_t nbytes = have & ~15u;
    const uint32_t tail = have - nbytes;
#ifdef PP_HOST_EMU
    if (have) memcpy((uint8_t *)sm.cw + keep, comp + off, have);
    memset((uint8_t *)sm.cw + keep + have, 0, need - have);
    (void)phase;
    (void)tail;
    return true;
#else
    if (threadIdx.x == 0 && nbytes) {
        mbar_expect_tx(sm.bar, nbytes);
        for (uint32_t o = 0; o < nbytes; o += 32768u) {
            const uint32_t n = nbytes - o < 32768u ? nbytes - o : 32768u;
            tma_load((uint8_t *)sm.cw + keep + o, comp + off + o, n, sm.bar);
        }
    }
    if (nbytes < need) {
        if (threadIdx.x < 16u)
            ((uint8_t *)sm.cw)[keep + nbytes + threadIdx.x] = threadIdx.x < tail ? comp[off + nbytes + threadIdx.x] : (uint8_t)0;
        for (uint32_t i = (keep + nbytes) / 4u + 4u + threadIdx.x; i < words; i += blockDim.x) sm.cw[i] = 0;
    }
    bool ok = true;
    if (nbytes) {
        ok = mbar_wait(sm.bar, phase & 1u);
        phase++;
    }
    return __syncthreads_and(ok ? 1 : 0) != 0;
#endif
}

// ---- block-wide exclusive prefix sum over a[0..T) (in place); returns the total ----
PP_DEV uint32_t block_excl_scan(const Sm &sm, uint32_t *a)
{
#ifdef PP_HOST_EMU
    uint32_t acc = 0;
    for (int t = 0; t < PP_NT; t++) {
        const uint32_t v = a[t];
        a[t] = acc;
        acc += v;
    }
    return acc;
#else
    const int t = (int)threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint32_t v = a[t];
    uint32_t s = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t x = __shfl_up_sync(0xffffffffu, s, d);
        if (lane >= d) s += x;
    }
    if (lane == 31) sm.wsum[warp] = s;
    __syncthreads();
    if (warp == 0) {
        const int nw = ((int)blockDim.x + 31) >> 5;
        const uint32_t w = lane < nw ? sm.wsum[lane] : 0u;
        uint32_t ws = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t x = __shfl_up_sync(0xffffffffu, ws, d);
            if (lane >= d) ws += x;
        }
        sm.wsum[lane] = ws - w;          // exclusive warp offsets
        if (lane == 31) sm.u[31] = ws;   // total
    }
    __syncthreads();
    a[t] = sm.wsum[warp] + s - v;
    const uint32_t total = sm.u[31];
    __syncthreads();
    return total;
#endif
}

// ---- HEADER: Huffman table construction -----------------------------------------
// Canonical codes (RFC 1951 3.2.2) into a two-level lookup table: `root` primary bits,
// sub-tables sized by the longest code under each primary prefix.  Validity rules follow
// zlib's inflate_table: over-subscribed sets are rejected, incomplete sets are rejected
// unless the set is a single 1-bit code (or, for distances, empty).
// mode 0: literal/length alphabet, 1: distance alphabet.  Returns 0 or -3 (Z_DATA_ERROR).
PP_DEV uint32_t sym_entry(int mode, uint32_t s, uint32_t l)
{
    if (mode == 0) {
        if (s < 256u) return mk_entry(K_LIT, l, l, 0x8000u | s);
        if (s == 256u) return mk_entry(K_EOB, l, l, 0);
        if (s < 286u) return mk_entry(K_BASE, l + kLenExtra[s - 257u], l, kLenBase[s - 257u]);
        return mk_entry(K_BAD, l, l, 0);
    }
    if (s < 30u) return mk_entry(K_BASE, l + kDistExtra[s], l, kDistBase[s]);
    return mk_entry(K_BAD, l, l, 0);
}

PP_DEV int build_table(const Sm &sm, uint32_t *tbl, int root, int cap, int nsym, int lens_off, int mode)
{
    const uint8_t *lens = sm.lens + lens_off;
    const int T = PP_NT;
    // 1. histogram of code lengths
    PP_FOR_T(t)
    if (t < 16) sm.count[t] = 0;
    PP_END_T
    PP_SYNC();
    PP_FOR_T(t)
    for (int s = t; s < nsym; s += T) PP_ATOMIC_ADD(&sm.count[lens[s]], 1u);
    PP_END_T
    PP_SYNC();
    // 2. validity, first code and first sorted slot of each length (thread 0)
    PP_T0_BEGIN
    {
        int maxlen = 0, left = 1, err = 0;
        for (int l = 1; l <= 15; l++) {
            if (sm.count[l]) maxlen = l;
            left <<= 1;
            left -= (int)sm.count[l];
            if (left < 0) err = 1;  // over-subscribed
        }
        if (left > 0 && maxlen != 1 && !(mode == 1 && maxlen == 0)) err = 1;  // incomplete set
        uint32_t code = 0, off = 0;
        sm.first[0] = 0;
        sm.offs[0] = 0;
        for (int l = 1; l <= 15; l++) {
            code = (code + (l > 1 ? sm.count[l - 1] : 0u)) << 1;
            sm.first[l] = code;
            sm.offs[l] = off;
            off += sm.count[l];
        }
        sm.u[0] = (uint32_t)err;
        sm.u[1] = (uint32_t)maxlen;
    }
    PP_T0_END
    PP_SYNC();
    if (sm.u[0]) { PP_SYNC(); return -3; }
    const int maxlen = (int)sm.u[1];
    // 3. canonical code of every symbol: first[len] + (symbols of the same length before it)
    PP_FOR_T(t)
    for (int s = t; s < nsym; s += T) {
        const uint32_t l = lens[s];
        if (l == 0) continue;
        uint32_t rank = 0;
        const uint32_t l4 = l * 0x01010101u;
        const uint32_t *w = reinterpret_cast<const uint32_t *>(lens);  // lens_off is a multiple of 4
        const int full = s >> 2;
        for (int i = 0; i < full; i++) rank += popc32(eq_bytes(w[i], l4));
        for (int i = full * 4; i < s; i++) rank += (lens[i] == l);
        sm.codes[lens_off + s] = (uint16_t)(sm.first[l] + rank);
        sm.sorted[sm.offs[l] + rank] = (uint16_t)s;
    }
    PP_END_T
    PP_SYNC();
    // 4. primary table: every slot finds its own code by canonical search
    const int nprim = 1 << root;
    PP_FOR_T(t)
    for (int i = t; i < nprim; i += T) {
        const uint32_t c = bitrev((uint32_t)i, root);  // the slot's bits as an MSB-first code prefix
        uint32_t ent = mk_entry(K_BAD, 1, 1, 0);
        const int lim = maxlen < root ? maxlen : root;
        for (int l = 1; l <= lim; l++) {
            const uint32_t cc = (c >> (root - l)) - sm.first[l];
            if (cc < sm.count[l]) {
                ent = sym_entry(mode, sm.sorted[sm.offs[l] + cc], (uint32_t)l);
                break;
            }
        }
        tbl[i] = ent;
    }
    PP_END_T
    PP_SYNC();
    if (maxlen <= root) return 0;
    // 5. sub-table geometry: prefix p owns the length-l codes in [p << (l-root), (p+1) << (l-root));
    //    long codes sit at the top of the code space.  Every prefix finds its own sub-table size in
    //    parallel (scratch: the resolve tile buffer, idle here); thread 0 only hands out the offsets.
    int pmin = nprim;
    for (int l = root + 1; l <= maxlen; l++)
        if (sm.count[l]) { pmin = (int)(sm.first[l] >> (l - root)); break; }
    uint8_t *subs = reinterpret_cast<uint8_t *>(sm.res);  // [nprim] sub-table index bits per prefix (0: none)
    PP_FOR_T(t)
    for (int p = pmin + t; p < nprim; p += T) {
        int sub = 0;
        for (int l = maxlen; l > root; l--) {
            const uint32_t lo = (uint32_t)p << (l - root), hi = ((uint32_t)p + 1u) << (l - root);
            const uint32_t f = sm.first[l], e = f + sm.count[l];
            if (sm.count[l] && lo < e && hi > f) { sub = l - root; break; }
        }
        subs[p] = (uint8_t)sub;
    }
    PP_END_T
    PP_SYNC();
    PP_T0_BEGIN
    {
        int used = nprim, err = 0;
        for (int p = pmin; p < nprim; p++) {
            const int sub = subs[p];
            if (sub) {
                if (used + (1 << sub) > cap) { err = 1; break; }
                tbl[bitrev((uint32_t)p, root)] =
                    mk_entry(K_SUB, (uint32_t)root, (uint32_t)root, (uint32_t)used) | ((uint32_t)sub << 13);
                used += 1 << sub;
            }
        }
        sm.u[0] = (uint32_t)err;
        sm.u[2] = (uint32_t)used;
    }
    PP_T0_END
    PP_SYNC();
    if (sm.u[0]) { PP_SYNC(); return -3; }
    const int used = (int)sm.u[2];
    PP_FOR_T(t)
    for (int i = nprim + t; i < used; i += T) tbl[i] = mk_entry(K_BAD, 1, 1, 0);
    PP_END_T
    PP_SYNC();
    // 6. long codes into their sub-tables (one symbol per thread)
    PP_FOR_T(t)
    for (int s = t; s < nsym; s += T) {
        const int l = lens[s];
        if (l <= root) continue;
        const uint32_t code = sm.codes[lens_off + s];
        const uint32_t pe = tbl[bitrev(code >> (l - root), root)];
        const int sub = (int)e_sub(pe);
        const uint32_t st = e_val(pe);
        const int sl = l - root;  // bits of this code inside the sub-table
        const uint32_t base = bitrev(code & ((1u << sl) - 1u), sl);
        const uint32_t ent = sym_entry(mode, (uint32_t)s, (uint32_t)l);
        for (int j = 0; j < (1 << (sub - sl)); j++) tbl[st + (base | ((uint32_t)j << sl))] = ent;
    }
    PP_END_T
    PP_SYNC();
    return 0;
}

PP_DEV int fixed_tables(const Sm &sm)
{
    PP_FOR_T(t)
    for (int s = t; s < 320; s += PP_NT)
        sm.lens[s] = (uint8_t)(s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : s < 288 ? 8 : 5);
    PP_END_T
    PP_SYNC();
    int rc = build_table(sm, sm.lit, kRootL, kLitCap, 288, 0, 0);
    if (rc) return rc;
    // zlib's fixed distance table is the 5-bit complete code over 32 symbols (30/31 invalid)
    return build_table(sm, sm.dist, kRootD, kDistCap, 32, 288, 1);
}

// Dynamic block header at window-relative bit `pos` (just past the 3 block-type bits):
// HLIT/HDIST/HCLEN, the code-length code, the run-length coded lengths, then both tables.
// On success *pos_out is the first symbol's bit.
//
// The run-length coded lengths are a serial bit stream (every symbol's position depends on
// the one before), and one GPU thread needs ~140 cycles per symbol for "refill, look up,
// shift, store".  So the lookups are done SPECULATIVELY for every bit position by all threads
// (node[p] = bits consumed | repeat count | symbol of the symbol that would start at p), and
// thread 0 only follows the chain p -> p + bits(node[p]): one dependent shared-memory load per
// symbol.  It records (first index, repeat, value) per symbol; all threads then expand the runs.
// Scratch: sm.lit (the table it is about to build), sm.dist[0..127], sm.wsum.
PP_DEV int dynamic_tables(const Sm &sm, uint32_t pos, uint32_t *pos_out)
{
    uint32_t *const chain = sm.lit;        // [320] symbols on the true path: first index | repeat << 9 | value << 17
    uint32_t *const node = sm.lit + 512;   // [T] one round of speculative decodes
    uint32_t *const clv = sm.wsum;         // [19] lengths of the code-length code
    const uint32_t nlen = peek_bits(sm.cw, pos, 5) + 257u;
    const uint32_t ndist = peek_bits(sm.cw, pos + 5, 5) + 1u;
    const uint32_t ncode = peek_bits(sm.cw, pos + 10, 4) + 4u;
    const uint32_t total = nlen + ndist;
    const uint32_t p0 = pos + 14u + 3u * ncode;
    PP_FOR_T(t)
    for (int i = t; i < 320; i += PP_NT) sm.lens[i] = 0;
    for (int i = t; i < 128; i += PP_NT) sm.dist[i] = 0;
    if (t < 19) clv[kClOrder[t]] = (uint32_t)t < ncode ? peek_bits(sm.cw, pos + 14u + 3u * (uint32_t)t, 3) : 0u;
    if (t == 0) {
        sm.u[0] = (nlen > 286u || ndist > 30u) ? 1u : 0u;  // too many length or distance symbols
        sm.u[3] = 0;   // lengths produced
        sm.u[4] = 0;   // chain entries
        sm.u[5] = p0;  // bit position of the walk
        sm.u[6] = 0;   // previous length (for symbol 16)
    }
    PP_END_T
    PP_SYNC();
    // code-length code: 7-bit direct lookup table in sm.dist: entry = sym | len << 8 | 0x8000, 0 = invalid.
    // One thread per symbol: its canonical code is first[len] + (symbols before it with the same length).
    PP_FOR_T(t)
    if (t < 19) {
        const uint32_t l = clv[t];
        uint64_t cnt = 0;  // eight 8-bit counters, one per length
        uint32_t rank = 0;
        for (int i = 0; i < 19; i++) {
            const uint32_t li = clv[i];
            cnt += 1ull << (8u * li);
            if (i < t && li == l) rank++;
        }
        int left = 1, bad = 0;
        uint32_t code = 0, prev = 0, mine = 0;
        for (uint32_t k = 1; k <= 7u; k++) {
            const uint32_t c = (uint32_t)(cnt >> (8u * k)) & 255u;
            left = (left << 1) - (int)c;
            if (left < 0) bad = 1;
            code = (code + prev) << 1;
            if (k == l) mine = code + rank;
            prev = c;
        }
        if (left > 0) bad = 1;  // zlib: the code-length code must be complete
        if (bad) {
            if (t == 0) sm.u[0] = 1;
        } else if (l) {
            for (uint32_t j = bitrev(mine, (int)l); j < 128u; j += 1u << l) sm.dist[j] = (uint32_t)t | (l << 8) | 0x8000u;
        }
    }
    PP_END_T
    PP_SYNC();
    if (sm.u[0]) { PP_SYNC(); return -3; }
    const uint32_t pmax = p0 + 4440u;  // 316 symbols of at most 14 bits: the true path never gets here
    for (uint32_t base = p0;; base += (uint32_t)PP_NT) {
        PP_FOR_T(t)
        {
            const uint32_t q = base + (uint32_t)t;
            uint32_t nd = 0;
            if (q < pmax) {
                const uint32_t bits = peek_bits(sm.cw, q, 14);
                const uint32_t e = sm.dist[bits & 127u];
                if (e) {
                    const uint32_t l = (e >> 8) & 15u, sym = e & 255u, x = bits >> l;
                    uint32_t n = l, rep = 1;
                    if (sym == 16u) { rep = 3u + (x & 3u); n += 2u; }
                    else if (sym == 17u) { rep = 3u + (x & 7u); n += 3u; }
                    else if (sym == 18u) { rep = 11u + (x & 127u); n += 7u; }
                    nd = n | (rep << 4) | (sym << 12) | 0x80000000u;
                }
            }
            node[t] = nd;
        }
        PP_END_T
        PP_SYNC();
        PP_T0_BEGIN
        {
            uint32_t q = sm.u[5], have = sm.u[3], k = sm.u[4], last = sm.u[6], err = 0;
            const uint32_t lim = base + (uint32_t)PP_NT;
            while (have < total && q < lim) {
                const uint32_t nd = node[q - base];
                if (!nd) { err = 1; break; }  // invalid code
                const uint32_t rep = (nd >> 4) & 255u, sym = (nd >> 12) & 31u;
                uint32_t val = 0;
                if (sym < 16u) {
                    val = sym;
                    last = sym;
                } else if (sym == 16u) {
                    if (have == 0) { err = 1; break; }  // invalid bit length repeat
                    val = last;
                } else {
                    last = 0;
                }
                if (have + rep > total) { err = 1; break; }  // invalid bit length repeat
                chain[k++] = have | (rep << 9) | (val << 17);
                have += rep;
                q += nd & 15u;
            }
            sm.u[5] = q;
            sm.u[3] = have;
            sm.u[4] = k;
            sm.u[6] = last;
            if (err) sm.u[0] = 1;
            sm.u[7] = (err || have >= total) ? 1u : 0u;
        }
        PP_T0_END
        PP_SYNC();
        if (sm.u[7]) break;
    }
    if (sm.u[0]) { PP_SYNC(); return -3; }
    const uint32_t nch = sm.u[4];
    // runs -> lens: lit/len lengths at 0, distance lengths (they follow in the stream) at 288
    PP_FOR_T(t)
    for (uint32_t k = (uint32_t)t; k < nch; k += (uint32_t)PP_NT) {
        const uint32_t c = chain[k], val = c >> 17;
        if (val) {
            const uint32_t h0 = c & 511u, h1 = h0 + ((c >> 9) & 255u);
            for (uint32_t h = h0; h < h1; h++) sm.lens[h < nlen ? h : h + 288u - nlen] = (uint8_t)val;
        }
    }
    PP_END_T
    PP_SYNC();
    if (sm.lens[256] == 0) { PP_SYNC(); return -3; }  // invalid code -- missing end-of-block
    *pos_out = sm.u[5];
    PP_SYNC();
    PP_PHASE(PH_H_PARSE);
    int rc = build_table(sm, sm.lit, kRootL, kLitCap, (int)nlen, 0, 0);
    PP_PHASE(PH_H_LIT);
    if (rc) return rc;
    return build_table(sm, sm.dist, kRootD, kDistCap, (int)ndist, 288, 1);
}

// ---- GUESS / SYNC / EMIT: one thread walks one segment ---------------------------------
// EMIT decodes the symbols that START in [start, limit) (window-relative bits) and writes one
// 32-bit TOKEN per symbol, token k of this thread at tok[t * kTokRows + k]:
//   [15:0]  entry: what the resolve stage starts from for every byte of the token —
//           0x8000 | byte for a literal, distance - 1 for a match;
//   [29:16] position of the token's first byte inside its RESOLVE TILE (tiles are R = 16 T
//           output bytes, cut in the window's virtual index space; R <= 16384);
//   [31:30] tile number modulo 4 (a row's tokens are in output order, so "the tokens of tile i"
//           is a run of equal tags; a token is at most 258 bytes, so no tag is skipped).
// A match that crosses a tile boundary is written a second time as a token that starts at the
// boundary: bytes repeat with period `distance` inside a match, so the continuation is an
// ordinary match of the same distance.  The token's length is not stored: the resolve stage
// sees where the next token starts.  `o` is the virtual output index of the segment's first
// byte (window output offset + misalignment of the window's first byte).
struct Seg {
    uint32_t end, out, flag, ntok;
};
PP_DEV uint32_t tok_pack(uint32_t entry, uint32_t p0, uint32_t rshift)
{
    return entry | ((p0 & ((1u << rshift) - 1u)) << 16) | ((p0 >> rshift) << 30);
}

template <int WRITE>
PP_DEV Seg decode_seg(const Sm &sm, uint32_t start, uint32_t limit, uint32_t *tok, uint32_t rshift, uint32_t t,
                      uint32_t o)
{
    const uint32_t *cw = sm.cw;
    uint32_t *row = tok + t * (uint32_t)kTokRows;
    uint32_t wp = start >> 5;
    const uint32_t sh = start & 31u;
    uint64_t buf = ((uint64_t)cw[wp] | ((uint64_t)cw[wp + 1] << 32)) >> sh;
    uint32_t cnt = 64u - sh;
    wp += 2;
    uint32_t out = 0, flag = F_NONE, k = 0;
    // One loop body for literals and matches (the lanes of a warp are in different kinds of symbol
    // all the time: a branch per kind makes every iteration pay for both): the distance lookup is
    // done for a literal too, on whatever bits follow, and consumes nothing.
    for (;;) {
        if (wp * 32u - cnt >= limit) break;
        if (cnt < 32u) { buf |= (uint64_t)cw[wp] << cnt; cnt += 32u; wp++; }
        uint32_t lo = (uint32_t)buf;
        uint32_t e = sm.lit[lo & ((1u << kRootL) - 1u)];
        if (e_kind(e) == K_SUB) e = sm.lit[e_val(e) + ((lo >> kRootL) & ((1u << e_sub(e)) - 1u))];
        const uint32_t kind = e_kind(e), tot = e_tot(e);
        if (kind > K_BASE) {  // end of block, or an invalid literal/length code
            if (kind == K_EOB) { buf >>= tot; cnt -= tot; flag = F_EOB; }
            else flag = F_BAD;
            break;
        }
        const bool ism = kind == K_BASE;
        const uint32_t len = e_val(e) + ((lo & ~(0xffffffffu << tot)) >> e_cl(e));  // (a literal: its entry, unused)
        buf >>= tot;
        cnt -= tot;
        if (cnt < 32u) { buf |= (uint64_t)cw[wp] << cnt; cnt += 32u; wp++; }
        lo = (uint32_t)buf;
        uint32_t d = sm.dist[lo & ((1u << kRootD) - 1u)];
        if (ism && e_kind(d) == K_SUB) d = sm.dist[e_val(d) + ((lo >> kRootD) & ((1u << e_sub(d)) - 1u))];
        if (ism && e_kind(d) != K_BASE) { flag = F_BAD; break; }  // invalid distance code
        const uint32_t dtot = ism ? e_tot(d) : 0u;
        // dist <= 32768 <= lead_len always, so "distance too far back" cannot occur:
        // the reference primes a full 32 KB dictionary (Core.cs:158)
        const uint32_t dist = e_val(d) + ((lo & ~(0xffffffffu << dtot)) >> e_cl(d));
        buf >>= dtot;
        cnt -= dtot;
        const uint32_t n = ism ? len : 1u;
        if (WRITE) {
            const uint32_t p0 = o + out, p1 = p0 + n - 1u;
            row[k++] = tok_pack(ism ? dist - 1u : e_val(e), p0, rshift);
            if ((p0 ^ p1) >> rshift) row[k++] = tok_pack(dist - 1u, (p1 >> rshift) << rshift, rshift);  // only a match crosses
        }
        out += n;
    }
    Seg r;
    r.end = wp * 32u - cnt;
    r.out = out;
    r.flag = flag;
    r.ntok = k;
    return r;
}

// GUESS / SYNC walk of one segment with CHECKPOINTS.  Slot c (c = 0..kCkpt-1) of a thread is the
// first symbol boundary at or after bit tgt0 + c * kCkptStep of its sub-sequence, stored as one
// word: position relative to `base` (the sub-sequence's first bit; < 2048) | output bytes from
// there to the end of the segment << 11 (a segment emits < 128 Ki bytes).  A decoder that restarts
// at a corrected bit falls back onto the old path's boundaries after a few symbols; as soon as it
// stands exactly on the old path's boundary for the slot it is passing, the rest of the segment is
// already known: it stops there and reuses the stored suffix (end, flag and later slots are
// unchanged), so a SYNC round costs a fraction of a full pass.  cp: [kCkpt][T] words (aliases the
// resolve tile buffer, which is idle while decoding).
constexpr int kCkpt = 7;
constexpr uint32_t kCkptStep = (uint32_t)kSubBits / (kCkpt + 1);
constexpr uint32_t kNoCkpt = 0xffffffffu;

PP_DEV Seg decode_count(const Sm &sm, uint32_t start, uint32_t limit, uint32_t base, uint32_t *cp, uint32_t T,
                        uint32_t t, bool have_old, uint32_t old_end, uint32_t old_flag)
{
    const uint32_t *cw = sm.cw;
    uint32_t wp = start >> 5;
    const uint32_t sh = start & 31u;
    uint64_t buf = ((uint64_t)cw[wp] | ((uint64_t)cw[wp + 1] << 32)) >> sh;
    uint32_t cnt = 64u - sh;
    wp += 2;
    uint32_t out = 0, flag = F_NONE;
    uint32_t c = 0, next_t = base + kCkptStep;
    bool reused = false;
    for (;;) {
        const uint32_t pos = wp * 32u - cnt;
        if (pos >= limit) break;
        if (pos >= next_t && c < (uint32_t)kCkpt) {
            const uint32_t old = cp[c * T + t];
            if (have_old && old != kNoCkpt && (old & 2047u) == pos - base) {  // on the old path: the rest is known
                out += old >> 11;
                reused = true;
                break;
            }
            cp[c * T + t] = (pos - base) | (out << 11);  // bytes BEFORE the boundary for now; turned into the suffix below
            c++;
            next_t += kCkptStep;
        }
        if (cnt < 32u) { buf |= (uint64_t)cw[wp] << cnt; cnt += 32u; wp++; }
        uint32_t lo = (uint32_t)buf;
        uint32_t e = sm.lit[lo & ((1u << kRootL) - 1u)];
        if (e_kind(e) == K_SUB) e = sm.lit[e_val(e) + ((lo >> kRootL) & ((1u << e_sub(e)) - 1u))];
        const uint32_t kind = e_kind(e), tot = e_tot(e);
        if (kind > K_BASE) {  // end of block, or an invalid literal/length code
            if (kind == K_EOB) { buf >>= tot; cnt -= tot; flag = F_EOB; }
            else flag = F_BAD;
            break;
        }
        // one body for literals and matches (see decode_seg)
        const bool ism = kind == K_BASE;
        const uint32_t len = e_val(e) + ((lo & ~(0xffffffffu << tot)) >> e_cl(e));
        buf >>= tot;
        cnt -= tot;
        if (cnt < 32u) { buf |= (uint64_t)cw[wp] << cnt; cnt += 32u; wp++; }
        lo = (uint32_t)buf;
        uint32_t d = sm.dist[lo & ((1u << kRootD) - 1u)];
        if (ism && e_kind(d) == K_SUB) d = sm.dist[e_val(d) + ((lo >> kRootD) & ((1u << e_sub(d)) - 1u))];
        if (ism && e_kind(d) != K_BASE) { flag = F_BAD; break; }  // invalid distance code
        const uint32_t dtot = ism ? e_tot(d) : 0u;
        buf >>= dtot;
        cnt -= dtot;
        out += ism ? len : 1u;
    }
    Seg r;
    if (reused) {
        r.end = old_end;
        r.flag = old_flag;
    } else {
        r.end = wp * 32u - cnt;
        r.flag = flag;
        for (uint32_t j = c; j < (uint32_t)kCkpt; j++) cp[j * T + t] = kNoCkpt;  // slots the new path never reached
    }
    r.out = out;
    r.ntok = 0;
    // slots passed on the way (before the reuse point, or all of them) now describe the new path
    for (uint32_t j = 0; j < c; j++) {
        const uint32_t w = cp[j * T + t];
        cp[j * T + t] = (w & 2047u) | ((out - (w >> 11)) << 11);
    }
    return r;
}

// Bit position right after the symbol whose output makes a segment's byte count reach `target`
// (target >= 1); one thread, once per chunk: zlib stops as soon as the wanted bytes are out
// (Core.cs:187), so the input it needed ends there and not at the end of the window.
PP_DEV uint32_t seg_pos_at_output(const Sm &sm, uint32_t start, uint32_t limit, uint32_t target)
{
    const uint32_t *cw = sm.cw;
    uint32_t wp = start >> 5;
    const uint32_t sh = start & 31u;
    uint64_t buf = ((uint64_t)cw[wp] | ((uint64_t)cw[wp + 1] << 32)) >> sh;
    uint32_t cnt = 64u - sh;
    wp += 2;
    uint32_t out = 0;
    while (wp * 32u - cnt < limit && out < target) {
        if (cnt < 32u) { buf |= (uint64_t)cw[wp] << cnt; cnt += 32u; wp++; }
        uint32_t lo = (uint32_t)buf;
        uint32_t e = sm.lit[lo & ((1u << kRootL) - 1u)];
        if (e_kind(e) == K_SUB) e = sm.lit[e_val(e) + ((lo >> kRootL) & ((1u << e_sub(e)) - 1u))];
        const uint32_t kind = e_kind(e), tot = e_tot(e);
        if (kind != K_LIT && kind != K_BASE) break;  // EOB / invalid: the caller has already seen it
        uint32_t n = 1;
        if (kind == K_BASE) n = e_val(e) + ((lo & ~(0xffffffffu << tot)) >> e_cl(e));
        buf >>= tot;
        cnt -= tot;
        if (kind == K_BASE) {
            if (cnt < 32u) { buf |= (uint64_t)cw[wp] << cnt; cnt += 32u; wp++; }
            lo = (uint32_t)buf;
            uint32_t d = sm.dist[lo & ((1u << kRootD) - 1u)];
            if (e_kind(d) == K_SUB) d = sm.dist[e_val(d) + ((lo >> kRootD) & ((1u << e_sub(d)) - 1u))];
            const uint32_t dtot = e_tot(d);
            buf >>= dtot;
            cnt -= dtot;
        }
        out += n;
    }
    return wp * 32u - cnt;
}

struct WindowOut {
    uint32_t next_bit;   // window-relative bit after the last symbol used
    uint32_t need_bit;   // window-relative bit up to which the compressed input was really needed
    uint32_t produced;   // output bytes (clipped to the room left)
    uint32_t flag;       // F_EOB: the block ended; F_BAD: invalid data; F_NONE: window or output exhausted
    uint32_t rounds;     // SYNC rounds (statistics)
};

// ---- RESOLVE ------------------------------------------------------------------------
// Tokens -> bytes for window output [0, total) that lands at outp[0..total).
// `a` = misalignment of outp (outp - a is 16-byte aligned); virtual index v = q + a.
// The output is walked in tiles of R = 16 bytes per thread; per tile:
//  SCATTER  (one tile ahead, overlapped with CHASE of the previous tile) the tile's tokens are
//           brought from the rows in global memory into shared memory BY POSITION: a warp takes
//           a row, its lanes take consecutive tokens (coalesced), each writes its 16-bit entry
//           at ent[position] and sets bit `position` of the tile's HEAD MASK.  Per row a cursor
//           remembers how far earlier tiles got.
//  EXPAND   lanes own INTERLEAVED bytes (byte j*32+lane of the warp's 512): the token a byte
//           belongs to is the last head at or before it — one mask word, one count-leading-
//           zeros — and its offset inside the token falls out of the same subtraction.  A
//           literal is final; a match byte whose source lies in front of the tile is final in
//           global memory (the 32 lanes of a load read one or two runs of consecutive bytes, a
//           few sectors) and is fetched right here; a source inside the tile becomes the
//           tile-relative index of that byte.  Overlapping matches (distance < offset) are
//           re-pointed at the same byte one or more periods earlier, in front of the match.
//  CHASE    in-tile sources are followed through shared memory with no barrier: an entry is
//           always either the byte or the index of an EARLIER byte with the same value and
//           every hop is published, so chains collapse like pointer jumping whatever the
//           interleaving of the warps.
//  STORE    thread t packs the sixteen final entries of group t and writes one 16-byte vector;
//           '\n' and NUL are counted on the way for the parse stage.
PP_DEV uint32_t div_small(uint32_t i, uint32_t d)  // floor(i / d) for i, d < 512
{
#ifdef PP_HOST_EMU
    return i / d;
#else
    return (uint32_t)__fdividef((float)i + 0.5f, (float)d);  // (i+0.5)/d is >= 0.5/d away from an integer
#endif
}
PP_DEV uint32_t clz32(uint32_t v)
{
#ifdef PP_HOST_EMU
    return v ? (uint32_t)__builtin_clz(v) : 32u;
#else
    return (uint32_t)__clz((int)v);
#endif
}
// Shared-memory accesses of the EXPAND step by 32-bit shared address (device) / plain pointer
// (emulation): keeps the per-byte address arithmetic to one add.
#ifdef PP_HOST_EMU
typedef uint8_t *SAddr;
PP_DEV SAddr saddr(const volatile void *p) { return (uint8_t *)p; }
PP_DEV uint32_t lds_u32(SAddr a, uint32_t off) { uint32_t v; memcpy(&v, a + off, 4); return v; }
PP_DEV uint32_t lds_u16(SAddr a, uint32_t off) { uint16_t v; memcpy(&v, a + off, 2); return v; }
PP_DEV void sts_u16(SAddr a, uint32_t off, uint32_t v) { const uint16_t x = (uint16_t)v; memcpy(a + off, &x, 2); }
PP_DEV void sts_u32(SAddr a, uint32_t off, uint32_t v) { memcpy(a + off, &v, 4); }
#else
typedef uint32_t SAddr;
PP_DEV SAddr saddr(const volatile void *p) { return (uint32_t)__cvta_generic_to_shared(const_cast<const void *>(p)); }
PP_DEV uint32_t lds_u32(SAddr a, uint32_t off)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a + off) : "memory");
    return v;
}
PP_DEV uint32_t lds_u16(SAddr a, uint32_t off)
{
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a + off) : "memory");
    return v;
}
PP_DEV void sts_u16(SAddr a, uint32_t off, uint32_t v)
{
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(a + off), "h"((uint16_t)v) : "memory");
}
PP_DEV void sts_u32(SAddr a, uint32_t off, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a + off), "r"(v) : "memory");
}
#endif
PP_DEV uint32_t res_pos(uint32_t q) { return q; }
PP_HD uint32_t res_entries_for(int T) { return (uint32_t)T * kTileB + (uint32_t)T * kTileB / 32u + 2u; }

// Shared-memory staging of a tile's tokens (lives in the compressed-window buffer, idle during
// RESOLVE): two buffers (the tile being expanded, the tile being scattered), each R u16 entries
// indexed by tile position + R/32 head-mask words.
static_assert(kSubW * 4 >= 2 * 2 * kTileB + 2 * kTileB / 8 + kTileB / 8 + 32,
              "the tile token staging (entries, head masks, word heads, row map: 102 bytes per thread) must fit the compressed-window buffer");
struct TileTok {
    uint16_t *ent;     // [2][R]
    uint32_t *mask;    // [2][R/32]
    uint32_t *kcur;    // [T] per row: tokens consumed by earlier tiles
    uint32_t *wprev;   // [R/32] per mask word: position of the last head BEFORE the word (EXPAND scratch)
    uint16_t *tfirst;  // [max_tiles_for(T)+2] per tile: first row that starts in it or later (the window's row map)
};
// tiles one window may produce before it is cut short (its row map must fit the staging buffer)
PP_HD uint32_t max_tiles_for(int T) { return 16u * (uint32_t)T; }
// Bytes at the front of cw that RESOLVE uses as scratch (ent, mask x2, wprev, tfirst — the layout resolve_window gives
// them), rounded up to 128: what lies behind survives a window.
PP_HD uint32_t resolve_scratch_bytes(int T)
{
    const uint32_t R = (uint32_t)T * kTileB;
    return (4u * R + 4u * 3u * (R / 32u) + 2u * (max_tiles_for(T) + 2u) + 127u) & ~127u;
}

// One row's tokens of tile `tile` -> ent / mask.  Device: the lanes of the calling warp take
// consecutive tokens, three loads in flight per lane.  Emulation: one thread walks the row.
PP_DEV void scatter_row(const Sm &sm, const uint32_t *tok, const TileTok &tt, uint32_t buf, uint32_t R, uint32_t tile,
                        uint32_t row, uint32_t lane)
{
    const uint32_t nt = sm.ntok[row], tag = tile & 3u;
    uint32_t k0 = tt.kcur[row];
    const uint32_t *rp = tok + row * (uint32_t)kTokRows;
    uint16_t *ent = tt.ent + buf * R;
    uint32_t *mask = tt.mask + buf * (R / 32u);
#ifdef PP_HOST_EMU
    (void)lane;
    uint32_t k = k0;
    for (; k < nt; k++) {
        const uint32_t tv = rp[k];
        if ((tv >> 30) != tag) break;
        const uint32_t pos = (tv >> 16) & 0x3fffu;
        ent[pos] = (uint16_t)tv;
        mask[pos >> 5] |= 1u << (pos & 31u);
    }
    tt.kcur[row] = k;
#else
    for (;;) {
        // a row's tokens of one tile are a run starting at the cursor: lane l looks at tokens k0+l, k0+l+32, k0+l+64
        uint32_t tv[3];
#pragma unroll
        for (int u = 0; u < 3; u++) {
            const uint32_t k = k0 + lane + 32u * (uint32_t)u;
            tv[u] = k < nt ? rp[k] : ((tag ^ 1u) << 30);  // past the row's end: a tag that never matches
        }
        uint32_t cnt = 0;
#pragma unroll
        for (int u = 0; u < 3; u++) {
            if ((tv[u] >> 30) == tag) {
                const uint32_t pos = (tv[u] >> 16) & 0x3fffu;
                ent[pos] = (uint16_t)tv[u];
                atomicOr(&mask[pos >> 5], 1u << (pos & 31u));
                cnt++;
            }
        }
        cnt = __reduce_add_sync(0xffffffffu, cnt);  // the run's length is the cursor's advance
        k0 += cnt;
        if (cnt < 96u) break;
    }
    if (lane == 0u) tt.kcur[row] = k0;
#endif
}

// Rows with tokens in tile `tile`: those that START in it — [tfirst[tile], tfirst[tile+1]) — and the
// one before them, which may reach into it.  Warp w takes rows lo+w, lo+w+nw, ...
PP_DEV void scatter_tile(const Sm &sm, const uint32_t *tok, const TileTok &tt, uint32_t buf, uint32_t R, uint32_t tile,
                         uint32_t t)
{
    const uint32_t f0 = tt.tfirst[tile], hi = tt.tfirst[tile + 1u];
    const uint32_t lo = f0 ? f0 - 1u : 0u;
    const uint32_t warp = t >> 5, lane = t & 31u, nw = ((uint32_t)PP_NT + 31u) >> 5;
#ifdef PP_HOST_EMU
    if (lane != 0u) return;
#endif
    for (uint32_t row = lo + warp; row < hi; row += nw) scatter_row(sm, tok, tt, buf, R, tile, row, lane);
}

// EXPAND of one tile for thread t (see RESOLVE above).  FULL: every byte of the tile is a valid
// output byte (all tiles but the first and the last of a window).  wprev[w] = position of the last
// head before mask word w (expand_prev, run by the same warp just before).
PP_DEV void expand_prev(const uint32_t *mask, uint32_t *wprev, uint32_t t)
{
    const uint32_t lane = t & 31u, warp = t >> 5;
    if (lane < (uint32_t)kTileB) {
        const uint32_t w = warp * (uint32_t)kTileB + lane;
        uint32_t off = w, m = 0;
        while (m == 0u && off > 0u) m = mask[--off];
        wprev[w] = m ? off * 32u + 31u - clz32(m) : 0u;  // (no head before a valid byte: cannot happen for consistent tokens)
    }
}

template <bool FULL>
PP_DEV void expand_tile(const Sm &sm, const uint16_t *ent, const uint32_t *mask, const uint32_t *wprev,
                        const uint8_t *vbase, uint32_t tb, uint32_t a, uint32_t vend, int32_t near_lo, uint32_t t)
{
    // A lane owns byte PAIRS: bytes 64 s + 2 lane, + 1 of the warp's 512 (s = 0..7).  The two bytes share
    // the mask word, the head search and — unless the second byte starts a token — the entry.
    const uint32_t lane = t & 31u, warp = t >> 5;
    const uint32_t qb = warp * (32u * kTileB) + 2u * lane;  // this lane's first byte; its other pairs follow 64 apart
    const uint32_t bw = lane >> 4, bp = 2u * (lane & 15u);  // the pair's mask word within the step (0/1), its first bit there
    const uint32_t below = 0xffffffffu >> (31u - bp);       // bits 0..bp
    const int32_t safe = near_lo - 1;                       // the byte just before the tile/window: final, inside the slot
    const SAddr ent_s = saddr(ent), res_s = saddr(sm.res) + 2u * qb;
    const uint32_t w0 = warp * (uint32_t)kTileB + bw;       // mask word of the lane's pair in step 0; + 2 per step
    const SAddr mask_w = saddr(mask) + 4u * w0, prev_w = saddr(wprev) + 4u * w0;
    const int32_t vq = (int32_t)(tb + qb);                  // virtual index of the lane's first byte
    // All sixteen source loads of the lane are issued before the first one is consumed (one exposed
    // round trip per tile instead of two): per byte only a 16-bit "what it will be" is kept meanwhile —
    // the literal, the tile-relative index of an in-tile source, or kFar = "the byte being loaded".
    constexpr uint32_t kFar = 0x7fffu;                      // not a tile index: R <= 16384
    uint32_t pre[kTileB / 2];                               // two per word: bytes 64 s + 2 lane, + 1
    uint32_t b[kTileB];
#pragma unroll
    for (int s = 0; s < kTileB / 2; s++) {
        const uint32_t q0 = qb + 64u * (uint32_t)s;
        const uint32_t v0 = (uint32_t)vq + 64u * (uint32_t)s;
        const uint32_t mw = lds_u32(mask_w, 8u * (uint32_t)s);
        const uint32_t m0 = mw & below;
        // first byte's token: the last head at or before it — in this word, else the last one before the word
        const uint32_t hp0 = m0 ? 32u * (w0 + 2u * (uint32_t)s) + 31u - clz32(m0) : lds_u32(prev_w, 8u * (uint32_t)s);
        uint32_t x0 = lds_u16(ent_s, 2u * hp0);
        uint32_t i0 = q0 - hp0;
        // second byte: a new token if its head bit is set, else the same token one byte further
        uint32_t x1 = x0, i1 = i0 + 1u;
        if ((mw >> (bp + 1u)) & 1u) { x1 = lds_u16(ent_s, 2u * (q0 + 1u)); i1 = 0u; }
        if (i0 > x0) {                                      // a match (never true for a literal: x >= 0x8000 > i) whose
            const uint32_t dist = x0 + 1u;                  // offset reached its distance: overlapping run, take the same byte
            x0 = dist * (div_small(i0, dist) + 1u) - 1u;    // one or more periods earlier, in front of the match
        }
        if (i1 > x1) {
            const uint32_t dist = x1 + 1u;
            x1 = dist * (div_small(i1, dist) + 1u) - 1u;
        }
        // virtual index of the source; for a literal (0x8000 | byte) a harmless address at most 256
        // bytes back (there is always that much in front of an output: a 32 KB window, or the
        // spare bytes the runtime keeps in front of the first slot)
        int32_t sv0 = (int32_t)v0 - (int32_t)(x0 & 0x7fffu) - 1;
        int32_t sv1 = (int32_t)v0 - (int32_t)(x1 & 0x7fffu);
        if (!FULL) {                                        // bytes outside the window: harmless literals
            if (v0 < a || v0 >= vend) { x0 = 0x8000u; sv0 = safe; }
            if (v0 + 1u < a || v0 + 1u >= vend) { x1 = 0x8000u; sv1 = safe; }
        }
        // unconditional loads (a source inside the tile reads the byte just before the tile instead)
        b[2 * s] = vbase[sv0 < safe ? sv0 : safe];
        b[2 * s + 1] = vbase[sv1 < safe ? sv1 : safe];
        const uint32_t p0 = (x0 & 0x8000u) ? x0 : (sv0 < near_lo ? kFar : (uint32_t)(sv0 - (int32_t)tb));
        const uint32_t p1 = (x1 & 0x8000u) ? x1 : (sv1 < near_lo ? kFar : (uint32_t)(sv1 - (int32_t)tb));
        pre[s] = (p0 & 0xffffu) | (p1 << 16);
    }
#pragma unroll
    for (int s = 0; s < kTileB / 2; s++) {
        uint32_t o0 = pre[s] & 0xffffu, o1 = pre[s] >> 16;
        if (o0 == kFar) o0 = 0x8000u | b[2 * s];
        if (o1 == kFar) o1 = 0x8000u | b[2 * s + 1];
        sts_u32(res_s, 128u * (uint32_t)s, o0 | (o1 << 16));
    }
}

PP_DEV void resolve_window(const Sm &sm, const uint32_t *tok, uint8_t *outp, uint32_t a, uint32_t total, uint32_t nlive,
                           uint32_t rshift)
{
    const int T = PP_NT;
    const uint32_t R = (uint32_t)T * kTileB;
    const uint32_t vend = a + total;         // valid bytes: a <= v < vend
    uint8_t *vbase = outp - a;               // 16-byte aligned; vbase[v] is the byte of virtual index v
    volatile uint16_t *res = sm.res;
    TileTok tt;
    tt.ent = reinterpret_cast<uint16_t *>(sm.cw);      // cw is restaged for every window anyway
    tt.mask = sm.cw + R;                               // 2 R u16 = R words
    tt.kcur = sm.ns;                                   // idle outside SYNC
    tt.wprev = tt.mask + 2u * (R / 32u);               // R/32 words
    tt.tfirst = reinterpret_cast<uint16_t *>(tt.wprev + R / 32u);  // max_tiles_for(T) + 2 u16
    const uint32_t ntiles = (vend + R - 1u) >> rshift;
    // the window's row map: tfirst[i] = first row that starts in tile i or later
    PP_FOR_T(t)
    for (uint32_t i = (uint32_t)t; i < 2u * (R / 32u); i += (uint32_t)T) tt.mask[i] = 0;
    for (uint32_t i = (uint32_t)t; i <= ntiles + 1u; i += (uint32_t)T) tt.tfirst[i] = (uint16_t)nlive;
    tt.kcur[t] = 0;
    PP_END_T
    PP_SYNC();
    PP_FOR_T(t)
    if ((uint32_t)t < nlive && (t == 0 || sm.outc[t] != sm.outc[t - 1])) {  // (an empty row starts where the next one does)
        const uint32_t ti = (a + sm.outc[t]) >> rshift;
        if (ti <= ntiles) {
            // 16-bit atomic min through the containing word
            uint32_t *w32 = reinterpret_cast<uint32_t *>(tt.tfirst) + (ti >> 1);
            const uint32_t sh = (ti & 1u) * 16u;
#ifdef PP_HOST_EMU
            if ((uint32_t)t < ((*w32 >> sh) & 0xffffu)) *w32 = (*w32 & ~(0xffffu << sh)) | ((uint32_t)t << sh);
#else
            uint32_t old = *w32;
            while ((uint32_t)t < ((old >> sh) & 0xffffu)) {
                const uint32_t assumed = old;
                old = atomicCAS(w32, assumed, (assumed & ~(0xffffu << sh)) | ((uint32_t)t << sh));
                if (old == assumed) break;
            }
#endif
        }
    }
    PP_END_T
    PP_SYNC();
    PP_T0_BEGIN
    for (uint32_t i = ntiles + 1u; i-- > 0u;)  // suffix minimum: a tile no row starts in inherits the next one's
        if (tt.tfirst[i] > tt.tfirst[i + 1u]) tt.tfirst[i] = tt.tfirst[i + 1u];
    PP_T0_END
    PP_SYNC();
    PP_FOR_T(t)
    scatter_tile(sm, tok, tt, 0, R, 0, (uint32_t)t);
    PP_END_T
    PP_SYNC();
    uint32_t tile = 0;
    for (uint32_t tb = 0; tb < vend; tb += R, tile++) {
        const uint32_t cb = tile & 1u;
        const int32_t near_lo = (int32_t)(tb > a ? tb : a);  // sources below this are final in global memory
        // EXPAND: the byte's token is the last head at or before it
        PP_FOR_W(t)
        expand_prev(tt.mask + cb * (R / 32u), tt.wprev, (uint32_t)t);
        PP_WARP_SPLIT(t)
        {
            const uint16_t *ent = tt.ent + cb * R;
            const uint32_t *mask = tt.mask + cb * (R / 32u);
            if (tb >= a && tb + R <= vend) expand_tile<true>(sm, ent, mask, tt.wprev, vbase, tb, a, vend, near_lo, (uint32_t)t);
            else expand_tile<false>(sm, ent, mask, tt.wprev, vbase, tb, a, vend, near_lo, (uint32_t)t);
        }
        PP_END_W
        PP_SYNC();
        PP_PHASE(PH_R_GATHER);
        // this tile's head mask is free again; the next tile's tokens go to the other buffer
        PP_FOR_T(t)
        if ((uint32_t)t < R / 32u) tt.mask[cb * (R / 32u) + (uint32_t)t] = 0;
        if (tb + R < vend) scatter_tile(sm, tok, tt, cb ^ 1u, R, tile + 1u, (uint32_t)t);
        PP_END_T
        // CHASE: follow in-tile sources through shared memory (published pointer jumping, no barrier)
        PP_FOR_W(t)
        {
            const uint32_t qb = ((uint32_t)t >> 5) * (32u * kTileB) + ((uint32_t)t & 31u);
            const uint32_t pb = qb;
            uint32_t e[kTileB];
            uint32_t pend = 0;
#pragma unroll
            for (int j = 0; j < kTileB; j++) {
                e[j] = res[pb + (uint32_t)j * 32u];
                pend |= ~e[j] & 0x8000u;
            }
            while (pend) {
                pend = 0;
#pragma unroll
                for (int j = 0; j < kTileB; j++) {
                    if (!(e[j] & 0x8000u)) {
                        uint32_t y = res[res_pos(e[j])];
                        if (!(y & 0x8000u)) y = res[res_pos(y)];  // two hops per round
                        e[j] = y;
                        res[pb + (uint32_t)j * 32u] = (uint16_t)y;  // publish the hop
                        pend |= ~y & 0x8000u;
                    }
                }
            }
        }
        // STORE: every entry of the warp's 512 bytes is 0x8000|byte in shared memory now; thread t takes
        // group t (bytes 16 t .. 16 t + 15 of the tile) and writes it with one 16-byte store
        PP_WARP_SPLIT(t)
        {
            const uint32_t q0 = (uint32_t)t * kTileB;
            const uint32_t v0 = tb + q0;
            uint32_t lo = v0 < a ? a - v0 : 0u;
            uint32_t hi = v0 < vend ? (vend - v0 < (uint32_t)kTileB ? vend - v0 : (uint32_t)kTileB) : 0u;
            if (lo > hi) lo = hi;
            if (lo < hi) {
                const uint4 *src = reinterpret_cast<const uint4 *>(sm.res + q0);
                const uint4 r0 = src[0], r1 = src[1];
                uint32_t w[4];
#ifdef PP_HOST_EMU
#define PP_PACK(x, y) (((x) & 0xffu) | (((x) >> 8) & 0xff00u) | (((y) & 0xffu) << 16) | (((y) << 8) & 0xff000000u))
#else
#define PP_PACK(x, y) __byte_perm((x), (y), 0x6420)
#endif
                w[0] = PP_PACK(r0.x, r0.y); w[1] = PP_PACK(r0.z, r0.w);
                w[2] = PP_PACK(r1.x, r1.y); w[3] = PP_PACK(r1.z, r1.w);
#undef PP_PACK
                uint32_t nl = 0, nul = 0;
                if (hi - lo == (uint32_t)kTileB) {
                    uint32_t nz = 0x80808080u;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint32_t x = ((w[j] & 0x7f7f7f7fu) ^ 0x0a0a0a0au) + 0x7f7f7f7fu;
                        nl += popc32(~(x | w[j]) & 0x80808080u);                 // bytes equal to '\n'
                        nz &= ((w[j] & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w[j];        // bit 7 of a byte survives iff it is not NUL
                    }
                    nul = nz != 0x80808080u;
                    uint4 o;
                    o.x = w[0]; o.y = w[1]; o.z = w[2]; o.w = w[3];
                    *reinterpret_cast<uint4 *>(vbase + v0) = o;
                } else {  // first / last group of the window
                    for (uint32_t j = lo; j < hi; j++) {
                        const uint32_t byte = (uint32_t)res[q0 + j] & 0xffu;
                        vbase[v0 + j] = (uint8_t)byte;
                        nl += (byte == 0x0au);
                        nul |= (byte == 0u);
                    }
                }
                sm.nl[t] += nl;
                sm.nul[t] |= nul;
            }
        }
        PP_END_W
        PP_SYNC();  // stores visible to the next tile's gathers; res free again; the next tile's tokens are in place
        PP_PHASE(PH_R_CHASE);
    }
}

// The decode half of a window — GUESS, SYNC, SCAN — which needs no history and writes no output:
// where every sub-sequence really starts, how many bytes each produces, where the window (or the
// block, or the wanted output) ends.  Also what the block scanner of the GPU-assisted CreateIndex
// runs (csrc/blockscan.cu).  s0: window-relative bit of the first symbol; a: misalignment of the
// output pointer; room: output bytes still wanted (> 0).
struct WindowCount {
    uint32_t next_bit, need_bit, produced, flag, rounds, nlive;
};
PP_DEV WindowCount count_window(const Sm &sm, uint32_t s0, uint32_t room)
{
    const int T = PP_NT;
    uint32_t *cp = reinterpret_cast<uint32_t *>(sm.res);  // checkpoints live in the (idle) resolve tile buffer
    // GUESS
    PP_FOR_T(t)
    {
        const uint32_t st = s0 + (uint32_t)t * kSubBits;
        const Seg r = decode_count(sm, st, s0 + (uint32_t)(t + 1) * kSubBits, st, cp, (uint32_t)T, (uint32_t)t,
                                   false, 0, 0);
        sm.start[t] = st;
        sm.end[t] = r.end;
        sm.outc[t] = r.out;
        sm.flag[t] = r.flag;
    }
    PP_END_T
    PP_SYNC();
    PP_PHASE(PH_GUESS);
    // SYNC: u[8] = first thread whose start is not its predecessor's end, u[9] = first flagged thread
    uint32_t rounds = 0;
    for (;;) {
        PP_T0_BEGIN
        sm.u[8] = (uint32_t)T;
        sm.u[9] = (uint32_t)T;
        PP_T0_END
        PP_SYNC();
        PP_FOR_T(t)
        if (t > 0 && sm.start[t] != sm.end[t - 1]) PP_ATOMIC_MIN(&sm.u[8], (uint32_t)t);
        if (sm.flag[t] != F_NONE) PP_ATOMIC_MIN(&sm.u[9], (uint32_t)t);
        PP_END_T
        PP_SYNC();
        const uint32_t m = sm.u[8], f = sm.u[9];
        if (m == (uint32_t)T || f < m) break;
        rounds++;
        // every thread from the first mismatch on restarts where its predecessor ended
        PP_FOR_T(t)
        sm.ns[t] = t ? sm.end[t - 1] : sm.start[0];
        PP_END_T
        PP_SYNC();
        PP_FOR_T(t)
        {
            const uint32_t ns = sm.ns[t];
            if ((uint32_t)t >= m && ns != sm.start[t]) {
                const uint32_t lim = s0 + (uint32_t)(t + 1) * kSubBits;
                Seg r;
                if (ns >= lim) {
                    r.end = ns; r.out = 0; r.flag = F_NONE; r.ntok = 0;
                    for (int j = 0; j < kCkpt; j++) cp[(uint32_t)j * (uint32_t)T + (uint32_t)t] = kNoCkpt;  // no path left to reuse
                }
                else r = decode_count(sm, ns, lim, s0 + (uint32_t)t * kSubBits, cp, (uint32_t)T, (uint32_t)t, true,
                                      sm.end[t], sm.flag[t]);
                sm.start[t] = ns;
                sm.end[t] = r.end;
                sm.outc[t] = r.out;
                sm.flag[t] = r.flag;
            }
        }
        PP_END_T
        PP_SYNC();
    }
    const uint32_t f = sm.u[9];
    PP_SYNC();
    PP_PHASE(PH_SYNC);
    // SCAN: threads past the first flagged one produce nothing
    PP_FOR_T(t)
    if ((uint32_t)t > f) sm.outc[t] = 0;
    PP_END_T
    PP_SYNC();
    // outc -> exclusive sums; keep each thread's own count in nl-free scratch: recomputed from neighbours below
    uint32_t total = block_excl_scan(sm, sm.outc);
    // live threads: up to the flagged one, cut where the output is full or the window would need more
    // resolve tiles than its row map holds
    const uint32_t capv = (max_tiles_for(T) - 1u) * ((uint32_t)T * kTileB) - 16u;
    PP_T0_BEGIN
    sm.u[10] = f < (uint32_t)T ? f + 1u : (uint32_t)T;  // nlive
    PP_T0_END
    PP_SYNC();
    PP_FOR_T(t)
    {
        const uint32_t endsum = (t + 1 < T) ? sm.outc[t + 1] : total;  // inclusive sum of thread t
        // the first thread whose output crosses the capacity is cut (never thread 0: one sub-sequence
        // emits < 128 Ki bytes, the capacity is 256 T^2 >= 256 Ki)
        if (t > 0 && (uint32_t)t <= f && endsum > capv) PP_ATOMIC_MIN(&sm.u[10], (uint32_t)t);
        // the first thread that completes the wanted output is the last live one
        if ((uint32_t)t <= f && endsum >= room) PP_ATOMIC_MIN(&sm.u[10], (uint32_t)t + 1u);
    }
    PP_END_T
    PP_SYNC();
    const uint32_t nlive = sm.u[10];
    uint32_t produced = nlive < (uint32_t)T ? sm.outc[nlive] : total;
    const uint32_t next_bit = sm.end[nlive - 1u];
    uint32_t flag = (f < (uint32_t)T && nlive == f + 1u) ? sm.flag[f] : (uint32_t)F_NONE;
    PP_T0_BEGIN
    sm.u[23] = next_bit;  // every symbol of the window was needed ...
    PP_T0_END
    PP_SYNC();
    if (produced >= room) {
        // Core.cs:187: the loop ends as soon as the wanted bytes are there (zlib stops mid-block)
        // ... except here: the input is needed only up to the symbol that completes the output
        PP_FOR_T(t)
        if ((uint32_t)t == nlive - 1u)
            sm.u[23] = seg_pos_at_output(sm, sm.start[t], s0 + (uint32_t)(t + 1) * kSubBits, room - sm.outc[t]);
        PP_END_T
        produced = room;
        if (flag == F_BAD) flag = F_NONE;
    }
    PP_SYNC();
    PP_PHASE(PH_SCAN);
    WindowCount c;
    c.next_bit = next_bit;
    c.need_bit = sm.u[23];
    c.produced = produced;
    c.flag = flag;
    c.rounds = rounds;
    c.nlive = nlive;
    PP_SYNC();
    return c;
}

// One window of a Huffman block: GUESS, SYNC, SCAN (count_window), then EMIT and RESOLVE.
// outp2 (dual output, GPU CreateIndex): the same tokens resolved a second time against a second history —
// one Huffman decode, two LZ77 resolves.  outp2 has outp's alignment modulo 16.
template <bool DUAL = false>
PP_DEV WindowOut huffman_window(const Sm &sm, uint32_t s0, uint32_t *tok, uint32_t rshift, uint8_t *outp, uint32_t room,
                                uint8_t *outp2 = nullptr)
{
    const int T = PP_NT;
    const WindowCount c = count_window(sm, s0, room);
    const uint32_t a = (uint32_t)((uintptr_t)outp & 15u);
    const uint32_t nlive = c.nlive, produced = c.produced;
    // EMIT
    PP_FOR_T(t)
    {
        uint32_t nt = 0;
        if ((uint32_t)t < nlive) {
            const uint32_t lim = s0 + (uint32_t)(t + 1) * kSubBits;
            const uint32_t st = sm.start[t];
            if (st < lim) nt = decode_seg<1>(sm, st, lim, tok, rshift, (uint32_t)t, a + sm.outc[t]).ntok;
        }
        sm.ntok[t] = nt;
    }
    PP_END_T
    PP_SYNC();
    PP_PHASE(PH_EMIT);
    // RESOLVE
    resolve_window(sm, tok, outp, a, produced, nlive, rshift);
    if (DUAL) resolve_window(sm, tok, outp2, a, produced, nlive, rshift);
    PP_PHASE(PH_RESOLVE);
    WindowOut w;
    w.next_bit = c.next_bit;
    w.need_bit = c.need_bit;
    w.produced = produced;
    w.flag = c.flag;
    w.rounds = c.rounds;
    return w;
}

// Stored block body: `n` bytes straight from the compressed buffer to the output.
PP_DEV void stored_copy(const Sm &sm, const uint8_t *src, uint8_t *dst, uint32_t n)
{
    const int T = PP_NT;
    PP_FOR_T(t)
    {
        uint32_t nl = 0, nul = 0;
        for (uint32_t j = (uint32_t)t; j < n; j += (uint32_t)T) {
            const uint32_t c = src[j];
            dst[j] = (uint8_t)c;
            nl += (c == 10u);
            nul |= (c == 0u);
        }
        sm.nl[t] += nl;
        sm.nul[t] |= nul;
    }
    PP_END_T
    PP_SYNC();
}

// Whole chunk: Core.ExtractDeflateIndex for one (from, to) pair.
// scratch: this CTA's token rows + group index (global memory, scratch_words_for(T) words).
// Pipelined upload: the compressed range reaches the device in file order, in pieces, on a copy stream
// while the kernel runs; after every piece the host publishes the number of bytes in place (then
// "everything").  A CTA waits, window by window, until the bytes it is about to stage are there.
// (Tried and measured slower on 10 M reads, 24.0 ms pulled / 24.7 ms file order: delivering "waves" of
// chunks column by column with 2-D copies so that every resident CTA gets its first bytes at once —
// strided 64 KB rows reach only ~38 GB/s against 55 GB/s for plain copies, 27.6 ms; a first wave pulled
// by the kernel while the copy engine brings the rest — the two halve each other's link, 24.7 ms.)
struct ByteGate {
    const volatile unsigned long long *mark;  // device memory, written by the copy stream (null: no gate)
    uint64_t total;                            // bytes of the range
    uint64_t shift;                            // bytes the kernel's base pointer was moved down (alignment)
};
PP_HD unsigned long long gate_need(const ByteGate &g, uint64_t hi) { return hi < g.total ? hi : g.total; }
// Wait until bytes [lo, hi) (kernel coordinates) are in place.  One thread polls; bounded (~4 s): a copy
// that never arrives must not hang the GPU.  Returns false on time-out.
PP_DEV bool gate_wait(const Sm &sm, const ByteGate *g, uint64_t lo, uint64_t hi)
{
#ifdef PP_HOST_EMU
    (void)sm; (void)g; (void)lo; (void)hi;
    return true;
#else
    if (!g || !g->mark) return true;
    if (threadIdx.x == 0) {
        lo = lo > g->shift ? lo - g->shift : 0u;
        hi = hi > g->shift ? hi - g->shift : 0u;
        (void)lo;
        const unsigned long long need = gate_need(*g, hi);
        unsigned ok = 1;
        if (*g->mark < need) {
            const long long t0 = clock64();
            while (*g->mark < need) {
                __nanosleep(100);
                if (clock64() - t0 > 8000000000LL) { ok = 0; break; }
            }
            atomicAdd(&g_phase_cycles[PH_WAIT], (unsigned long long)(clock64() - t0));
        }
        sm.u[17] = ok;
    }
    __syncthreads();
    const bool ok = sm.u[17] != 0;
    __syncthreads();
    return ok;
#endif
}

// Dual output (GPU CreateIndex, createindex.cu): every chunk is written a second time, dual.slot_delta bytes
// further on in the slots buffer, against the history dual.lead_delta bytes further on in the lead buffer.
struct DualOut {
    uint64_t slot_delta;  // multiple of 128
    uint64_t lead_delta;  // multiple of 16
};
template <bool DUAL = false, bool PULL = false>
PP_DEV void inflate_chunk(const Sm &sm, const ChunkDesc &d, const uint8_t *comp, uint64_t comp_bytes, uint8_t *slots,
                          const uint8_t *lead_src, uint32_t *scratch, ChunkResult &res, uint32_t &stage_phase,
                          const ByteGate *gate = nullptr, DualOut dual = DualOut{0, 0})
{
    const int T = PP_NT;
    uint8_t *slot = slots + d.slot_off;
    if (DUAL) {  // the second history in front of the second output
        const uint4 *s4 = reinterpret_cast<const uint4 *>(lead_src + d.lead_src + dual.lead_delta);
        uint4 *d4 = reinterpret_cast<uint4 *>(slot + dual.slot_delta);
        const uint32_t n4 = d.lead_len / 16u;
        PP_FOR_T(t)
        for (uint32_t i = (uint32_t)t; i < n4; i += (uint32_t)T) d4[i] = s4[i];
        PP_END_T
    }
    // 1. history: copy the checkpoint window (Core.cs:158 inflateSetDictionary) in front of the output
    {
        // (lead_src == kLeadInPlace: the history is already there — the checkpoint windows were inflated
        // straight into the slots by a pre-pass of this same kernel)
        const uint4 *s4 = reinterpret_cast<const uint4 *>(lead_src + (d.lead_src == kLeadInPlace ? 0 : d.lead_src));
        uint4 *d4 = reinterpret_cast<uint4 *>(slot);
        const uint32_t n4 = d.lead_src == kLeadInPlace ? 0u : d.lead_len / 16u;
        PP_FOR_T(t)
        for (uint32_t i = (uint32_t)t; i < n4; i += (uint32_t)T) d4[i] = s4[i];
        sm.nl[t] = 0;
        sm.nul[t] = 0;
        PP_END_T
    }
    PP_SYNC();
#ifndef PP_HOST_EMU
    if (threadIdx.x == 0) {
        const long long now_ = clock64();
        sm.u[24] = (uint32_t)now_;
        sm.u[25] = (uint32_t)((unsigned long long)now_ >> 32);
    }
#endif
    uint8_t *out = slot + d.lead_len;
    const uint32_t out_len = d.out_len;
    uint32_t *tok = scratch;
    uint32_t rshift = 4;  // log2 of the resolve tile (16 T bytes; T is a power of two)
    while ((1u << rshift) < (uint32_t)T * kTileB) rshift++;
    const uint32_t cww = cw_words_for(T);
    // 2. bit cursor: 8*Input - Bits (Core.cs:151-157 inflatePrime semantics)
    uint64_t bit = d.in_bit;
    uint32_t produced = 0;
    int status = 0;
    bool need_header = true, last = false;
    uint64_t prev_base = ~0ull;  // PULL: base of the window whose staged bytes (those behind RESOLVE's scratch) are still in cw
    const uint32_t keep_from = PULL ? resolve_scratch_bytes(T) : 0u;
    while (produced < out_len) {
        if ((bit >> 3) > d.in_limit) { status = -3; break; }  // Core.cs:174: out of input
        const uint64_t base_byte = (bit >> 3) & ~(uint64_t)15;
        PP_PHASE(PH_OTHER);
        if (!gate_wait(sm, gate, base_byte, base_byte + 4ull * cww)) { status = -100; break; }
        if (PULL && prev_base != ~0ull && base_byte >= prev_base + keep_from && base_byte < prev_base + 4ull * cww) {
            if (!stage_window_reuse(sm, comp, comp_bytes, base_byte, cww, stage_phase, (uint32_t)(base_byte - prev_base))) { status = -100; break; }
        } else {
            if (!stage_window(sm, comp, comp_bytes, base_byte, cww, stage_phase)) { status = -100; break; }
        }
        prev_base = base_byte;
        PP_PHASE(PH_STAGE);
        uint32_t s0 = (uint32_t)(bit - base_byte * 8u);
        if (need_header) {
            const uint32_t hdr = peek_bits(sm.cw, s0, 3);
            s0 += 3;
            last = (hdr & 1u) != 0;
            const uint32_t type = hdr >> 1;
            if (type == 0u) {
                // stored: skip to the byte boundary, LEN / NLEN, then raw bytes
                const uint32_t bp = (s0 + 7u) & ~7u;
                const uint32_t len = peek_bits(sm.cw, bp, 16), nlen = peek_bits(sm.cw, bp + 16u, 16);
                if ((len ^ 0xffffu) != nlen) { status = -3; break; }  // invalid stored block lengths
                const uint64_t byte0 = base_byte + (bp >> 3) + 4u;
                if (byte0 + len > comp_bytes || byte0 + len > d.in_limit) { status = -3; break; }
                uint32_t n = len;
                if (n > out_len - produced) n = out_len - produced;
                PP_SYNC();
                if (!gate_wait(sm, gate, byte0, byte0 + n)) { status = -100; break; }
                stored_copy(sm, comp + byte0, out + produced, n);
                if (DUAL) stored_copy(sm, comp + byte0, out + dual.slot_delta + produced, n);
                produced += n;
                bit = (byte0 + len) * 8u;
                PP_PHASE(PH_STORED);
                if (last) break;  // Z_STREAM_END (Core.cs:185)
                continue;
            }
            int rc;
            if (type == 1u) rc = fixed_tables(sm);
            else if (type == 2u) rc = dynamic_tables(sm, s0, &s0);
            else rc = -3;  // invalid block type
            if (rc) { status = rc; break; }
            if (base_byte * 8u + s0 > d.in_limit * 8u) { status = -3; break; }  // the header itself ran past the input
            PP_PHASE(PH_HEADER);
            need_header = false;
        }
        const WindowOut w = huffman_window<DUAL>(sm, s0, tok, rshift, out + produced, out_len - produced,
                                                 DUAL ? out + dual.slot_delta + produced : nullptr);
        produced += w.produced;
        bit = base_byte * 8u + w.next_bit;
        // Core.cs:174: the reference throws DATA_ERROR when zlib wants input past the end of fileBuffer
        if (base_byte * 8u + w.need_bit > d.in_limit * 8u) { status = -3; break; }
#ifdef PP_HOST_EMU
        g_stat[0]++;
        g_stat[1] += w.rounds;
        if (w.rounds > g_stat[2]) g_stat[2] = w.rounds;
        g_stat[3] += (w.flag == F_EOB);
#endif
        if (w.flag == F_BAD) { status = -3; break; }
        if (w.flag == F_EOB) {
            need_header = true;
            if (last) break;  // Z_STREAM_END (Core.cs:185)
        }
    }
    PP_SYNC();
    PP_PHASE(PH_OTHER);
    // 3. NUL terminator / clean tail for the parse stage
    {
        const uint32_t to = ((d.lead_len + d.out_len + 1u + 127u) & ~127u) - d.lead_len;
        PP_FOR_T(t)
        for (uint32_t i = produced + (uint32_t)t; i < to; i += (uint32_t)T) {
            out[i] = 0;
            if (DUAL) out[dual.slot_delta + i] = 0;
        }
        PP_END_T
    }
    // 4. results
    PP_SYNC();
    PP_T0_BEGIN
    {
        uint32_t nl = 0, nul = 0;
        for (int t = 0; t < T; t++) { nl += sm.nl[t]; nul |= sm.nul[t]; }
        res.status = status;
        res.produced = produced;
        res.newlines = nl;
        res.min_byte = nul ? 0u : 1u;
        res.end_bit = bit;
    }
    PP_T0_END
    PP_SYNC();
}

}  // namespace ppinf

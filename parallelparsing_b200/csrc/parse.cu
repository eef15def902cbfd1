// Kernel 2 — FASTQ parse, plus the small kernels around it.
//
// Replaces Parsing.Parse / ParseLine / CombinedMemory (Decompressor/Parsing.cs:11-117).
// The reference walks the bytes of `from.offset ++ inflated ++ zeros` one at a
// time; for well-formed input that is "every '\n' ends a line, every 4 lines are a
// record, stop at the first NUL".  The fast kernel computes exactly that in one
// streaming pass and proves, per chunk, that the input is in the regime where the
// two definitions agree; a chunk that is not (a NUL inside the data, or an empty
// id/'+' line, whose first byte Parsing.cs:19,30 would skip unchecked) is flagged
// and re-parsed by the exact kernel below, which restates Parsing.cs literally.
//
// Layout: the combined memory of chunk k is CONTIGUOUS in the slots buffer: the
// inflate kernel puts the checkpoint window right before the output and
// from.offset is by construction the tail of that window (Core.cs:86-94,107), so
// no concatenation or copy is needed.  Output is structure-of-arrays: four u32
// arrays (one per FASTQ line) holding, for every record, the combined-memory index
// of the first byte of that line.
#include <cstdlib>

#include "kernels.cuh"

namespace pp {

constexpr int kConsumers = 256;                            // eight consumer warps
constexpr int kParseWarps = kConsumers / 32;
constexpr int kParseThreads = kConsumers + 32;             // + the producer warp
constexpr int kRows = 4;                                   // 16-byte vectors per lane per step
constexpr int kWarpBytes = 32 * 16 * kRows;                // 2 KB per warp per step
constexpr int kIterBytes = kWarpBytes * kParseWarps;       // 16 KB per CTA per step = one ring stage

// 0x80 in every byte of w that equals '\n' (exact for any byte values), three instructions:
// per byte, ((b & 0x7f) ^ 0x0a) + 0x7f has bit 7 set iff the low seven bits differ from 0x0a
// (no carry leaves the byte), and bit 7 of b itself must be clear
__device__ __forceinline__ uint32_t nl_bytes(uint32_t w)
{
    uint32_t x, t;
    asm("lop3.b32 %0, %1, %2, %3, 0x6a;" : "=r"(x) : "r"(w), "r"(0x7f7f7f7fu), "r"(0x0a0a0a0au));  // (w & A) ^ B
    x += 0x7f7f7f7fu;
    asm("lop3.b32 %0, %1, %2, %3, 0x02;" : "=r"(t) : "r"(x), "r"(w), "r"(0x80808080u));  // ~(x | w) & C
    return t;
}

// Tile-parallel: the combined memory of every chunk is cut into 64 KB tiles and the tiles of
// ALL chunks form one work list, so the kernel fills the GPU whatever the chunk count.  A tile
// needs the number of '\n' before it inside its chunk: tiles are handed out by a global ticket
// counter and chained with a decoupled look-back (each tile publishes its own count, then the
// inclusive prefix; warp 0 sums its predecessors' counts, 32 at a time, until it meets a
// prefix).  Tickets go round-robin over the chunks (tile r of every chunk, then tile r+1 ...),
// so the predecessor of a running tile has normally finished long ago and the look-back is one
// step.  One pass over the bytes, no second read.
//
// The CTA is warp-specialised.  A producer warp draws the tickets and streams each tile into a
// 4 x 16 KB shared-memory ring with TMA bulk copies (cp.async.bulk + mbarrier complete_tx), one
// tile ahead of the eight consumer warps, so no register holds bytes in flight and HBM latency is
// off the consumers' path.  The consumers turn every 16-byte vector into a 16-bit newline mask
// (SWAR), regroup the masks through shared memory so that each lane owns 64 CONSECUTIVE bytes (one
// count to scan and one short loop per step instead of four), rank the newlines with warp/block
// scans, write their positions into a shared array by tile-local ordinal, and then emit the line
// starts with all lanes busy: thread j handles newline j, consecutive lanes write consecutive
// records.  (The first version let every lane walk its own mask bits and store straight to global
// memory: five of 32 lanes active, 134 warp instructions per 512 bytes; this one needs ~70.)
// lines: four arrays of `stride` u32 each (line 0..3), record r of the chunk at rec_base + r.
constexpr int kTileIters = 4;                               // 16 KB steps per tile
constexpr int kTileBytesP = kIterBytes * kTileIters;        // 64 KB
#define kFlagAgg (1ull << 32)
#define kFlagPrefix (2ull << 32)

// 16 flag bits (bit i = byte i is '\n') of one 16-byte vector in the LOW HALF of the result; the high
// half is not defined
__device__ __forceinline__ uint32_t nl_mask16(const uint4 x)
{
    // two words per multiply: the flags of the first word at bits 8j+3, of the second at bits 8j+7;
    // times 2^21 + 2^14 + 2^7 + 1 they land, in order and without any carry, in bits 24..31
    const uint32_t u = (nl_bytes(x.x) >> 4) | nl_bytes(x.y);
    const uint32_t v = (nl_bytes(x.z) >> 4) | nl_bytes(x.w);
    return __byte_perm(u * 0x00204081u, v * 0x00204081u, 0x0073);
}

constexpr uint32_t kPosCap = 2048;  // newline positions per emission round (a tile of 150 bp reads holds ~700)
struct TileMsg {
    uint64_t data_off;
    int64_t rec_base;
    uint32_t total, rec_count, skip, k, ti, tile, first_tile, done;
};
struct ParseSm {
    uint4 stage[kTileIters][kIterBytes / 16];
    uint32_t pos[kPosCap + 4];
    uint16_t tr[kParseWarps][kWarpBytes / 16];  // per warp: the 16-bit masks of one step, by vector number
    uint32_t warp_tot[kTileIters * kParseWarps];
    TileMsg msg[4];
    unsigned long long full[kTileIters], empty[kTileIters];
    uint32_t before;
};
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ppinf::smem_u32(bar)) : "memory");
}
// try_wait sleeps in hardware for up to `hint_ns` per attempt, so a waiting warp issues almost nothing
__device__ __forceinline__ void mbar_wait_or_trap(unsigned long long *bar, uint32_t parity, uint32_t hint_ns)
{
    const long long t0 = clock64();
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(ppinf::smem_u32(bar)), "r"(parity), "r"(hint_ns)
            : "memory");
        if (ok) return;
        if (clock64() - t0 > 4000000000LL) __trap();  // ~2 s: a transfer that never lands must not hang the GPU
    }
}

// ticket -> tile; false when the ticket names nothing to do
__device__ __forceinline__ bool ticket_to_tile(const ParseDesc *__restrict__ pdesc, int n, const uint32_t *__restrict__ tile_base,
                                               uint32_t tk, int order, TileMsg &g)
{
    int k;
    uint32_t ti;
    if (order) {
        k = (int)(tk % (uint32_t)n);
        ti = tk / (uint32_t)n;
        if (ti >= tile_base[k + 1] - tile_base[k]) return false;
    } else {
        int lo = 0, hi = n - 1;  // last chunk whose tile_base <= tk
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (tile_base[mid] <= tk) lo = mid; else hi = mid - 1;
        }
        k = lo;
        ti = tk - tile_base[k];
    }
    const ParseDesc d = pdesc[k];
    if (d.exact) return false;  // handled by pp_exact_emit_kernel
    const uint32_t head = (uint32_t)(d.data_off & 15u);
    if ((uint64_t)ti * kTileBytesP >= (uint64_t)head + d.total) return false;  // the chunk produced less than planned
    g.data_off = d.data_off;
    g.rec_base = d.rec_base;
    g.total = d.total;
    g.rec_count = d.rec_count;
    g.skip = d.skip;
    g.k = (uint32_t)k;
    g.ti = ti;
    g.first_tile = tile_base[k];
    g.tile = g.first_tile + ti;
    g.done = 0;
    return true;
}

__device__ __forceinline__ void parse_producer(ParseSm &sm, const uint8_t *__restrict__ slots, const ParseDesc *__restrict__ pdesc,
                                               int n, const uint32_t *__restrict__ tile_base, uint32_t n_tickets, int order,
                                               uint32_t *ticket)
{
    uint32_t tile_no = 0;
    uint32_t tk = atomicAdd(ticket, 1u);
    while (tk < n_tickets) {
        const uint32_t next = atomicAdd(ticket, 1u);  // in flight while this tile is set up
        TileMsg g;
        if (ticket_to_tile(pdesc, n, tile_base, tk, order, g)) {
            // slot tile_no & 3 was last used by tile_no - 4: every consumer left that tile before it
            // released the last stage of tile_no - 3, which the loads of tile_no - 2 waited for
            sm.msg[tile_no & 3u] = g;
            const uint32_t head = (uint32_t)(g.data_off & 15u);
            const uint8_t *src = slots + g.data_off - head + (uint64_t)g.ti * kTileBytesP;
            const uint32_t span16 = (head + g.total + 15u) & ~15u;
            const uint32_t t0 = g.ti * (uint32_t)kTileBytesP;
#pragma unroll
            for (int it = 0; it < kTileIters; it++) {
                mbar_wait_or_trap(&sm.empty[it], (tile_no & 1u) ^ 1u, 20000u);
                const uint32_t o = t0 + (uint32_t)it * kIterBytes;
                const uint32_t bytes = o >= span16 ? 0u : (span16 - o < (uint32_t)kIterBytes ? span16 - o : (uint32_t)kIterBytes);
                if (bytes) {
                    ppinf::mbar_expect_tx(&sm.full[it], bytes);
                    ppinf::tma_load(sm.stage[it], src + (uint32_t)it * kIterBytes, bytes, &sm.full[it]);
                } else {
                    mbar_arrive(&sm.full[it]);
                }
            }
            tile_no++;
        }
        tk = next;
    }
    sm.msg[tile_no & 3u].done = 1u;
    mbar_wait_or_trap(&sm.empty[0], (tile_no & 1u) ^ 1u, 20000u);
    mbar_arrive(&sm.full[0]);
}

// positions of the newlines of this lane's 64-byte blocks -> sm.pos[tile-local ordinal - lo]
template <bool kChecked>
__device__ __forceinline__ void compact_positions(ParseSm &sm, const uint32_t (&mlo)[kTileIters], const uint32_t (&mhi)[kTileIters],
                                                  const uint32_t (&ex)[kTileIters / 2], const uint32_t (&before)[kTileIters],
                                                  uint32_t rel, uint32_t lo)
{
#pragma unroll
    for (int it = 0; it < kTileIters; it++) {
        uint32_t o = before[it] - lo + ((ex[it >> 1] >> ((it & 1) * 16)) & 0xffffu);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            uint32_t mask = h ? mhi[it] : mlo[it];
            const uint32_t off = rel + (uint32_t)it * kIterBytes + (uint32_t)h * 32u;  // wraps only for masked-out bytes
            while (mask) {
                const uint32_t b = (uint32_t)__ffs((int)mask) - 1u;
                mask &= mask - 1u;
                if (!kChecked || o <= kPosCap) sm.pos[o] = off + b;
                o++;
            }
        }
    }
}

__global__ void __launch_bounds__(kParseThreads, 3) pp_parse_kernel(const uint8_t *__restrict__ slots,
                                                                   const ParseDesc *__restrict__ pdesc, int n,
                                                                   const uint32_t *__restrict__ tile_base,
                                                                   uint32_t n_tickets, int order,
                                                                   uint32_t *__restrict__ lines, int64_t stride,
                                                                   ParseOut *__restrict__ pout,
                                                                   const ScanTotals *__restrict__ totals,
                                                                   unsigned long long *tile_state, uint32_t *ticket)
{
    extern __shared__ __align__(128) uint8_t parse_smem_raw[];
    ParseSm &sm = *reinterpret_cast<ParseSm *>(parse_smem_raw);
    if (totals->overflow) return;
    const int lane = (int)(threadIdx.x & 31u), warp = (int)(threadIdx.x >> 5);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < kTileIters; i++) {
            ppinf::mbar_init(&sm.full[i], 1);
            ppinf::mbar_init(&sm.empty[i], kParseWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == kParseWarps) {  // producer
        if (lane == 0) parse_producer(sm, slots, pdesc, n, tile_base, n_tickets, order, ticket);
        return;
    }
    volatile unsigned long long *st = tile_state;
    for (uint32_t tile_no = 0;; tile_no++) {
        const uint32_t par = tile_no & 1u;
        mbar_wait_or_trap(&sm.full[0], par, 2000u);
        const volatile TileMsg *g = &sm.msg[tile_no & 3u];
        if (g->done) break;
        const uint64_t data_off = g->data_off;
        const uint32_t total = g->total, skip = g->skip, k = g->k, ti = g->ti, tile = g->tile;
        const uint32_t rec_total = g->rec_count + skip;  // records in the chunk before skipping
        uint32_t *const l_base = lines + g->rec_base;
        const int64_t first = (int64_t)g->first_tile;
        const uint8_t *data = slots + data_off;
        const uint32_t head = (uint32_t)(data_off & 15u);  // bytes before `data` in its first vector
        const uint32_t span = head + total;                  // bytes from the first vector to the end
        const uint32_t t0 = ti * (uint32_t)kTileBytesP;
        if (threadIdx.x == 0 && ti == 0 && rec_total > skip) l_base[0] = 0;  // record 0 starts at 0 (when not skipped)
        // the byte right after the tile: is it a newline?
        uint32_t after_pos = 0xffffffffu;  // combined-memory index of that byte when it is one
        {
            const uint32_t nxt = t0 + (uint32_t)kTileBytesP;
            if (threadIdx.x == kConsumers - 1 && nxt < span && data[nxt - head] == '\n') after_pos = nxt - head;
        }
        // 1. newline masks from the staged bytes: 16 bits per vector, then regrouped through shared memory so
        //    that every lane holds the 64-bit mask of 64 CONSECUTIVE bytes (one count, one loop per step)
        uint32_t mlo[kTileIters], mhi[kTileIters];
        const bool edge = t0 == 0u || t0 + (uint32_t)kTileBytesP > span;  // block-uniform
        const uint32_t toff = t0 + (uint32_t)warp * kWarpBytes + (uint32_t)lane * 16u;  // (it, r) adds it * 16 KB + r * 512
        uint16_t *const tr = sm.tr[warp];
#pragma unroll
        for (int it = 0; it < kTileIters; it++) {
            if (it) mbar_wait_or_trap(&sm.full[it], par, 2000u);
            uint4 v[kRows];
#pragma unroll
            for (int r = 0; r < kRows; r++) v[r] = sm.stage[it][warp * (kWarpBytes / 16) + r * 32 + lane];
#pragma unroll
            for (int r = 0; r < kRows; r++) {
                uint32_t mask = nl_mask16(v[r]);
                // bytes outside [head, span) are not part of the chunk (only a chunk's first and last tile)
                if (edge) {
                    mask &= 0xffffu;
                    const uint32_t off = toff + (uint32_t)it * kIterBytes + (uint32_t)r * 512u;
                    if (off >= span) mask = 0;
                    if (off < head) mask &= 0xffffu << (head - off);
                    if (off + 16u > span && off < span) mask &= 0xffffu >> (off + 16u - span);
                }
                tr[r * 32 + lane] = (uint16_t)mask;
            }
            __syncwarp();
            const uint2 mm = *reinterpret_cast<const uint2 *>(tr + lane * 4);  // vectors 4*lane .. 4*lane+3
            mlo[it] = mm.x;
            mhi[it] = mm.y;
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty[it]);  // this warp is done with the stage
        }
        // 2. per-step counts, packed two per register, inclusive warp scans
        uint32_t cnt2[kTileIters / 2], ex[kTileIters / 2];
#pragma unroll
        for (int h = 0; h < kTileIters / 2; h++) {
            cnt2[h] = (uint32_t)(__popc(mlo[2 * h]) + __popc(mhi[2 * h])) |
                      ((uint32_t)(__popc(mlo[2 * h + 1]) + __popc(mhi[2 * h + 1])) << 16);
            ex[h] = cnt2[h];
        }
#pragma unroll
        for (int sft = 1; sft < 32; sft <<= 1) {
#pragma unroll
            for (int h = 0; h < kTileIters / 2; h++) {
                const uint32_t a = __shfl_up_sync(0xffffffffu, ex[h], sft);
                if (lane >= sft) ex[h] += a;
            }
        }
#pragma unroll
        for (int h = 0; h < kTileIters / 2; h++) {
            const uint32_t t = __shfl_sync(0xffffffffu, ex[h], 31);
            if (lane == 0) {
                sm.warp_tot[(2 * h) * kParseWarps + warp] = t & 0xffffu;
                sm.warp_tot[(2 * h + 1) * kParseWarps + warp] = t >> 16;
            }
            ex[h] -= cnt2[h];  // inclusive -> exclusive (no borrow between the halves)
        }
        consumer_sync();
        // newlines of the tile before this warp's block of step `it`: one warp scan over the 32 block totals
        uint32_t before[kTileIters], all;
        {
            static_assert(kTileIters * kParseWarps == 32, "one total per lane");
            const uint32_t mine = sm.warp_tot[lane];
            uint32_t inc = mine;
#pragma unroll
            for (int sft = 1; sft < 32; sft <<= 1) {
                const uint32_t a = __shfl_up_sync(0xffffffffu, inc, sft);
                if (lane >= sft) inc += a;
            }
            all = __shfl_sync(0xffffffffu, inc, 31);
#pragma unroll
            for (int it = 0; it < kTileIters; it++) before[it] = __shfl_sync(0xffffffffu, inc - mine, it * kParseWarps + warp);
        }
        if (threadIdx.x == 0) st[tile] = (ti == 0 ? kFlagPrefix : kFlagAgg) | all;  // successors can go on
        uint32_t flags = 0, run = 0;
        uint32_t lo = 0;
        do {  // one round unless the tile holds more than kPosCap newlines
            if (lo) consumer_sync();  // the previous round's positions have been consumed
            // 3. compaction (slot kPosCap = first position of the next round)
            const uint32_t rel = t0 + (uint32_t)warp * kWarpBytes + (uint32_t)lane * 64u - head;
            if (all <= kPosCap) compact_positions<false>(sm, mlo, mhi, ex, before, rel, 0u);
            else compact_positions<true>(sm, mlo, mhi, ex, before, rel, lo);
            if (threadIdx.x == kConsumers - 1 && all - lo <= kPosCap) sm.pos[all - lo] = after_pos;
            // 4. decoupled look-back over the earlier tiles of this chunk (warp 0, 32 tiles per step)
            if (lo == 0 && warp == 0) {
                uint32_t excl = 0;
                if (ti != 0) {
                    int64_t j = (int64_t)tile - 1;
                    for (;;) {
                        const int64_t idx = j - lane;
                        unsigned long long v;
                        uint32_t ready, pfx;
                        do {
                            v = idx >= first ? st[idx] : kFlagPrefix;  // before the chunk: prefix 0
                            ready = __ballot_sync(0xffffffffu, (v >> 32) != 0ull);
                            pfx = __ballot_sync(0xffffffffu, (v >> 32) == 2ull);
                            // lanes 0..p are needed, p = nearest tile holding a prefix (all 32 if none)
                        } while ((ready | ~(pfx ? ((pfx & (0u - pfx)) << 1) - 1u : 0xffffffffu)) != 0xffffffffu);
                        const uint32_t need = pfx ? ((pfx & (0u - pfx)) << 1) - 1u : 0xffffffffu;
                        uint32_t val = ((need >> lane) & 1u) ? (uint32_t)v : 0u;
#pragma unroll
                        for (int sft = 16; sft > 0; sft >>= 1) val += __shfl_xor_sync(0xffffffffu, val, sft);
                        excl += val;
                        if (pfx) break;
                        j -= 32;
                    }
                    if (lane == 0) st[tile] = kFlagPrefix | (unsigned long long)(excl + all);
                }
                if (lane == 0) {
                    sm.before = excl;
                    if (t0 + (uint32_t)kTileBytesP >= span) pout[k].newlines = excl + all;
                }
            }
            consumer_sync();
            if (lo == 0) run = sm.before;
            // 5. newline `o` of the chunk ends line o: line L = o + 1 starts right after it
            const uint32_t cnt = all - lo < kPosCap ? all - lo : kPosCap;
            for (uint32_t j = threadIdx.x; j < cnt; j += (uint32_t)kConsumers) {
                const uint32_t pos = sm.pos[j], nxt = sm.pos[j + 1];
                const uint32_t L = run + lo + j + 1u;
                const uint32_t rec = L >> 2, f = L & 3u;
                // an empty line 0 or 2 is where Parsing.cs:19,30 diverge from "every '\n' ends a line"
                if (nxt == pos + 1u && (L & 1u) == 0u && rec < rec_total) flags |= 1u;
                if (pos == 0u) flags |= 1u;  // the chunk starts with an empty id line
                if (rec >= skip) {
                    if (rec < rec_total) l_base[(int64_t)f * stride + (rec - skip)] = pos + 1u;
                    else if (L == rec_total * 4u) pout[k].parse_end = pos + 1u;
                }
            }
            lo += kPosCap;
        } while (lo < all);
        if (flags) atomicOr(&pout[k].flags, flags);
    }
}

// ---- exact parser: Parsing.cs:11-69 literally, one thread per chunk -------------
struct ExactCursor {
    const uint8_t *d;
    uint32_t total;
    __device__ uint32_t at(uint32_t i) const { return i < total ? d[i] : 0u; }  // zero tail (H3)
};
// ParseLine (Parsing.cs:54-69): returns false when it meets NUL before '\n'
__device__ bool exact_line(const ExactCursor &c, uint32_t &pos)
{
    for (;;) {
        const uint32_t b = c.at(pos);
        if (b == '\n') break;
        if (b == 0u) return false;
        pos++;
    }
    pos++;
    return true;
}
// emit != nullptr: writes line starts; returns the record count
__device__ uint32_t exact_parse(const ExactCursor &c, uint32_t skip, uint32_t *l_base, int64_t stride, bool emit,
                                uint32_t *parse_end)
{
    uint32_t i = 0, n = 0, end = 0;
    while (i < c.total) {
        if (c.at(i) == 0u) break;  // :16
        const uint32_t l0 = i;
        i++;                       // :19 skip '@' unchecked
        if (!exact_line(c, i)) break;
        const uint32_t l1 = i;
        if (!exact_line(c, i)) break;
        const uint32_t l2 = i;
        i++;                       // :30 skip '+' unchecked
        if (!exact_line(c, i)) break;
        const uint32_t l3 = i;
        if (!exact_line(c, i)) break;
        if (emit && n >= skip) {
            const uint32_t r = n - skip;
            l_base[r] = l0;
            l_base[stride + r] = l1;
            l_base[2 * stride + r] = l2;
            l_base[3 * stride + r] = l3;
        }
        end = i;
        n++;
    }
    if (parse_end) *parse_end = end;
    return n;
}

__global__ void pp_exact_count_kernel(const uint8_t *__restrict__ slots, const ChunkDesc *__restrict__ descs,
                                      const ChunkResult *__restrict__ results, const ParseOut *__restrict__ pout,
                                      int n, int64_t *__restrict__ exact_counts)
{
    const int k = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (k >= n) return;
    const bool need = exact_counts[k] >= 0 || results[k].min_byte == 0u || (pout && (pout[k].flags & 1u));
    if (!need || results[k].status != 0) return;
    const ChunkDesc d = descs[k];
    ExactCursor c{slots + d.slot_off + d.lead_len - d.prefix_len, d.prefix_len + results[k].produced};
    exact_counts[k] = (int64_t)exact_parse(c, 0, nullptr, 0, false, nullptr);
}

__global__ void pp_exact_emit_kernel(const uint8_t *__restrict__ slots, const ParseDesc *__restrict__ pdesc, int n,
                                     uint32_t *__restrict__ lines, int64_t stride, ParseOut *__restrict__ pout,
                                     const ScanTotals *__restrict__ totals)
{
    const int k = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (k >= n || totals->overflow) return;
    const ParseDesc d = pdesc[k];
    if (!d.exact) return;
    ExactCursor c{slots + d.data_off, d.total};
    uint32_t end = 0;
    const uint32_t cnt = exact_parse(c, d.skip, lines + d.rec_base, stride, true, &end);
    pout[k].parse_end = end;
    pout[k].flags = 2u;
    pout[k].newlines = 0;
    pout[k].records = cnt - (d.skip < cnt ? d.skip : cnt);
}

// ---- byte statistics for buffers that did not come out of the inflate kernel ------
// (pp_parse on caller-provided bytes): '\n' count and minimum byte up to the first NUL.
__global__ void __launch_bounds__(256) pp_bytes_stats_kernel(const uint8_t *__restrict__ slots,
                                                             const ChunkDesc *__restrict__ descs,
                                                             ChunkResult *__restrict__ results, int n)
{
    __shared__ uint32_t s_nl, s_min;
    const int k = (int)blockIdx.x;
    if (k >= n) return;
    if (threadIdx.x == 0) { s_nl = 0; s_min = 255u; }
    __syncthreads();
    const ChunkDesc d = descs[k];
    const uint8_t *p = slots + d.slot_off + d.lead_len;
    uint32_t nl = 0, mn = 255u;
    for (uint32_t i = threadIdx.x; i < d.out_len; i += blockDim.x) {
        const uint32_t b = p[i];
        nl += (b == 10u);
        mn = b < mn ? b : mn;
    }
    atomicAdd(&s_nl, nl);
    atomicMin(&s_min, mn);
    __syncthreads();
    if (threadIdx.x == 0) {
        results[k].status = 0;
        results[k].produced = d.out_len;
        results[k].newlines = s_nl;
        results[k].min_byte = s_min;
        results[k].end_bit = 0;
    }
}

// ---- record-base scan ---------------------------------------------------------------
// One CTA: per-chunk record counts -> exclusive prefix sum -> ParseDesc.  Replaces the
// host-side bookkeeping BatchedFASTQ does with its ConcurrentQueue (BatchedFASTQ.cs:69).
constexpr int kScanThreads = 1024;
__global__ void __launch_bounds__(kScanThreads) pp_scan_kernel(const ChunkDesc *__restrict__ descs,
                                                               const ChunkResult *__restrict__ results,
                                                               const int64_t *__restrict__ exact_counts, int n,
                                                               uint32_t strict, int64_t capacity,
                                                               ParseDesc *__restrict__ pdesc,
                                                               ParseOut *__restrict__ pout,
                                                               ScanTotals *__restrict__ totals)
{
    __shared__ int64_t s_part[kScanThreads];
    __shared__ int64_t s_bytes[kScanThreads];
    __shared__ int64_t s_scanned[kScanThreads];
    __shared__ int s_first_k, s_status, s_exact;
    const int t = (int)threadIdx.x;
    if (t == 0) { s_first_k = n; s_status = 0; s_exact = 0; }
    __syncthreads();
    const int per = (n + kScanThreads - 1) / kScanThreads;
    const int lo = min(n, t * per), hi = min(n, lo + per);
    int64_t sum = 0, bytes = 0, scanned = 0;
    int first_status = 0, first_k = n, nexact = 0;
    for (int k = lo; k < hi; k++) {
        const ChunkDesc d = descs[k];
        const ChunkResult r = results[k];
        uint32_t cnt = 0, skip = 0, exact = 0;
        if (r.status == 0) {
            cnt = (d.prefix_nl + r.newlines) >> 2;
            if (exact_counts && exact_counts[k] >= 0) { cnt = (uint32_t)exact_counts[k]; exact = 1; }
            if (strict && d.prefix_nl >= 4u) skip = min(cnt, d.prefix_nl >> 2);
        } else if (first_status == 0) {
            first_status = r.status;
            first_k = k;
        }
        ParseDesc p;
        p.data_off = d.slot_off + d.lead_len - d.prefix_len;
        p.total = r.status == 0 ? d.prefix_len + r.produced : 0u;
        p.rec_count = cnt - skip;
        p.skip = skip;
        p.exact = exact;
        p.rec_base = sum;  // local to this thread's range; rebased below
        pdesc[k] = p;
        ParseOut o;  // the parse kernels fill these in (several CTAs per chunk: start from a clean slate)
        o.parse_end = 0;
        o.flags = 0;
        o.newlines = 0;
        o.records = cnt - skip;
        pout[k] = o;
        sum += cnt - skip;
        bytes += r.produced;
        scanned += p.total;
        nexact += (int)exact;
    }
    s_part[t] = sum;
    s_bytes[t] = bytes;
    s_scanned[t] = scanned;
    if (first_status) atomicMin(&s_first_k, first_k);
    if (nexact) atomicAdd(&s_exact, nexact);
    __syncthreads();
    if (first_status && s_first_k == first_k) s_status = first_status;
    if (t == 0) {
        // n <= a few 100k chunks: a serial pass over 1024 partials is negligible
        int64_t acc = 0, b = 0, sc = 0;
        for (int i = 0; i < kScanThreads; i++) {
            const int64_t v = s_part[i];
            s_part[i] = acc;
            acc += v;
            b += s_bytes[i];
            sc += s_scanned[i];
        }
        totals->total_records = acc;
        totals->total_bytes = b;
        totals->scanned_bytes = sc;
        totals->overflow = acc > capacity ? 1 : 0;
    }
    __syncthreads();
    const int64_t base = s_part[t];
    for (int k = lo; k < hi; k++) pdesc[k].rec_base += base;
    if (t == 0) {
        totals->first_status = s_status;
        totals->exact_chunks = s_exact;
    }
}

// ---- on-device consumer: histogram of the sequence lines -------------------------------------
// One CTA per chunk (grid-stride), one warp per record: the 32 lanes stream the record's sequence
// line; the five letters FASTQ sequences are made of live in registers, anything else goes through
// a shared-memory histogram.  Replaces the host loops of Decompressor/Program.cs:51-52.
__global__ void __launch_bounds__(256) pp_base_histogram_kernel(const uint8_t *__restrict__ slots,
                                                                const ParseDesc *__restrict__ pdesc, int n,
                                                                const uint32_t *__restrict__ lines, int64_t stride,
                                                                unsigned long long *__restrict__ counts)
{
    __shared__ unsigned int s_hist[256];
    for (int i = (int)threadIdx.x; i < 256; i += (int)blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const int lane = (int)(threadIdx.x & 31u), warp = (int)(threadIdx.x >> 5), nwarps = (int)(blockDim.x >> 5);
    unsigned long long cA = 0, cC = 0, cG = 0, cT = 0, cN = 0;
    for (int k = (int)blockIdx.x; k < n; k += (int)gridDim.x) {
        const ParseDesc d = pdesc[k];
        const uint8_t *data = slots + d.data_off;
        const uint32_t *l1 = lines + stride + d.rec_base, *l2 = lines + 2 * stride + d.rec_base;
        for (uint32_t r = (uint32_t)warp; r < d.rec_count; r += (uint32_t)nwarps) {
            const uint32_t a = l1[r], b = l2[r] - 1u;  // sequence = [l1, l2 - 1): without its '\n'
            for (uint32_t p = a + (uint32_t)lane; p < b; p += 32u) {
                const unsigned int c = data[p];
                if (c == 'A') cA++;
                else if (c == 'C') cC++;
                else if (c == 'G') cG++;
                else if (c == 'T') cT++;
                else if (c == 'N') cN++;
                else atomicAdd(&s_hist[c], 1u);
            }
        }
    }
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) {
        cA += __shfl_xor_sync(0xffffffffu, cA, sft);
        cC += __shfl_xor_sync(0xffffffffu, cC, sft);
        cG += __shfl_xor_sync(0xffffffffu, cG, sft);
        cT += __shfl_xor_sync(0xffffffffu, cT, sft);
        cN += __shfl_xor_sync(0xffffffffu, cN, sft);
    }
    if (lane == 0) {
        if (cA) atomicAdd(&counts['A'], cA);
        if (cC) atomicAdd(&counts['C'], cC);
        if (cG) atomicAdd(&counts['G'], cG);
        if (cT) atomicAdd(&counts['T'], cT);
        if (cN) atomicAdd(&counts['N'], cN);
    }
    __syncthreads();
    for (int i = (int)threadIdx.x; i < 256; i += (int)blockDim.x)
        if (s_hist[i]) atomicAdd(&counts[i], (unsigned long long)s_hist[i]);
}

// ---- on-device consumer: records whose sequence line contains a pattern -----------------------
// One warp per record (grid-stride over chunks): lane i tests the start positions a+i, a+i+32, ...
// of the sequence line against the pattern (held in shared memory up to 256 bytes).  Replaces the host loop of
// Benchmark/Naive.cs:167-180 (`record.Sequence.Contains(pattern)`, ordinal).
constexpr int kMaxPattern = 256;
__global__ void __launch_bounds__(256) pp_pattern_count_kernel(const uint8_t *__restrict__ slots,
                                                               const ParseDesc *__restrict__ pdesc, int n,
                                                               const uint32_t *__restrict__ lines, int64_t stride,
                                                               const uint8_t *__restrict__ pattern, int plen,
                                                               unsigned long long *__restrict__ count)
{
    __shared__ uint8_t s_pat[kMaxPattern];
    const bool in_smem = plen <= kMaxPattern;  // longer patterns are compared straight from global memory
    if (in_smem)
        for (int i = (int)threadIdx.x; i < plen; i += (int)blockDim.x) s_pat[i] = pattern[i];
    __syncthreads();
    const uint8_t *pat = in_smem ? s_pat : pattern;
    const int lane = (int)(threadIdx.x & 31u), warp = (int)(threadIdx.x >> 5), nwarps = (int)(blockDim.x >> 5);
    unsigned long long hits = 0;
    for (int k = (int)blockIdx.x; k < n; k += (int)gridDim.x) {
        const ParseDesc d = pdesc[k];
        const uint8_t *data = slots + d.data_off;
        const uint32_t *l1 = lines + stride + d.rec_base, *l2 = lines + 2 * stride + d.rec_base;
        for (uint32_t r = (uint32_t)warp; r < d.rec_count; r += (uint32_t)nwarps) {
            const uint32_t a = l1[r], b = l2[r] - 1u;  // sequence = [l1, l2 - 1): without its '\n'
            bool found = false;
            if (b - a >= (uint32_t)plen) {
                const uint32_t last = b - (uint32_t)plen;  // last start position
                for (uint32_t s = a + (uint32_t)lane; s <= last && !found; s += 32u) {
                    int j = 0;
                    while (j < plen && data[s + (uint32_t)j] == pat[j]) j++;
                    found = j == plen;
                }
            }
            if (__any_sync(0xffffffffu, found) && lane == 0) hits++;
        }
    }
    if (lane == 0 && hits) atomicAdd(count, hits);
}

cudaError_t launch_pattern_count(const uint8_t *slots, const ParseDesc *pdesc, int n, const uint32_t *lines,
                                 int64_t line_stride, const uint8_t *pattern, int plen, unsigned long long *count,
                                 int sm_count, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(count, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess || n <= 0) return e;
    if (plen < 1) return cudaErrorInvalidValue;
    const int grid = n < sm_count * 8 ? n : sm_count * 8;
    pp_pattern_count_kernel<<<grid, 256, 0, st>>>(slots, pdesc, n, lines, line_stride, pattern, plen, count);
    return cudaGetLastError();
}

cudaError_t launch_base_histogram(const uint8_t *slots, const ParseDesc *pdesc, int n, const uint32_t *lines,
                                  int64_t line_stride, unsigned long long *counts, int sm_count, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(counts, 0, 256 * sizeof(unsigned long long), st);
    if (e != cudaSuccess || n <= 0) return e;
    const int grid = n < sm_count * 8 ? n : sm_count * 8;
    pp_base_histogram_kernel<<<grid, 256, 0, st>>>(slots, pdesc, n, lines, line_stride, counts);
    return cudaGetLastError();
}

// ---- launchers ------------------------------------------------------------------------
cudaError_t launch_bytes_stats(const uint8_t *slots, const ChunkDesc *descs, ChunkResult *results, int n,
                               cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    pp_bytes_stats_kernel<<<n, 256, 0, st>>>(slots, descs, results, n);
    return cudaGetLastError();
}
cudaError_t launch_scan(const ChunkDesc *descs, const ChunkResult *results, const int64_t *exact_counts, int n,
                        uint32_t strict, int64_t capacity, ParseDesc *pdesc, ParseOut *pout, ScanTotals *totals,
                        cudaStream_t st)
{
    pp_scan_kernel<<<1, kScanThreads, 0, st>>>(descs, results, exact_counts, n, strict, capacity, pdesc, pout, totals);
    return cudaGetLastError();
}
uint32_t parse_tile_bytes() { return (uint32_t)kTileBytesP; }

// Resident CTAs of the parse kernel per SM on the current device (also opts the kernel into its
// dynamic shared memory size).  Called once per device context: the attribute lives in the
// device's primary context, and the answer is kept in the pp_ctx, not in a process-wide cache.
int parse_max_ctas_per_sm()
{
    int per_sm = 0;
    if (cudaFuncSetAttribute(pp_parse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ParseSm)) != cudaSuccess)
        return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pp_parse_kernel, kParseThreads, sizeof(ParseSm)) != cudaSuccess)
        return 0;
    return per_sm;
}

// tile_base: n+1 entries (device); max_tiles: largest tile count of one chunk;
// work: total_tiles u64 tile states followed by the u32 ticket
cudaError_t launch_parse(const uint8_t *slots, const ParseDesc *pdesc, int n, const uint32_t *tile_base,
                         uint32_t total_tiles, uint32_t max_tiles, uint32_t *lines, int64_t line_stride,
                         ParseOut *pout, const ScanTotals *totals, unsigned long long *work, int sm_count,
                         int per_sm, cudaStream_t st)
{
    if (n <= 0 || total_tiles == 0) return cudaSuccess;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    cudaError_t e = cudaMemsetAsync(work, 0, ((size_t)total_tiles + 1) * sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    uint32_t *ticket = reinterpret_cast<uint32_t *>(work + total_tiles);
    // round-robin over the chunks unless very uneven chunks would waste most tickets
    const uint64_t rr = (uint64_t)max_tiles * (uint64_t)n;
    int order = rr <= 4ull * total_tiles && rr < 0xffffffffull ? 1 : 0;
    if (const char *e = getenv("PPB200_PARSE_ORDER")) order = atoi(e) ? order : 0;  // 0 forces linear tickets
    const uint32_t n_tickets = order ? (uint32_t)rr : total_tiles;
    // the look-back needs every CTA of the grid resident
    const uint32_t resident = (uint32_t)sm_count * (uint32_t)per_sm;
    const uint32_t grid = n_tickets < resident ? n_tickets : resident;
    pp_parse_kernel<<<grid, kParseThreads, sizeof(ParseSm), st>>>(slots, pdesc, n, tile_base, n_tickets, order, lines, line_stride,
                                                                  pout, totals, work, ticket);
    return cudaGetLastError();
}
cudaError_t launch_exact_count(const uint8_t *slots, const ChunkDesc *descs, const ChunkResult *results,
                               const ParseOut *pout, int n, int64_t *exact_counts, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    pp_exact_count_kernel<<<(n + 31) / 32, 32, 0, st>>>(slots, descs, results, pout, n, exact_counts);
    return cudaGetLastError();
}
cudaError_t launch_exact_emit(const uint8_t *slots, const ParseDesc *pdesc, int n, uint32_t *lines,
                              int64_t line_stride, ParseOut *pout, const ScanTotals *totals, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    pp_exact_emit_kernel<<<(n + 31) / 32, 32, 0, st>>>(slots, pdesc, n, lines, line_stride, pout, totals);
    return cudaGetLastError();
}

}  // namespace pp

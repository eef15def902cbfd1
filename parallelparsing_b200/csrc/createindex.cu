// GPU CreateIndex: Core.BuildDeflateIndex (Decompressor/Core.cs:14-131) without the
// serial inflate(Z_BLOCK) pass.
//
// The reference inflates the whole file on one thread, stops at every block end (Core.cs:64, :98),
// counts '@' bytes (Core.cs:86) and drops a checkpoint — Input/Bits, the last 32 KB of output, the bytes
// since the last '@' — whenever more than chunksize-8 of them have gone by (Core.cs:105-109).  Here:
//
//   1. SCAN     pp_blockscan_kernel (blockscan_core.cuh) finds every block start and its output offset
//               without producing a byte: all the Z_BLOCK stops.
//   2. DECODE   consecutive blocks are grouped into segments and every segment is inflated by the
//               product's inflate kernel (dual-output variant: one Huffman decode, the tokens resolved
//               twice) into TWO outputs, each against a 32 KB dictionary that holds no data but
//               a code of its own positions: A[i] = i mod 256, B[i] = (i/256 + 1 + i mod 256) mod 256.  A
//               byte that comes out equal in both runs never touched the dictionary: it is a literal of
//               this segment (or a copy of one) and final.  A byte that differs is a copy, through any
//               number of hops, of dictionary position i = a + 256 ((b - a - 1) mod 256): its value is
//               byte i of the 32 KB of output in front of the segment, whatever they turn out to be.
//   3. CHAIN    the true 32 KB window behind every segment: each segment's tail is a function of the
//               window in front of it, and the functions are composed by a parallel scan over the segments.
//   4. RESOLVE  every differing byte of run A is replaced by its window byte (all segments in parallel);
//               run A now is the inflated file.
//   5. COUNT    per block: the number of '@', the first and the last one, the largest gap between two.
//   6. the host walks the block list exactly as Core.cs:98-125 does (index.cpp: index_plan_points) and
//      a gather kernel collects the windows and offsets of the points it chose.
//
// The gzip trailer is checked like zlib does (ISIZE and CRC-32, the latter computed on the device).
// Single-member gzip only; anything else answers PP_E_UNSUPPORTED and the host
// pp_index_create remains the general path.
#include <cuda_runtime.h>
#include <zlib.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

#include "blockscan_core.cuh"
#include "index.hpp"
#include "kernels.cuh"
#include "ppb200.h"

extern "C" int pp_internal_ctx_device(const pp_ctx *ctx, int *device, int *sm_count, cudaStream_t *stream);
extern "C" int pp_internal_ctx_inflate(pp_ctx *ctx, int n_chunks, pp::InflateLaunch *cfg);
extern "C" void pp_internal_ctx_lock(pp_ctx *ctx, int lock);

namespace pp {

constexpr uint32_t kWin = PP_WINSIZE;
constexpr uint32_t kPiece = 65536;   // bytes of one RESOLVE work item
constexpr uint32_t kCrcPart = 4096;  // bytes one thread runs the CRC over
constexpr uint32_t kCrcPiece = 256u * kCrcPart;  // ... and one CTA
constexpr uint32_t kNone = 0xffffffffu;

// One decode segment: a run of consecutive deflate blocks.
struct CiSeg {
    uint64_t a_off;    // its output in the slots buffer, run A
    uint64_t b_off;    // ... run B
    uint64_t out_off;  // offset of its first byte in the stream's output
    uint32_t out_len;
    uint32_t pad;
};
struct CiBlkIn {
    uint64_t addr;  // slots offset of the block's first output byte (run A)
    uint32_t len;
    uint32_t pad;
};
struct CiBlkOut {
    uint32_t ats, first, last, maxgap;
};
struct CiCopy {
    uint64_t src;  // offset in the stream's output
    uint64_t dst;  // offset in the gather buffer
    uint32_t len;
    uint32_t pad;
};

__device__ __forceinline__ uint32_t ci_pos(uint32_t a, uint32_t b) { return a + 256u * ((b - a - 1u) & 0xffu); }

// ---- CHAIN -----------------------------------------------------------------------------------------
// The window behind segment s is a function of the window in front of it: M_s[j] = a final byte, or "byte p
// of the window in front" (0x8000 | p) — for the segment's last min(len, 32768) bytes straight from the two
// runs, for the rest (a segment shorter than a window) the old window sliding down.  Functions of this
// kind compose — (M_s o M_{s-1})[j] = M_s[j] if final, else M_{s-1}[p] — and composition is associative,
// so the windows behind ALL segments come out of an inclusive scan over s: log2(S) rounds of gathers
// (Hillis-Steele), every round over all segments in parallel, instead of a walk through the segments in
// order.  What still points in front of segment 0 afterwards points at nothing and reads as zero (RESOLVE
// flags the bytes for which that is an error).
__device__ __forceinline__ uint32_t ci_ld_unaligned(const uint8_t *base, int64_t off)
{
    const uint8_t *p = base + off;
    const uintptr_t q = (uintptr_t)p & ~(uintptr_t)3;
    const uint32_t sh = 8u * (uint32_t)((uintptr_t)p & 3u);
    const uint32_t w0 = *reinterpret_cast<const uint32_t *>(q), w1 = *reinterpret_cast<const uint32_t *>(q + 4);
    return __funnelshift_r(w0, w1, sh);
}

// maps[s][j], j = 0 .. 32767 (u16): grid.x = S, every thread four entries at a time
__global__ void __launch_bounds__(256)
    pp_ci_tailmap_kernel(const CiSeg *__restrict__ segs, const uint8_t *__restrict__ slots, uint16_t *__restrict__ maps,
                         uint32_t *bad)
{
    const int s = (int)blockIdx.x;
    const CiSeg g = segs[s];
    const uint32_t n = g.out_len < kWin ? g.out_len : kWin;
    const uint8_t *pa = slots + g.a_off, *pb = slots + g.b_off;
    uint2 *dst = reinterpret_cast<uint2 *>(maps + (size_t)s * kWin);
    for (uint32_t w = threadIdx.x; w < kWin / 4u; w += blockDim.x) {
        const uint32_t j = 4u * w;
        const int64_t i0 = (int64_t)g.out_len + (int64_t)j - (int64_t)kWin;  // segment offset of window byte j
        uint32_t a = 0, b = 0;
        if (i0 > -4) {  // the word holds segment bytes (what lies in front of the segment is its lead area: ignored)
            a = ci_ld_unaligned(pa, i0);
            b = ci_ld_unaligned(pb, i0);
        }
        uint32_t e[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t jb = j + (uint32_t)q;
            if (jb >= kWin - n) {
                const uint32_t av = (a >> (8 * q)) & 0xffu, bv = (b >> (8 * q)) & 0xffu;
                if (av == bv) e[q] = av;
                else {
                    const uint32_t pos = ci_pos(av, bv);
                    if (pos >= kWin) atomicOr(bad, 1u);
                    e[q] = pos >= kWin ? 0u : 0x8000u | pos;
                }
            } else e[q] = 0x8000u | (jb + n);  // the old window slides
        }
        dst[w] = make_uint2(e[0] | (e[1] << 16), e[2] | (e[3] << 16));
    }
}

// One scan round: dst[s] = src[s] o src[s-d]  (s >= d), dst[s] = src[s] (s < d).  grid.x = S.
__global__ void __launch_bounds__(256)
    pp_ci_compose_kernel(const uint16_t *__restrict__ src, uint16_t *__restrict__ dst, int d)
{
    const int s = (int)blockIdx.x;
    const uint4 *in = reinterpret_cast<const uint4 *>(src + (size_t)s * kWin);
    uint4 *out = reinterpret_cast<uint4 *>(dst + (size_t)s * kWin);
    const uint16_t *prev = s >= d ? src + (size_t)(s - d) * kWin : nullptr;
    for (uint32_t w = threadIdx.x; w < kWin / 8u; w += blockDim.x) {
        uint4 v = in[w];
        if (prev && ((v.x | v.y | v.z | v.w) & 0x80008000u)) {
            uint32_t x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; q++) {
                uint32_t lo = x[q] & 0xffffu, hi = x[q] >> 16;
                if (lo & 0x8000u) lo = prev[lo & 0x7fffu];
                if (hi & 0x8000u) hi = prev[hi & 0x7fffu];
                x[q] = lo | (hi << 16);
            }
            v = make_uint4(x[0], x[1], x[2], x[3]);
        }
        out[w] = v;
    }
}

// The windows as bytes: wall[s][j] = maps[s][j], what still points in front of the stream = 0.
__global__ void __launch_bounds__(256) pp_ci_wall_kernel(const uint16_t *__restrict__ maps, uint8_t *__restrict__ wall)
{
    const int s = (int)blockIdx.x;
    const uint2 *in = reinterpret_cast<const uint2 *>(maps + (size_t)s * kWin);
    uint32_t *out = reinterpret_cast<uint32_t *>(wall + (size_t)s * kWin);
    for (uint32_t w = threadIdx.x; w < kWin / 4u; w += blockDim.x) {
        const uint2 v = in[w];
        const uint32_t e[4] = {v.x & 0xffffu, v.x >> 16, v.y & 0xffffu, v.y >> 16};
        uint32_t r = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) r |= ((e[q] & 0x8000u) ? 0u : (e[q] & 0xffu)) << (8 * q);
        out[w] = r;
    }
}

// ---- RESOLVE ---------------------------------------------------------------------------------------
__device__ __forceinline__ int ci_piece_seg(const uint32_t *piece_base, int S, uint32_t piece)
{
    int lo = 0, hi = S;  // piece_base[lo] <= piece < piece_base[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (piece_base[mid] <= piece) lo = mid;
        else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256)
    pp_ci_resolve_kernel(const CiSeg *__restrict__ segs, const uint32_t *__restrict__ piece_base, int S, uint32_t npieces,
                         uint8_t *slots, const uint8_t *__restrict__ wall, uint32_t *bad)
{
    for (uint32_t piece = blockIdx.x; piece < npieces; piece += gridDim.x) {
        const int s = ci_piece_seg(piece_base, S, piece);
        const CiSeg g = segs[s];
        const uint32_t from = (piece - piece_base[s]) * kPiece;
        const uint32_t to = from + kPiece < g.out_len ? from + kPiece : g.out_len;
        const uint8_t *w = s ? wall + (size_t)(s - 1) * kWin : nullptr;
        // window positions below this hold nothing (the stream is younger than 32 KB): zlib's
        // "invalid distance too far back"
        const uint32_t valid_from = g.out_off >= kWin ? 0u : kWin - (uint32_t)g.out_off;
        uint4 *A = reinterpret_cast<uint4 *>(slots + g.a_off);
        const uint4 *B = reinterpret_cast<const uint4 *>(slots + g.b_off);
        const bool fast = w != nullptr && valid_from == 0u;  // a whole window in front of the segment: every position is legal
        // four 16-byte groups per thread and trip: eight independent loads in flight
        for (uint32_t i0 = from / 16u + threadIdx.x; i0 * 16u < to; i0 += 4u * blockDim.x) {
            uint4 av[4], bv4[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t i = i0 + (uint32_t)u * blockDim.x;
                if (i * 16u < to) {
                    av[u] = A[i];
                    bv4[u] = B[i];
                } else av[u] = bv4[u] = make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t i = i0 + (uint32_t)u * blockDim.x;
                const uint4 a = av[u], b = bv4[u];
                if (a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w) continue;  // sixteen final bytes
                uint32_t aw[4] = {a.x, a.y, a.z, a.w};
                const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
                if (fast && i * 16u + 16u <= g.out_len) {
                    // four bytes at a time: hi = (b - a - 1) per byte is the position's high byte, a its low byte
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const uint32_t eq = __vcmpeq4(aw[q], bw[q]);  // 0xff in every byte that is final
                        if (eq == 0xffffffffu) continue;
                        const uint32_t hi = __vsub4(__vsub4(bw[q], aw[q]), 0x01010101u);
                        if (hi & ~eq & 0x80808080u) atomicOr(bad, 2u);  // a position >= 32768: no such dictionary byte
                        uint32_t r = aw[q];
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            if ((eq >> (8 * k)) & 1u) continue;
                            const uint32_t pos = __byte_perm(aw[q], hi, (uint32_t)((4 + k) << 4 | k)) & 0x7fffu;
                            r = __byte_perm(r, (uint32_t)w[pos], k == 0 ? 0x3214 : k == 1 ? 0x3240 : k == 2 ? 0x3410 : 0x4210);
                        }
                        aw[q] = r;
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        if (aw[q] == bw[q]) continue;
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const uint32_t x0 = (aw[q] >> (8 * k)) & 0xffu, x1 = (bw[q] >> (8 * k)) & 0xffu;
                            if (x0 == x1 || i * 16u + 4u * q + k >= g.out_len) continue;
                            const uint32_t pos = ci_pos(x0, x1);
                            uint32_t x = 0;
                            if (pos >= kWin || pos < valid_from || !w) atomicOr(bad, 2u);
                            else x = w[pos];
                            aw[q] = (aw[q] & ~(0xffu << (8 * k))) | (x << (8 * k));
                        }
                    }
                }
                A[i] = make_uint4(aw[0], aw[1], aw[2], aw[3]);
            }
        }
    }
}

// ---- COUNT -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    pp_ci_count_kernel(const CiBlkIn *__restrict__ in, int nb, const uint8_t *__restrict__ slots, CiBlkOut *__restrict__ out)
{
    __shared__ uint32_t s_cnt[256], s_first[256], s_last[256], s_gap[256];
    const uint32_t t = threadIdx.x;
    for (int blk = (int)blockIdx.x; blk < nb; blk += (int)gridDim.x) {
        const CiBlkIn q = in[blk];
        const uint64_t origin = q.addr & ~15ull;
        const uint32_t head = (uint32_t)(q.addr - origin);
        const uint32_t groups = (uint32_t)(((uint64_t)head + q.len + 15u) / 16u);
        const uint32_t per = (groups + 255u) / 256u;
        const uint32_t g0 = t * per < groups ? t * per : groups, g1 = g0 + per < groups ? g0 + per : groups;
        const uint4 *src = reinterpret_cast<const uint4 *>(slots + origin);
        uint32_t cnt = 0, first = kNone, last = kNone, gap = 0;
        for (uint32_t g = g0; g < g1; g++) {
            const uint4 v = src[g];
            const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int wi = 0; wi < 4; wi++) {
                const uint32_t x = wv[wi] ^ 0x40404040u;
                uint32_t m = ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x | 0x7f7f7f7fu);  // 0x80 in every byte that is '@'
                while (m) {
                    const uint32_t k = (uint32_t)(__ffs((int)m) - 1) >> 3;
                    m &= m - 1u;
                    const uint32_t rel = 16u * g + 4u * (uint32_t)wi + k - head;  // wraps for bytes in front of the block
                    if (rel < q.len) {
                        if (first == kNone) first = rel;
                        else if (rel - last > gap) gap = rel - last;
                        last = rel;
                        cnt++;
                    }
                }
            }
        }
        s_cnt[t] = cnt;
        s_first[t] = first;
        s_last[t] = last;
        s_gap[t] = gap;
        __syncthreads();
        if (t == 0) {
            CiBlkOut r = {0u, kNone, kNone, 0u};
            for (int i = 0; i < 256; i++) {
                if (!s_cnt[i]) continue;
                if (r.first == kNone) r.first = s_first[i];
                else if (s_first[i] - r.last > r.maxgap) r.maxgap = s_first[i] - r.last;
                if (s_gap[i] > r.maxgap) r.maxgap = s_gap[i];
                r.last = s_last[i];
                r.ats += s_cnt[i];
            }
            out[blk] = r;
        }
        __syncthreads();
    }
}

// ---- GATHER ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    pp_ci_gather_kernel(const CiCopy *__restrict__ items, int n, const CiSeg *__restrict__ segs, int S,
                        const uint8_t *__restrict__ slots, uint8_t *__restrict__ dst)
{
    __shared__ int s_seg;
    for (int it = (int)blockIdx.x; it < n; it += (int)gridDim.x) {
        const CiCopy c = items[it];
        if (threadIdx.x == 0) {
            int lo = 0, hi = S;  // the last segment that starts at or before c.src
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (segs[mid].out_off <= c.src) lo = mid;
                else hi = mid;
            }
            s_seg = lo;
        }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < c.len; i += blockDim.x) {
            const uint64_t g = c.src + i;
            int s = s_seg;
            while (s + 1 < S && g >= segs[s].out_off + segs[s].out_len) s++;
            dst[c.dst + i] = slots[segs[s].a_off + (g - segs[s].out_off)];
        }
        __syncthreads();
    }
}

// ---- CRC-32 (the gzip trailer's, RFC 1952 8.) -------------------------------------------------------
// CRC-32 is linear: crc(X ++ Y) = crc(X) * x^(8|Y|) + crc(Y) over GF(2)[x] mod the CRC polynomial (bit-reflected
// representation, x^0 = 0x80000000 — the identity zlib's crc32_combine is built on).  Every thread runs the
// byte-wise table method over 4 KB of a 1 MB piece, shifts its value by the bytes that
// follow it in the piece, and the CTA xors the 256 contributions into the piece's CRC; the host chains the
// pieces with crc32_combine.
__device__ __forceinline__ uint32_t ci_mulmod(uint32_t a, uint32_t b)
{
    uint32_t p = 0;
    for (int i = 31; i >= 0; i--) {
        if ((a >> i) & 1u) p ^= b;
        b = (b & 1u) ? (b >> 1) ^ 0xedb88320u : b >> 1;
    }
    return p;
}
__device__ __forceinline__ uint32_t ci_xpow(uint32_t n)  // x^n
{
    uint32_t r = 0x80000000u, base = 0x40000000u;
    while (n) {
        if (n & 1u) r = ci_mulmod(r, base);
        base = ci_mulmod(base, base);
        n >>= 1;
    }
    return r;
}

__global__ void __launch_bounds__(256)
    pp_ci_crc_kernel(const CiSeg *__restrict__ segs, const uint32_t *__restrict__ piece_base, int S, uint32_t npieces,
                     const uint8_t *__restrict__ slots, uint32_t *__restrict__ crcs)
{
    __shared__ uint32_t tab[4][256], shift_parts[256], red[8];  // tab[k][i]: byte i followed by k zero bytes
    const uint32_t t = threadIdx.x;
    {
        uint32_t c = t;
        for (int k = 0; k < 8; k++) c = (c & 1u) ? 0xedb88320u ^ (c >> 1) : c >> 1;
        tab[0][t] = c;
        shift_parts[t] = ci_xpow(8u * kCrcPart * t);  // t parts further on
    }
    __syncthreads();
    for (int k = 1; k < 4; k++) {
        const uint32_t p = tab[k - 1][t];
        tab[k][t] = (p >> 8) ^ tab[0][p & 0xffu];
    }
    __syncthreads();
    for (uint32_t piece = blockIdx.x; piece < npieces; piece += gridDim.x) {
        const int s = ci_piece_seg(piece_base, S, piece);
        const CiSeg g = segs[s];
        const uint32_t pfrom = (piece - piece_base[s]) * kCrcPiece;
        const uint32_t plen = g.out_len - pfrom < kCrcPiece ? g.out_len - pfrom : kCrcPiece;
        const uint32_t K = (plen + kCrcPart - 1u) / kCrcPart, r = plen - kCrcPart * (K - 1u);  // r: bytes of the last part
        uint32_t c = 0;
        if (t < K) {
            const uint32_t from = pfrom + kCrcPart * t, n = t + 1u < K ? kCrcPart : r;
            const uint4 *src = reinterpret_cast<const uint4 *>(slots + g.a_off + from);
            c = 0xffffffffu;
            const uint32_t n16 = n / 16u;
            for (uint32_t i0 = 0; i0 < n16; i0 += 4u) {  // four loads in flight in front of the dependent CRC steps
                uint4 v4[4];
#pragma unroll
                for (int u = 0; u < 4; u++) v4[u] = i0 + (uint32_t)u < n16 ? src[i0 + (uint32_t)u] : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (i0 + (uint32_t)u >= n16) break;
                    const uint32_t wv[4] = {v4[u].x, v4[u].y, v4[u].z, v4[u].w};
#pragma unroll
                    for (int q = 0; q < 4; q++) {  // four bytes per step: four independent lookups
                        c ^= wv[q];
                        c = tab[3][c & 0xffu] ^ tab[2][(c >> 8) & 0xffu] ^ tab[1][(c >> 16) & 0xffu] ^ tab[0][c >> 24];
                    }
                }
            }
            const uint8_t *p = slots + g.a_off + from;
            for (uint32_t i = n & ~15u; i < n; i++) c = tab[0][(c ^ p[i]) & 0xffu] ^ (c >> 8);
            c ^= 0xffffffffu;
            // K-2-t whole parts and the last part's r bytes follow this one
            if (t + 1u < K) c = ci_mulmod(ci_mulmod(c, shift_parts[K - 2u - t]), r == kCrcPart ? shift_parts[1] : ci_xpow(8u * r));
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) c ^= __shfl_xor_sync(0xffffffffu, c, o);
        if ((t & 31u) == 0) red[t >> 5] = c;
        __syncthreads();
        if (t == 0) {
            uint32_t x = 0;
            for (int i = 0; i < 8; i++) x ^= red[i];
            crcs[piece] = x;
        }
        __syncthreads();
    }
}

struct Dev {  // stream-ordered allocation from the device's pool (release threshold lifted by pp_open)
    void *p = nullptr;
    cudaStream_t st = nullptr;
    ~Dev()
    {
        if (p) cudaFreeAsync(p, st);
    }
    cudaError_t alloc(size_t n, cudaStream_t stream)
    {
        st = stream;
        return cudaMallocAsync(&p, n ? n : 1, stream);
    }
    template <class T> T *as() const { return (T *)p; }
};

}  // namespace pp

#define CKI(call)                                                                                                   \
    do {                                                                                                            \
        cudaError_t e_ = (call);                                                                                    \
        if (e_ == cudaErrorMemoryAllocation) { /* not enough device memory: the caller has the host pass */          \
            (void)cudaGetLastError();                                                                               \
            return PP_MEM_ERROR;                                                                                    \
        }                                                                                                           \
        if (e_ != cudaSuccess) {                                                                                    \
            fprintf(stderr, "ppb200: %s failed: %s (%s:%d)\n", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return PP_E_CUDA;                                                                                       \
        }                                                                                                           \
    } while (0)

namespace {
struct CtxLock {
    pp_ctx *c;
    explicit CtxLock(pp_ctx *ctx) : c(ctx) { pp_internal_ctx_lock(c, 1); }
    ~CtxLock() { pp_internal_ctx_lock(c, 0); }
};
struct Ev {
    cudaEvent_t e = nullptr;
    Ev() { cudaEventCreate(&e); }
    ~Ev() { if (e) cudaEventDestroy(e); }
};
}  // namespace

namespace {
struct HostTrace {  // PPB200_CI_TRACE=1: where the host side of the last stage spends its time (stderr)
    bool on = getenv("PPB200_CI_TRACE") != nullptr;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void lap(const char *what)
    {
        if (!on) return;
        const auto n = std::chrono::steady_clock::now();
        fprintf(stderr, "ppb200 create_gpu: %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
        t = n;
    }
};
}  // namespace

static int create_gpu(pp_ctx *ctx, const uint8_t *gz, size_t gz_len, uint32_t chunksize, uint32_t flags, pp_index *ix,
                      pp_create_stats *stt)
{
    using namespace pp;
    const size_t hdr = gzip_member_header_len(gz, gz_len);
    if (!hdr || gz_len < hdr + 8) return PP_DATA_ERROR;
    int device = 0, sm_count = 0;
    cudaStream_t st = nullptr;
    if (pp_internal_ctx_device(ctx, &device, &sm_count, &st) != PP_OK) return PP_E_ARG;
    CtxLock lock(ctx);
    CKI(cudaSetDevice(device));
    Ev ev[9];
    int nev = 0;
    auto mark = [&]() { cudaEventRecord(ev[nev++].e, st); };

    // 0. the file
    Dev comp;
    const size_t comp_base = gz_len & ~(size_t)15, comp_pad = 4096 + 16;
    CKI(comp.alloc(comp_base + comp_pad, st));
    mark();  // 0
    CKI(cudaMemsetAsync(comp.as<uint8_t>() + comp_base, 0, comp_pad, st));
    CKI(cudaMemcpyAsync(comp.p, gz, gz_len, cudaMemcpyHostToDevice, st));
    mark();  // 1

    // 1. SCAN
    std::vector<ppinf::BlockRec> chain;
    uint64_t land = 0, total_out = 0;
    float scan_kernel_ms = 0.f;
    int passes = 0;
    int rc = scan_blocks_resident(device, sm_count, st, comp.as<uint8_t>(), gz, gz_len, hdr, 0, chain, land, total_out,
                                  scan_kernel_ms, passes);
    if (rc == PP_BUF_ERROR) return PP_E_UNSUPPORTED;  // blocks of < 64 compressed bytes on average: the record areas overflow
    if (rc != PP_OK) return rc;
    mark();  // 2
    const uint64_t end_byte = (land + 7u) >> 3;
    if (end_byte + 8u > gz_len) return PP_DATA_ERROR;     // no room for the trailer: zlib runs out of input (Core.cs:42-45)
    if (end_byte + 8u != gz_len) return PP_E_UNSUPPORTED;  // more members (or garbage) behind the first: Core.cs:116-121
    const uint8_t *trailer = gz + end_byte;
    const uint32_t want_crc = (uint32_t)trailer[0] | ((uint32_t)trailer[1] << 8) | ((uint32_t)trailer[2] << 16) | ((uint32_t)trailer[3] << 24);
    const uint32_t want_len = (uint32_t)trailer[4] | ((uint32_t)trailer[5] << 8) | ((uint32_t)trailer[6] << 16) | ((uint32_t)trailer[7] << 24);
    if (want_len != (uint32_t)total_out) return PP_DATA_ERROR;  // "incorrect length check"
    const size_t nb = chain.size();
    if (!nb) return PP_DATA_ERROR;

    // 2. segments of consecutive blocks
    const uint64_t target_bits = 8u * std::min<uint64_t>(512u << 10, std::max<uint64_t>(32u << 10, gz_len / 1500u));
    std::vector<CiSeg> segs;
    std::vector<uint32_t> seg_first;  // first block of each segment
    for (size_t i = 0; i < nb;) {
        size_t j = i + 1;
        const uint64_t out0 = chain[i].out;
        while (j < nb && chain[j].bit - chain[i].bit < target_bits && chain[j + 1 < nb ? j + 1 : j].out - out0 < (1u << 30)) j++;
        const uint64_t out1 = j < nb ? chain[j].out : total_out;
        if (out1 - out0 >= (1ull << 31)) return PP_E_UNSUPPORTED;  // one block of more than 2 GB
        CiSeg g{};
        g.out_off = out0;
        g.out_len = (uint32_t)(out1 - out0);
        segs.push_back(g);
        seg_first.push_back((uint32_t)i);
        i = j;
    }
    const int S = (int)segs.size();
    // run A's slots, then run B's in the same layout: chunk s of run B = chunk s of run A + slot_delta
    std::vector<ChunkDesc> descs((size_t)S);
    uint64_t slot_off = 512;
    for (int s = 0; s < S; s++) {
        ChunkDesc &d = descs[(size_t)s];
        d.in_bit = chain[seg_first[(size_t)s]].bit;
        d.in_limit = gz_len;
        d.slot_off = slot_off;
        d.lead_src = 0;
        d.lead_len = kWin;
        d.out_len = segs[(size_t)s].out_len;
        d.prefix_len = d.prefix_nl = 0;
        segs[(size_t)s].a_off = slot_off + kWin;
        slot_off += ((uint64_t)kWin + d.out_len + 1u + 127u) & ~127ull;
    }
    const uint64_t slot_delta = slot_off - 512;
    for (int s = 0; s < S; s++) segs[(size_t)s].b_off = segs[(size_t)s].a_off + slot_delta;
    slot_off += slot_delta;
    std::vector<uint32_t> piece_base((size_t)S + 1);
    uint32_t npieces = 0;
    for (int s = 0; s < S; s++) {
        piece_base[(size_t)s] = npieces;
        npieces += (segs[(size_t)s].out_len + kPiece - 1u) / kPiece;
    }
    piece_base[(size_t)S] = npieces;
    std::vector<uint32_t> crc_base((size_t)S + 1);
    uint32_t ncrc = 0;
    for (int s = 0; s < S; s++) {
        crc_base[(size_t)s] = ncrc;
        ncrc += (segs[(size_t)s].out_len + kCrcPiece - 1u) / kCrcPiece;
    }
    crc_base[(size_t)S] = ncrc;

    // 3. DECODE twice
    std::vector<uint8_t> coded(2 * kWin);
    for (uint32_t i = 0; i < kWin; i++) {
        coded[i] = (uint8_t)i;
        coded[kWin + i] = (uint8_t)((i >> 8) + 1u + (i & 0xffu));
    }
    Dev slots, lead, d_descs, d_res, d_segs, d_pb, d_cb, d_bad, wall;
    CKI(slots.alloc(slot_off + 512, st));
    CKI(lead.alloc(2 * kWin, st));
    CKI(d_descs.alloc(sizeof(ChunkDesc) * descs.size(), st));
    CKI(d_res.alloc(sizeof(ChunkResult) * descs.size(), st));
    CKI(d_segs.alloc(sizeof(CiSeg) * (size_t)S, st));
    CKI(d_pb.alloc(sizeof(uint32_t) * piece_base.size(), st));
    CKI(d_cb.alloc(sizeof(uint32_t) * crc_base.size(), st));
    CKI(d_bad.alloc(sizeof(uint32_t), st));
    CKI(wall.alloc((size_t)S * kWin, st));
    CKI(cudaMemcpyAsync(lead.p, coded.data(), coded.size(), cudaMemcpyHostToDevice, st));
    CKI(cudaMemcpyAsync(d_descs.p, descs.data(), sizeof(ChunkDesc) * descs.size(), cudaMemcpyHostToDevice, st));
    CKI(cudaMemcpyAsync(d_segs.p, segs.data(), sizeof(CiSeg) * (size_t)S, cudaMemcpyHostToDevice, st));
    CKI(cudaMemcpyAsync(d_pb.p, piece_base.data(), sizeof(uint32_t) * piece_base.size(), cudaMemcpyHostToDevice, st));
    CKI(cudaMemcpyAsync(d_cb.p, crc_base.data(), sizeof(uint32_t) * crc_base.size(), cudaMemcpyHostToDevice, st));
    CKI(cudaMemsetAsync(d_bad.p, 0, sizeof(uint32_t), st));
    InflateLaunch cfg;
    if (pp_internal_ctx_inflate(ctx, S, &cfg) != PP_OK) return PP_E_ARG;
    mark();  // 3
    CKI(launch_inflate_dual(d_descs.as<ChunkDesc>(), S, comp.as<uint8_t>(), gz_len, slots.as<uint8_t>(), lead.as<uint8_t>(),
                            d_res.as<ChunkResult>(), cfg, slot_delta, kWin, st));
    mark();  // 4

    // 4. CHAIN, RESOLVE
    {
        Dev maps0, maps1;
        CKI(maps0.alloc((size_t)S * kWin * sizeof(uint16_t), st));
        CKI(maps1.alloc((size_t)S * kWin * sizeof(uint16_t), st));
        uint16_t *src = maps0.as<uint16_t>(), *dst = maps1.as<uint16_t>();
        pp_ci_tailmap_kernel<<<S, 256, 0, st>>>(d_segs.as<CiSeg>(), slots.as<uint8_t>(), src, d_bad.as<uint32_t>());
        CKI(cudaGetLastError());
        for (int d = 1; d < S; d *= 2) {
            pp_ci_compose_kernel<<<S, 256, 0, st>>>(src, dst, d);
            CKI(cudaGetLastError());
            std::swap(src, dst);
        }
        pp_ci_wall_kernel<<<S, 256, 0, st>>>(src, wall.as<uint8_t>());
        CKI(cudaGetLastError());
    }
    mark();  // 5
    const int wide = sm_count * 8;
    if (npieces) {
        pp_ci_resolve_kernel<<<(int)std::min<uint32_t>(npieces, (uint32_t)wide), 256, 0, st>>>(
            d_segs.as<CiSeg>(), d_pb.as<uint32_t>(), S, npieces, slots.as<uint8_t>(), wall.as<uint8_t>(), d_bad.as<uint32_t>());
        CKI(cudaGetLastError());
    }
    mark();  // 6

    // 5. COUNT and CRC
    std::vector<CiBlkIn> bin(nb);
    {
        int s = 0;
        for (size_t i = 0; i < nb; i++) {
            while (s + 1 < S && seg_first[(size_t)s + 1] <= i) s++;
            const uint64_t o1 = i + 1 < nb ? chain[i + 1].out : total_out;
            bin[i].addr = segs[(size_t)s].a_off + (chain[i].out - segs[(size_t)s].out_off);
            bin[i].len = (uint32_t)(o1 - chain[i].out);
            bin[i].pad = 0;
        }
    }
    Dev d_bin, d_bout, d_crc;
    CKI(d_bin.alloc(sizeof(CiBlkIn) * nb, st));
    CKI(d_bout.alloc(sizeof(CiBlkOut) * nb, st));
    CKI(d_crc.alloc(sizeof(uint32_t) * (size_t)ncrc, st));
    CKI(cudaMemcpyAsync(d_bin.p, bin.data(), sizeof(CiBlkIn) * nb, cudaMemcpyHostToDevice, st));
    pp_ci_count_kernel<<<(int)std::min<size_t>(nb, (size_t)wide), 256, 0, st>>>(d_bin.as<CiBlkIn>(), (int)nb, slots.as<uint8_t>(),
                                                                                 d_bout.as<CiBlkOut>());
    CKI(cudaGetLastError());
    if (ncrc) {
        pp_ci_crc_kernel<<<(int)std::min<uint32_t>(ncrc, (uint32_t)wide), 256, 0, st>>>(
            d_segs.as<CiSeg>(), d_cb.as<uint32_t>(), S, ncrc, slots.as<uint8_t>(), d_crc.as<uint32_t>());
        CKI(cudaGetLastError());
    }
    mark();  // 7
    HostTrace tr;
    std::vector<ChunkResult> res(descs.size());
    std::vector<CiBlkOut> bout(nb);
    std::vector<uint32_t> crcs((size_t)ncrc);
    uint32_t bad = 0;
    CKI(cudaMemcpyAsync(res.data(), d_res.p, sizeof(ChunkResult) * res.size(), cudaMemcpyDeviceToHost, st));
    CKI(cudaMemcpyAsync(bout.data(), d_bout.p, sizeof(CiBlkOut) * nb, cudaMemcpyDeviceToHost, st));
    CKI(cudaMemcpyAsync(crcs.data(), d_crc.p, sizeof(uint32_t) * crcs.size(), cudaMemcpyDeviceToHost, st));
    CKI(cudaMemcpyAsync(&bad, d_bad.p, sizeof bad, cudaMemcpyDeviceToHost, st));
    CKI(cudaStreamSynchronize(st));
    tr.lap("statistics to the host");
    for (size_t i = 0; i < res.size(); i++)
        if (res[i].status < 0 || res[i].produced != descs[i].out_len) return res[i].status < 0 ? res[i].status : PP_DATA_ERROR;
    if (bad) return PP_DATA_ERROR;  // a distance that reaches in front of the stream ("invalid distance too far back")
    {
        // CRC of the whole output from the pieces, in stream order
        uLong crc = crc32(0L, Z_NULL, 0);
        const uLong op_full = crc32_combine_gen((z_off_t)kCrcPiece);
        for (int s = 0; s < S; s++) {
            const uint32_t len = segs[(size_t)s].out_len;
            for (uint32_t p = crc_base[(size_t)s]; p < crc_base[(size_t)s + 1]; p++) {
                const uint32_t plen = std::min(kCrcPiece, len - (p - crc_base[(size_t)s]) * kCrcPiece);
                crc = plen == kCrcPiece ? crc32_combine_op(crc, crcs[p], op_full) : crc32_combine(crc, crcs[p], (z_off_t)plen);
            }
        }
        if ((uint32_t)crc != want_crc) return PP_DATA_ERROR;  // "incorrect data check"
    }
    tr.lap("crc combine");

    // 6. the points (Core.cs:98-125), then their windows and offsets
    std::vector<CiBlockStat> bs(nb);
    for (size_t i = 0; i < nb; i++) bs[i] = {chain[i].bit, chain[i].out, bout[i].ats, bout[i].first, bout[i].last, bout[i].maxgap};
    std::vector<CiPointPlan> plan;
    rc = index_plan_points(bs.data(), nb, total_out, gz_len, chunksize, flags, plan);
    if (rc != PP_OK) return rc;
    tr.lap("plan points");
    index_from_plan(ix, plan);
    tr.lap("size the index");
    const size_t np = plan.size();
    std::vector<CiCopy> items;
    uint64_t off_bytes = 0;
    for (size_t k = 0; k < np; k++) {
        const uint64_t out = (uint64_t)plan[k].output, have = std::min<uint64_t>(out, kWin);
        for (uint64_t c = 0; c < have; c += 4096) {  // the window: the last `have` bytes in front of the point
            CiCopy it{out - have + c, (uint64_t)k * kWin + (kWin - have) + c, (uint32_t)std::min<uint64_t>(4096, have - c), 0};
            items.push_back(it);
        }
        const uint64_t olen = (uint64_t)(plan[k].output - plan[k].off_from);
        for (uint64_t c = 0; c < olen; c += 4096) {
            CiCopy it{(uint64_t)plan[k].off_from + c, (uint64_t)np * kWin + off_bytes + c, (uint32_t)std::min<uint64_t>(4096, olen - c), 0};
            items.push_back(it);
        }
        off_bytes += olen;
    }
    Dev d_items, d_gather;
    CKI(d_items.alloc(sizeof(CiCopy) * items.size(), st));
    CKI(d_gather.alloc(np * (size_t)kWin + off_bytes, st));
    CKI(cudaMemsetAsync(d_gather.p, 0, np * (size_t)kWin, st));
    if (!items.empty()) {
        CKI(cudaMemcpyAsync(d_items.p, items.data(), sizeof(CiCopy) * items.size(), cudaMemcpyHostToDevice, st));
        pp_ci_gather_kernel<<<(int)std::min<size_t>(items.size(), (size_t)wide), 256, 0, st>>>(
            d_items.as<CiCopy>(), (int)items.size(), d_segs.as<CiSeg>(), S, slots.as<uint8_t>(), d_gather.as<uint8_t>());
        CKI(cudaGetLastError());
    }
    CKI(cudaMemcpyAsync(ix->windows, d_gather.p, np * (size_t)kWin, cudaMemcpyDeviceToHost, st));
    if (off_bytes) CKI(cudaMemcpyAsync(ix->offsets.data(), d_gather.as<uint8_t>() + np * (size_t)kWin, off_bytes, cudaMemcpyDeviceToHost, st));
    mark();  // 8
    CKI(cudaStreamSynchronize(st));
    tr.lap("gather + D2H");
    if (stt) {
        float ms[8] = {};
        for (int i = 0; i < 8; i++) cudaEventElapsedTime(&ms[i], ev[i].e, ev[i + 1].e);
        stt->h2d_ms = ms[0];
        stt->scan_ms = ms[1];
        stt->scan_kernel_ms = scan_kernel_ms;
        stt->plan_ms = ms[2];
        stt->inflate_ms = ms[3];
        stt->chain_ms = ms[4];
        stt->resolve_ms = ms[5];
        stt->count_crc_ms = ms[6];
        stt->gather_ms = ms[7];
        cudaEventElapsedTime(&stt->total_ms, ev[0].e, ev[8].e);
        stt->blocks = (int64_t)nb;
        stt->segments = S;
        stt->scan_passes = passes;
        stt->points = (int32_t)np;
        stt->total_out = (int64_t)total_out;
    }
    return PP_OK;
}

extern "C" int pp_index_create_gpu(pp_ctx *ctx, const uint8_t *gz, size_t gz_len, uint32_t chunksize, uint32_t flags,
                                   pp_index **out, pp_create_stats *stats)
{
    if (!ctx || !out || !gz) return PP_E_ARG;
    *out = nullptr;
    if (stats) memset(stats, 0, sizeof *stats);
    try {
        std::unique_ptr<pp_index> ix(new pp_index());
        const int rc = create_gpu(ctx, gz, gz_len, chunksize, flags, ix.get(), stats);
        if (rc != PP_OK) return rc;
        *out = ix.release();
        return PP_OK;
    } catch (...) {
        return PP_MEM_ERROR;
    }
}

// GPU-assisted CreateIndex, first slice: pp_scan_blocks — the start bit and the
// output offset of every deflate block of a gzip member, i.e. the stops Core.BuildDeflateIndex gets
// from inflate(Z_BLOCK) (Decompressor/Core.cs:64, :98) and the only places a checkpoint may sit.
// Kernel + stitching; see blockscan_core.cuh for the method.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "blockscan_core.cuh"
#include "kernels.cuh"
#include "ppb200.h"

namespace pp {

using ppinf::BlockRec;
using ppinf::ScanSegIn;
using ppinf::ScanSegOut;

__global__ void __launch_bounds__(ppinf::kMaxThreads, 1)
    pp_blockscan_kernel(const ScanSegIn *__restrict__ segs, int nseg, const uint8_t *__restrict__ comp, uint64_t comp_bytes,
                        uint32_t shift_bytes, BlockRec *__restrict__ recs, ScanSegOut *__restrict__ outs)
{
    extern __shared__ __align__(128) uint8_t pp_smem_raw[];
    ppinf::Sm sm;
    ppinf::sm_carve(sm, pp_smem_raw, (int)blockDim.x);
    if (threadIdx.x == 0) {
        ppinf::mbar_init(sm.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        sm.u[24] = 0;
        sm.u[25] = 0;
    }
    __syncthreads();
    uint32_t stage_phase = 0;
    for (int s = (int)blockIdx.x; s < nseg; s += (int)gridDim.x)
        ppinf::scan_segment(sm, segs[s], comp, comp_bytes, 8ull * shift_bytes, recs, outs[s], stage_phase);
}

static constexpr int kScanThreads = 512;

cudaError_t launch_blockscan(const ScanSegIn *segs, int nseg, const uint8_t *comp, uint64_t comp_bytes, BlockRec *recs,
                             ScanSegOut *outs, int sm_count, cudaStream_t st)
{
    if (nseg <= 0) return cudaSuccess;
    static bool attr_set[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const size_t smem = ppinf::sm_bytes_for(kScanThreads);
    if (!attr_set[dev & 63]) {  // idempotent: the same value every time, so a race between contexts is harmless
        e = cudaFuncSetAttribute(pp_blockscan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_set[dev & 63] = true;
    }
    const uint32_t shift = (uint32_t)((uintptr_t)comp & 15u);
    const int grid = std::min(nseg, sm_count * 2);
    pp_blockscan_kernel<<<grid, kScanThreads, smem, st>>>(segs, nseg, comp - shift, comp_bytes + shift, shift, recs, outs);
    return cudaGetLastError();
}

}  // namespace pp

// ---- gzip member header (RFC 1952 2.3): offset of the first deflate byte, or 0 -----------------
static size_t gzip_header_len(const uint8_t *gz, size_t n)
{
    if (n < 18 || gz[0] != 0x1f || gz[1] != 0x8b || gz[2] != 8) return 0;
    const uint8_t flg = gz[3];
    size_t p = 10;
    if (flg & 4) {  // FEXTRA
        if (p + 2 > n) return 0;
        p += 2 + ((size_t)gz[p] | ((size_t)gz[p + 1] << 8));
    }
    for (int f = 0; f < 2; f++)   // FNAME, FCOMMENT: zero-terminated
        if (flg & (f ? 16 : 8)) {
            while (p < n && gz[p]) p++;
            p++;
        }
    if (flg & 2) p += 2;  // FHCRC
    return p < n ? p : 0;
}

// ---- host walk of one fixed-codes block (RFC 1951 3.2.6), for bridging seams the search cannot see ----
// p: first bit after the 3 header bits.  Counts the bytes the block produces; false: invalid or too long.
static bool host_walk_fixed(const uint8_t *gz, size_t n, uint64_t &p, uint64_t &out, uint64_t max_bits)
{
    static const uint16_t lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint8_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint8_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    const uint64_t end = std::min<uint64_t>((uint64_t)n * 8u, p + max_bits);
    auto bit = [&](uint64_t i) -> uint32_t { return (gz[i >> 3] >> (i & 7u)) & 1u; };
    for (;;) {
        if (p + 9u + 5u + 5u + 13u > end) return false;
        uint32_t code = 0, sym = 0xffffu;
        for (int len = 1; len <= 9 && sym == 0xffffu; len++) {
            code = (code << 1) | bit(p++);
            if (len == 7 && code <= 0x17u) sym = 256u + code;
            else if (len == 8 && code >= 0x30u && code <= 0xbfu) sym = code - 0x30u;
            else if (len == 8 && code >= 0xc0u && code <= 0xc7u) sym = 280u + (code - 0xc0u);
            else if (len == 9 && code >= 0x190u) sym = 144u + (code - 0x190u);
        }
        if (sym == 0xffffu) return false;
        if (sym < 256u) { out++; continue; }
        if (sym == 256u) return true;
        if (sym > 285u) return false;
        uint32_t length = lbase[sym - 257u];
        for (uint32_t k = 0; k < lext[sym - 257u]; k++) length += bit(p++) << k;
        uint32_t dc = 0;
        for (int k = 0; k < 5; k++) dc = (dc << 1) | bit(p++);
        if (dc > 29u) return false;
        p += dext[dc];
        out += length;
    }
}

// Test hook (not part of the ABI): the host walk of the fixed-codes block whose header starts at `bit`.
extern "C" int pp_internal_walk_fixed(const uint8_t *gz, size_t n, uint64_t bit, uint64_t *next_bit, uint64_t *out_bytes)
{
    uint64_t p = bit + 3u, o = 0;
    if (!host_walk_fixed(gz, n, p, o, ~0ull >> 1)) return PP_DATA_ERROR;
    *next_bit = p;
    *out_bytes = o;
    return PP_OK;
}

#define CKS(call)                                                                                 \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            fprintf(stderr, "ppb200: %s failed: %s (%s:%d)\n", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            rc = PP_E_CUDA;                                                                       \
            goto done;                                                                            \
        }                                                                                         \
    } while (0)

extern "C" int pp_internal_ctx_device(const pp_ctx *ctx, int *device, int *sm_count, cudaStream_t *stream);

namespace pp {

// The scan proper on a stream that is already resident in device memory (d_comp: the whole gzip file,
// at least 4096 + 16 zero bytes behind it).  Appends every block start to `chain` (output offsets
// global), sets land (bit after the final block) and total_out.
int scan_blocks_resident(int device, int sm_count, cudaStream_t st, const uint8_t *d_comp, const uint8_t *h_gz, size_t gz_len,
                         size_t hdr, int64_t segment_bytes, std::vector<BlockRec> &chain, uint64_t &land_out, uint64_t &total_out,
                         float &ms_total, int &npass)
{
    // Every segment pays one search (about half a block of bit positions probed) before its walk, and a CTA
    // walks one segment: the default is one segment per resident CTA — as few searches as keep every SM busy.
    if (segment_bytes <= 0)
        segment_bytes = (int64_t)std::min<uint64_t>(8u << 20, std::max<uint64_t>(128u << 10, gz_len / (2u * (uint64_t)std::max(sm_count, 1))));
    if (segment_bytes < 4096) segment_bytes = 4096;
    int rc = PP_OK;
    ScanSegIn *d_segs = nullptr;
    ScanSegOut *d_outs = nullptr;
    BlockRec *d_recs = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    try {
        CKS(cudaSetDevice(device));
        const uint64_t comp_bytes = gz_len;   // the gzip trailer (8 bytes) is simply never reached
        const uint64_t stream_bits = (uint64_t)gz_len * 8u;
        const int nseg = (int)std::max<uint64_t>(1, (gz_len - hdr + (uint64_t)segment_bytes - 1) / (uint64_t)segment_bytes);
        const uint32_t rec_cap = (uint32_t)std::max<int64_t>(256, segment_bytes / 64);  // blocks average >= 64 compressed bytes, else PP_BUF_ERROR
        CKS(cudaMallocAsync((void **)&d_segs, sizeof(ScanSegIn) * (size_t)nseg, st));
        CKS(cudaMallocAsync((void **)&d_outs, sizeof(ScanSegOut) * (size_t)nseg, st));
        CKS(cudaMallocAsync((void **)&d_recs, sizeof(BlockRec) * (size_t)nseg * rec_cap, st));
        CKS(cudaEventCreate(&e0));
        CKS(cudaEventCreate(&e1));
        // pass 0: every segment searches (but the first) and walks
        std::vector<ScanSegIn> segs((size_t)nseg);
        std::vector<ScanSegOut> outs((size_t)nseg);
        std::vector<std::vector<BlockRec>> recs((size_t)nseg);  // per segment, only the records it wrote
        std::vector<BlockRec> stage;
        for (int s = 0; s < nseg; s++) {
            ScanSegIn &g = segs[(size_t)s];
            g.start_bit = s ? ((uint64_t)hdr + (uint64_t)s * (uint64_t)segment_bytes) * 8u : (uint64_t)hdr * 8u;
            g.end_bit = s + 1 < nseg ? ((uint64_t)hdr + (uint64_t)(s + 1) * (uint64_t)segment_bytes) * 8u : stream_bits;
            g.search = s ? 1u : 0u;
            g.rec_off = (uint32_t)s * rec_cap;
            g.rec_cap = rec_cap;
            // bit 0: the search may take a block with BFINAL set — where the stream's last block can start: the last segment
            // and whatever lies within 1 MiB of the end (a longer final block costs a re-walk of its seam, no more)
            g.pad = (s + 1 == nseg || g.end_bit + (8u << 20) >= stream_bits) ? 1u : 0u;
        }
        std::vector<int> todo((size_t)nseg);
        for (int s = 0; s < nseg; s++) todo[(size_t)s] = s;
        std::vector<ScanSegIn> batch;
        int verified = 0;            // segments [0, verified) are in the chain
        uint64_t land = 0, out_base = 0;
        bool final_seen = false;
        for (;;) {
            // run the segments in `todo`
            batch.clear();
            for (int s : todo) batch.push_back(segs[(size_t)s]);
            CKS(cudaMemcpyAsync(d_segs, batch.data(), sizeof(ScanSegIn) * batch.size(), cudaMemcpyHostToDevice, st));
            CKS(cudaEventRecord(e0, st));
            CKS(launch_blockscan(d_segs, (int)batch.size(), d_comp, comp_bytes, d_recs, d_outs, sm_count, st));
            CKS(cudaEventRecord(e1, st));
            std::vector<ScanSegOut> bo(batch.size());
            CKS(cudaMemcpyAsync(bo.data(), d_outs, sizeof(ScanSegOut) * batch.size(), cudaMemcpyDeviceToHost, st));
            CKS(cudaStreamSynchronize(st));
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            ms_total += ms;
            npass++;
            if (npass == 1) {
                // all segments, record areas at a regular pitch: ONE 2-D copy of the used head of every area
                uint32_t maxn = 0;
                for (const ScanSegOut &o : bo) maxn = std::max(maxn, o.nrec);
                if (maxn) {
                    stage.resize((size_t)maxn * bo.size());
                    CKS(cudaMemcpy2DAsync(stage.data(), sizeof(BlockRec) * maxn, d_recs, sizeof(BlockRec) * rec_cap,
                                          sizeof(BlockRec) * maxn, bo.size(), cudaMemcpyDeviceToHost, st));
                    CKS(cudaStreamSynchronize(st));
                }
                for (size_t i = 0; i < todo.size(); i++) {
                    outs[(size_t)todo[i]] = bo[i];
                    recs[(size_t)todo[i]].assign(stage.begin() + (size_t)i * maxn, stage.begin() + (size_t)i * maxn + bo[i].nrec);
                }
            } else {
                for (size_t i = 0; i < todo.size(); i++) {
                    const int s = todo[i];
                    outs[(size_t)s] = bo[i];
                    recs[(size_t)s].resize(bo[i].nrec);
                    if (bo[i].nrec)
                        CKS(cudaMemcpy(recs[(size_t)s].data(), d_recs + segs[(size_t)s].rec_off, sizeof(BlockRec) * bo[i].nrec,
                                       cudaMemcpyDeviceToHost));
                }
            }
            // stitch: extend the chain while each segment's walk starts where the chain landed
            todo.clear();
            while (verified < nseg && !final_seen) {
                const int s = verified;
                const ScanSegOut &o = outs[(size_t)s];
                const bool anchored = (s == 0 && segs[0].search == 0) || segs[(size_t)s].search == 0;
                if (o.status < 0 && o.status != -5 && (anchored || o.first_bit == land)) { rc = PP_DATA_ERROR; goto done; }
                if (o.status == -5) { rc = PP_BUF_ERROR; goto done; }
                bool bridged = false;
                if (s > 0 && !anchored && o.first_bit != land && land < o.first_bit && o.status >= 0 && h_gz) {
                    // Stored and fixed-codes blocks between where the chain landed and the first block the segment's
                    // search saw (the search sees dynamic headers only): the short fixed block and the empty stored
                    // block of a sync flush that parallel gzip writers leave at every seam.  They are walked here, on
                    // the host copy (up to 1 Mbit of fixed-codes data; beyond that the segment is walked again).
                    uint64_t p = land, add = 0;
                    size_t keep = chain.size();
                    while (p < o.first_bit) {
                        const uint64_t by = p >> 3;
                        const uint32_t h3 = (((uint32_t)h_gz[by] | ((uint32_t)(by + 1 < gz_len ? h_gz[by + 1] : 0) << 8)) >> (p & 7u)) & 7u;
                        if (h3 == 2u) {  // fixed codes, not final
                            uint64_t q = p + 3u, o2 = 0;
                            if (!host_walk_fixed(h_gz, gz_len, q, o2, 1u << 20)) break;
                            chain.push_back({p, out_base + add});
                            add += o2;
                            p = q;
                            continue;
                        }
                        if (h3 != 0u) break;  // not a non-final stored block
                        const uint64_t lb = (p + 3u + 7u) >> 3;
                        if (lb + 4u > gz_len) break;
                        const uint32_t len = (uint32_t)h_gz[lb] | ((uint32_t)h_gz[lb + 1] << 8);
                        const uint32_t nlen = (uint32_t)h_gz[lb + 2] | ((uint32_t)h_gz[lb + 3] << 8);
                        if ((len ^ 0xffffu) != nlen) break;
                        chain.push_back({p, out_base + add});
                        add += len;
                        p = (lb + 4u + len) * 8u;
                    }
                    if (p == o.first_bit) {
                        out_base += add;
                        land = p;
                        bridged = true;
                    } else chain.resize(keep);
                }
                if (s > 0 && !anchored && !bridged && o.first_bit != land) {
                    if (land >= segs[(size_t)s].end_bit) {   // the chain already walked past this whole segment
                        verified++;
                        continue;
                    }
                    // the seam does not close (a stored/fixed block the search cannot see, or a false positive):
                    // walk this segment again from where the chain landed, no search
                    segs[(size_t)s].start_bit = land;
                    segs[(size_t)s].search = 0;
                    todo.push_back(s);
                    break;
                }
                for (uint32_t r = 0; r < o.nrec; r++) {
                    BlockRec br = recs[(size_t)s][r];
                    br.out += out_base;
                    chain.push_back(br);
                }
                out_base += o.out_bytes;
                land = o.land_bit;
                if (o.status == 1) final_seen = true;
                verified++;
            }
            if (todo.empty()) break;
            if (npass > nseg + 2) { rc = PP_E_CUDA; goto done; }   // cannot happen: every pass verifies a segment
        }
        if (!final_seen) { rc = PP_DATA_ERROR; goto done; }        // the stream ended without a final block
        land_out = land;
        total_out = out_base;
    } catch (...) {
        rc = PP_MEM_ERROR;
    }
done:
    if (d_segs) cudaFreeAsync(d_segs, st);
    if (d_outs) cudaFreeAsync(d_outs, st);
    if (d_recs) cudaFreeAsync(d_recs, st);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    return rc;
}

size_t gzip_member_header_len(const uint8_t *gz, size_t n) { return gzip_header_len(gz, n); }

}  // namespace pp

extern "C" int pp_scan_blocks(pp_ctx *ctx, const uint8_t *gz, size_t gz_len, int64_t segment_bytes, int64_t *start_bits,
                              int64_t *out_offsets, int64_t cap, int64_t *count, int64_t *end_bit, int64_t *total_out,
                              float *kernel_ms, int32_t *passes)
{
    using namespace pp;
    if (!ctx || !gz || !count) return PP_E_ARG;
    *count = 0;
    const size_t hdr = gzip_header_len(gz, gz_len);
    if (!hdr) return PP_DATA_ERROR;
    int device = 0, sm_count = 0;
    cudaStream_t st = nullptr;
    if (pp_internal_ctx_device(ctx, &device, &sm_count, &st) != PP_OK) return PP_E_ARG;
    int rc = PP_OK;
    uint8_t *d_comp = nullptr;
    float ms_total = 0.f;
    int npass = 0;
    uint64_t land = 0, out_total = 0;
    std::vector<BlockRec> chain;      // accepted block starts, output offsets global
    try {
        CKS(cudaSetDevice(device));
        const size_t comp_base = gz_len & ~(size_t)15, comp_pad = 4096 + 16;   // zeroed tail: reads past the end see zeros
        CKS(cudaMallocAsync((void **)&d_comp, comp_base + comp_pad, st));   // the device's pool (threshold lifted by pp_open)
        CKS(cudaMemsetAsync(d_comp + comp_base, 0, comp_pad, st));
        CKS(cudaMemcpyAsync(d_comp, gz, gz_len, cudaMemcpyHostToDevice, st));
        rc = scan_blocks_resident(device, sm_count, st, d_comp, gz, gz_len, hdr, segment_bytes, chain, land, out_total,
                                  ms_total, npass);
        if (rc != PP_OK) goto done;
        *count = (int64_t)chain.size();
        for (int64_t i = 0; i < (int64_t)chain.size() && i < cap; i++) {
            if (start_bits) start_bits[i] = (int64_t)chain[(size_t)i].bit;
            if (out_offsets) out_offsets[i] = (int64_t)chain[(size_t)i].out;
        }
        if (end_bit) *end_bit = (int64_t)land;
        if (total_out) *total_out = (int64_t)out_total;
        if ((int64_t)chain.size() > cap && (start_bits || out_offsets)) rc = PP_BUF_ERROR;
    } catch (...) {
        rc = PP_MEM_ERROR;
    }
done:
    if (kernel_ms) *kernel_ms = ms_total;
    if (passes) *passes = npass;
    if (d_comp) cudaFreeAsync(d_comp, st);
    return rc;
}

// Host runtime behind the C ABI: device context, DecompressAll jobs,
// Decompress(checkpoint) and Parse on caller buffers.
//
// DecompressAll (BatchedFASTQ, Decompressor/BatchedFASTQ.cs:54-98) becomes a
// plan ("job") over a contiguous chunk range:
//   upload   one H2D copy of the compressed byte range [Input_first-1, Input_last)
//            (LazyFileReader.cs:63-69 reads the same bytes chunk by chunk) and one of
//            the checkpoint windows, both from pinned host memory;
//   execute  inflate kernel (one CTA per chunk) -> exact-count kernel (only chunks that
//            need the quirk-exact parser do work) -> record-base scan -> parse kernel;
//   download per-chunk results.
// Chunks are independent: a multi-GPU run gives every rank its own
// pp_ctx and a disjoint chunk range; there is no cross-GPU exchange.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "index.hpp"
#include "kernels.cuh"
#include "ppb200.h"

using namespace pp;

#define CK(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            fprintf(stderr, "ppb200: %s failed: %s (%s:%d)\n", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return PP_E_CUDA;                                                                     \
        }                                                                                         \
    } while (0)

// Device memory of a job comes from the device's stream-ordered pool (cudaMallocAsync) with the release
// threshold lifted (pp_open): after the first DecompressAll the next one's allocations are served from
// memory the previous one returned, without a trip to the driver (a cold cudaMalloc/cudaFree of a few GB
// of slots costs more than the decode itself).
static cudaError_t pool_alloc(void **p, size_t bytes, cudaStream_t st) { return cudaMallocAsync(p, bytes ? bytes : 1, st); }
template <class T> static cudaError_t pool_alloc(T **p, size_t bytes, cudaStream_t st) { return pool_alloc((void **)p, bytes, st); }
static void pool_free(void *p, cudaStream_t st)
{
    if (p) cudaFreeAsync(p, st);
}

struct pp_ctx;
static cudaError_t ctx_arena_get(pp_ctx *c, size_t bytes, uint8_t **out, size_t *cap);  // caller holds c->mu
static void ctx_arena_put(pp_ctx *c, uint8_t *p, size_t cap);                            // takes c->mu

// Small RAII helper for the single-call entry points.
struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, n ? n : 1); }
    template <class T> T *as() { return (T *)p; }
};

struct pp_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;       // kernels, and every copy of the plain (non-pipelined) paths
    cudaStream_t copy_stream = nullptr;  // H2D of a pipelined upload (overlaps the kernels)
    cudaStream_t d2h_stream = nullptr;   // D2H of a streamed download (overlaps the kernels)
    std::mutex mu;
    int sm_count = 0;
    // inflate launch geometries: few chunks -> large CTAs (one per SM), many chunks -> more CTAs per SM
    InflateLaunch wide, dense;
    uint32_t *d_map = nullptr;
    int *d_counter = nullptr;
    int parse_per_sm = 0;  // resident CTAs of the parse kernel per SM (its look-back needs a resident grid)
    // pinned + mapped host arenas of finished jobs, kept for the next job of this context: cudaHostAlloc and
    // cudaFreeHost are driver calls that map into every device of the process (1-2 ms each with several GPUs)
    std::vector<std::pair<uint8_t *, size_t>> arenas;
    const InflateLaunch &inflate_cfg(int n_chunks) const { return n_chunks <= wide.grid ? wide : dense; }
};

static int ctx_setup_inflate(pp_ctx *c)
{
    int t_wide = 1024, t_dense = 512;
    if (const char *e = getenv("PPB200_INFLATE_T")) {
        const int v = atoi(e);
        if (v >= 32 && v <= 1024 && (v & (v - 1)) == 0) t_wide = t_dense = v;  // a power of two: resolve tiles are 16 T bytes
    }
    // dynamic shared memory opt-in: once per context, for the largest CTA size (never per launch)
    CK(inflate_set_max_smem(std::max(t_wide, t_dense)));
    const int occ_w = inflate_max_ctas_per_sm(t_wide), occ_d = inflate_max_ctas_per_sm(t_dense);
    if (occ_w <= 0 || occ_d <= 0) {
        fprintf(stderr, "ppb200: inflate kernel does not fit an SM (threads %d/%d)\n", t_wide, t_dense);
        return PP_E_CUDA;
    }
    c->wide.threads = t_wide;
    c->wide.grid = c->sm_count * occ_w;
    c->dense.threads = t_dense;
    c->dense.grid = c->sm_count * occ_d;
    const size_t bytes = std::max(inflate_scratch_bytes(t_wide, c->wide.grid), inflate_scratch_bytes(t_dense, c->dense.grid));
    CK(cudaMalloc(&c->d_map, bytes));
    CK(cudaMalloc(&c->d_counter, sizeof(int)));
    c->wide.map = c->dense.map = c->d_map;
    c->wide.counter = c->dense.counter = c->d_counter;
    c->parse_per_sm = parse_max_ctas_per_sm();
    if (c->parse_per_sm < 1) {
        fprintf(stderr, "ppb200: parse kernel does not fit an SM\n");
        return PP_E_CUDA;
    }
    return PP_OK;
}

static cudaError_t ctx_arena_get(pp_ctx *c, size_t bytes, uint8_t **out, size_t *cap)
{
    size_t best = c->arenas.size();
    for (size_t i = 0; i < c->arenas.size(); i++)
        if (c->arenas[i].second >= bytes && (best == c->arenas.size() || c->arenas[i].second < c->arenas[best].second)) best = i;
    if (best != c->arenas.size()) {
        *out = c->arenas[best].first;
        *cap = c->arenas[best].second;
        c->arenas.erase(c->arenas.begin() + (long)best);
        return cudaSuccess;
    }
    size_t want = 64 << 10;
    while (want < bytes) want *= 2;
    void *p = nullptr;
    const cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocMapped);
    if (e != cudaSuccess) return e;
    *out = (uint8_t *)p;
    *cap = want;
    return cudaSuccess;
}
static void ctx_arena_put(pp_ctx *c, uint8_t *p, size_t cap)
{
    if (!p) return;
    uint8_t *drop = nullptr;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        c->arenas.emplace_back(p, cap);
        if (c->arenas.size() > 8) {  // keep the eight largest
            size_t small = 0;
            for (size_t i = 1; i < c->arenas.size(); i++)
                if (c->arenas[i].second < c->arenas[small].second) small = i;
            drop = c->arenas[small].first;
            c->arenas.erase(c->arenas.begin() + (long)small);
        }
    }
    if (drop) cudaFreeHost(drop);
}

static constexpr uint64_t kTile = 2048;  // padding granule of the compressed buffers
static inline uint64_t align_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

struct pp_job {
    pp_ctx *ctx = nullptr;
    const pp_index *ix = nullptr;
    int32_t first = 0, n = 0;
    uint32_t flags = 0;
    // host plan
    std::vector<ChunkDesc> descs;
    uint64_t comp_file_lo = 0;   // file offset of device compressed byte 0 (16 B aligned)
    uint64_t comp_copy = 0;      // bytes copied from the file
    uint64_t comp_alloc = 0;     // device bytes (tile multiple, + one spare tile)
    uint64_t slots_bytes = 0;
    uint64_t lead_bytes = 0;
    bool lead_direct = true;     // leads are exactly the index windows (one contiguous copy)
    uint8_t *h_lead = nullptr;   // pinned staging when !lead_direct
    int64_t rec_cap = 0;
    // device
    uint8_t *d_comp = nullptr, *d_lead = nullptr, *d_slots = nullptr;
    ChunkDesc *d_descs = nullptr;
    ChunkResult *d_results = nullptr;
    ParseDesc *d_pdesc = nullptr;
    ParseOut *d_pout = nullptr;
    ScanTotals *d_totals = nullptr;
    int64_t *d_exact = nullptr;
    uint32_t *d_lines = nullptr;
    uint32_t *d_tile_base = nullptr;        // first parse tile of every chunk (n+1 entries)
    unsigned long long *d_parse_work = nullptr;  // look-back tile states + ticket
    uint32_t total_tiles = 0, max_tiles = 0;
    cudaStream_t st_alloc = nullptr;  // stream the pooled allocations are ordered on (the context's)
    int device = 0;
    uint8_t *h_arena = nullptr;       // ONE pinned, mapped allocation behind every small host mirror below
    size_t h_arena_cap = 0;
    pp_ctx *arena_owner = nullptr;    // the context whose arena cache h_arena goes back to
    // pinned host mirrors
    ChunkResult *h_results = nullptr;
    ParseDesc *h_pdesc = nullptr;
    ParseOut *h_pout = nullptr;
    ScanTotals *h_totals = nullptr;
    bool have_results = false;
    bool zero_copy = false;
    const uint8_t *zc_comp = nullptr;  // device-visible alias of the caller's pinned gz buffer
    // pipelined upload (PP_JOB_PIPELINE): pieces of the compressed range go over a copy stream while the
    // inflate kernel runs; after every piece the host publishes the bytes in place (d_avail)
    bool pipeline = false;
    ppinf::ByteGate gate = {nullptr, 0, 0};  // see ByteGate
    unsigned long long *d_avail = nullptr;  // the gate's mark
    unsigned long long *h_marks = nullptr;  // pinned: mark values, one per column copy (+ reset, + "everything")
    int n_marks = 0;
    cudaEvent_t ev_reset = nullptr, ev_lead = nullptr, ev_exec_done = nullptr;
    bool exec_pending = false;
    // compact windows (PP_JOB_COMPACT_WINDOWS): the checkpoint windows cross PCIe zlib-compressed and a
    // pre-pass of the inflate kernel unpacks them straight into the slots' lead areas
    bool compact = false;
    uint64_t cwin_lo = 0, cwin_bytes = 0;   // range of the index's compact blob this job needs
    uint8_t *d_cwin = nullptr;
    ChunkDesc *d_wdescs = nullptr;
    ChunkResult *d_wresults = nullptr, *h_wresults = nullptr;
    // streamed download (pp_job_execute_to_host): done[k] = 1 (mapped pinned memory) when chunk k is final
    uint32_t *h_done = nullptr;
    cudaEvent_t ev[8] = {};
    pp_job_info info{};
};

// --------------------------------------------------------------------------- ctx

static int check_device(int device)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return PP_E_NO_DEVICE;
    if (device < 0 || device >= count) return PP_E_ARG;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PP_E_CUDA;
    if (prop.major != 10) {
        fprintf(stderr, "ppb200: device %d is sm_%d%d; this library is built for sm_100a only\n", device, prop.major,
                prop.minor);
        return PP_E_NO_DEVICE;
    }
    return PP_OK;
}

extern "C" int pp_internal_ctx_device(const pp_ctx *ctx, int *device, int *sm_count, cudaStream_t *stream)
{
    if (!ctx) return PP_E_ARG;
    if (device) *device = ctx->device;
    if (sm_count) *sm_count = ctx->sm_count;
    if (stream) *stream = ctx->stream;
    return PP_OK;
}

// pp_index_create_gpu (createindex.cu) drives the inflate kernel itself: the launch geometry for n chunks,
// and the context lock that serialises users of the context's token scratch.
extern "C" int pp_internal_ctx_inflate(pp_ctx *ctx, int n_chunks, pp::InflateLaunch *cfg)
{
    if (!ctx || !cfg) return PP_E_ARG;
    *cfg = ctx->inflate_cfg(n_chunks);
    return PP_OK;
}
extern "C" void pp_internal_ctx_lock(pp_ctx *ctx, int lock)
{
    if (!ctx) return;
    if (lock) ctx->mu.lock();
    else ctx->mu.unlock();
}

extern "C" void pp_internal_unpin_index(const pp_index *ix)
{
    if (ix && ix->pinned_base) {
        cudaHostUnregister(ix->pinned_base);
        ix->pinned_base = nullptr;
        ix->pinned_bytes = 0;
    }
}

extern "C" void pp_internal_unpin_cwin(const pp_index *ix)
{
    if (ix && ix->cwin_pinned) {
        cudaHostUnregister(ix->cwin_pinned);
        ix->cwin_pinned = nullptr;
    }
}

static void pin_index_cwin(const pp_index *ix)
{
    std::lock_guard<std::mutex> lk(ix->cw_mu);
    if (ix->cwin.empty() || ix->cwin_pinned == ix->cwin.data()) return;
    if (ix->cwin_pinned) cudaHostUnregister(ix->cwin_pinned);
    ix->cwin_pinned = nullptr;
    if (cudaHostRegister(ix->cwin.data(), ix->cwin.size(), cudaHostRegisterDefault) == cudaSuccess)
        ix->cwin_pinned = ix->cwin.data();
    else
        cudaGetLastError();  // stay pageable: copies still work, just slower (pull mode then refuses)
}

static void pin_index_windows(const pp_index *ix)
{
    const size_t bytes = (size_t)ix->count() * PP_WINSIZE;
    if (!ix->windows || bytes == 0) return;
    if (ix->pinned_base == ix->windows && ix->pinned_bytes >= bytes) return;
    pp_internal_unpin_index(ix);
    // pin the whole allocation so later points added in place stay covered
    const size_t cap_bytes = ix->win_cap * (size_t)PP_WINSIZE;
    if (cudaHostRegister(ix->windows, cap_bytes, cudaHostRegisterDefault) == cudaSuccess) {
        ix->pinned_base = ix->windows;
        ix->pinned_bytes = cap_bytes;
    } else {
        cudaGetLastError();  // stay pageable: copies still work, just slower
    }
}

extern "C" {

int pp_open(int32_t device, pp_ctx **out)
{
    if (!out) return PP_E_ARG;
    *out = nullptr;
    int rc = check_device(device);
    if (rc != PP_OK) return rc;
    pp_ctx *c = new (std::nothrow) pp_ctx();
    if (!c) return PP_MEM_ERROR;
    c->device = device;
    if (cudaSetDevice(device) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        return PP_E_CUDA;
    }
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    {
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;  // keep freed job memory for the next job
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        } else {
            cudaGetLastError();
        }
    }
    rc = ctx_setup_inflate(c);
    if (rc != PP_OK) {
        pp_close(c);
        return rc;
    }
    *out = c;
    return PP_OK;
}

void pp_close(pp_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (cudaStream_t st : {ctx->stream, ctx->copy_stream, ctx->d2h_stream})
        if (st) {
            cudaStreamSynchronize(st);
            cudaStreamDestroy(st);
        }
    cudaFree(ctx->d_map);
    cudaFree(ctx->d_counter);
    for (auto &a : ctx->arenas) cudaFreeHost(a.first);
    delete ctx;
}

int pp_host_alloc(size_t bytes, void **out)
{
    if (!out) return PP_E_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return PP_E_NO_DEVICE;
    // two spare TMA tiles behind the data: PP_JOB_ZEROCOPY kernels read whole tiles
    CK(cudaHostAlloc(out, bytes + 4 * kTile, cudaHostAllocPortable));
    memset((uint8_t *)*out + bytes, 0, 4 * kTile);
    return PP_OK;
}
void pp_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}
int pp_host_register(void *p, size_t bytes)
{
    if (!p) return PP_E_ARG;
    CK(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return PP_OK;
}
void pp_host_unregister(void *p)
{
    if (p) cudaHostUnregister(p);
}

// --------------------------------------------------------------------------- job

void pp_job_free(pp_job *j)
{
    if (!j) return;
    if (j->ctx) {
        std::lock_guard<std::mutex> lk(j->ctx->mu);
        cudaSetDevice(j->ctx->device);
        cudaStreamSynchronize(j->ctx->stream);
        cudaStreamSynchronize(j->ctx->copy_stream);
        cudaStreamSynchronize(j->ctx->d2h_stream);
    } else if (j->st_alloc) {
        cudaSetDevice(j->device);
        cudaStreamSynchronize(j->st_alloc);
    }
    cudaStream_t st = j->st_alloc;
    for (void *p : {(void *)j->d_comp, (void *)j->d_lead, (void *)j->d_slots, (void *)j->d_descs, (void *)j->d_results,
                    (void *)j->d_pdesc, (void *)j->d_pout, (void *)j->d_totals, (void *)j->d_exact, (void *)j->d_lines,
                    (void *)j->d_tile_base, (void *)j->d_parse_work, (void *)j->d_avail, (void *)j->d_cwin,
                    (void *)j->d_wdescs, (void *)j->d_wresults})
        pool_free(p, st);
    if (j->arena_owner) ctx_arena_put(j->arena_owner, j->h_arena, j->h_arena_cap);
    else cudaFreeHost(j->h_arena);
    for (cudaEvent_t e : {j->ev_reset, j->ev_lead, j->ev_exec_done})
        if (e) cudaEventDestroy(e);
    cudaFreeHost(j->h_lead);
    for (auto &e : j->ev)
        if (e) cudaEventDestroy(e);
    delete j;
}

static int job_alloc_lines(pp_job *j, int64_t cap)
{
    pool_free(j->d_lines, j->st_alloc);
    j->d_lines = nullptr;
    j->rec_cap = std::max<int64_t>(cap, 16);
    CK(pool_alloc(&j->d_lines, (size_t)j->rec_cap * 4 * sizeof(uint32_t), j->st_alloc));
    return PP_OK;
}

// Lay out one chunk: where its compressed bits start, where its slot lives, what goes in
// front of its output.  `lead_len` bytes of history precede the output; normally that is
// the 32 KB checkpoint window, of which from.offset is the tail (Core.cs:86-94,107).
static void count_prefix(const uint8_t *p, int32_t n, uint32_t *nl, bool *has_nul)
{
    uint32_t c = 0;
    bool z = false;
    for (int32_t i = 0; i < n; i++) {
        c += (p[i] == '\n');
        z |= (p[i] == 0);
    }
    *nl = c;
    *has_nul = z;
}

}  // extern "C"

static int job_create_inner(pp_job *j, pp_ctx *ctx, const pp_index *ix, size_t gz_len, int32_t first_chunk, int n)
{
    auto fail = [&](int rc) { return rc; };
    try {
        j->descs.resize((size_t)std::max(n, 1));
        std::vector<uint8_t> forced_exact((size_t)std::max(n, 1), 0);
        // compressed range: one byte before Input_first (the Bits live there) .. Input_last
        uint64_t lo = 0, hi = 0;
        if (n > 0) {
            const int64_t in0 = ix->input[(size_t)first_chunk];
            lo = in0 > 0 ? (uint64_t)(in0 - 1) : 0;
            lo &= ~(uint64_t)127;  // keep file offsets and device offsets congruent mod 128
            hi = std::min<uint64_t>((uint64_t)ix->input[(size_t)(first_chunk + n)], gz_len);
            if (hi < lo) hi = lo;
        }
        j->comp_file_lo = lo;
        j->comp_copy = hi - lo;
        j->comp_alloc = align_up(j->comp_copy, kTile) + 2 * kTile;
        // 512 spare bytes in front of the first slot: a decode with no history in front of it (the
        // window pre-pass) may touch up to 256 bytes before its output
        uint64_t slot_off = 512, lead_off = 0;
        int64_t scanned = 0;
        for (int k = 0; k < n; k++) {
            const int c = first_chunk + k;
            ChunkDesc &d = j->descs[(size_t)k];
            const int64_t in = ix->input[(size_t)c], out_lo = ix->output[(size_t)c], out_hi = ix->output[(size_t)c + 1];
            const int32_t bits = ix->bits[(size_t)c];
            const int64_t len = out_hi - out_lo;
            if (len < 0 || len > 0x7fffffffLL || (uint64_t)in < lo) return fail(PP_E_ARG);
            d.in_bit = ((uint64_t)in - lo) * 8u - (uint64_t)bits;
            d.in_limit = std::min<uint64_t>((uint64_t)ix->input[(size_t)c + 1], gz_len) - lo;
            d.out_len = (uint32_t)len;
            const int32_t ol = ix->off_len[(size_t)c];
            d.prefix_len = (uint32_t)ol;
            bool nul = false;
            count_prefix(ix->offset(c), ol, &d.prefix_nl, &nul);
            // from.offset must be the tail of the history placed before the output
            bool tail_ok = true;
            if (ol <= PP_WINSIZE) {
                tail_ok = ol == 0 || memcmp(ix->window(c) + PP_WINSIZE - ol, ix->offset(c), (size_t)ol) == 0;
                d.lead_len = PP_WINSIZE;
            } else {
                tail_ok = memcmp(ix->offset(c) + ol - PP_WINSIZE, ix->window(c), PP_WINSIZE) == 0;
                d.lead_len = (uint32_t)align_up((uint64_t)ol, 16);
                j->lead_direct = false;
            }
            if (!tail_ok) return fail(PP_E_ARG);  // index whose offset is not the stream tail: not produced by CreateIndex
            if (nul) forced_exact[(size_t)k] = 1;
            d.lead_src = lead_off;
            lead_off += d.lead_len;
            d.slot_off = slot_off;
            slot_off += align_up((uint64_t)d.lead_len + d.out_len + 1, 128);
            scanned += (int64_t)d.prefix_len + d.out_len;
        }
        j->slots_bytes = std::max<uint64_t>(slot_off, 1024);
        j->lead_bytes = std::max<uint64_t>(lead_off, 16);
        // compact windows: only when every lead is exactly the checkpoint window
        std::vector<ChunkDesc> wdescs;
        if (j->compact && (!j->lead_direct || n == 0)) j->compact = false;
        if (j->compact) {
            if (!index_build_compact_windows(ix)) return fail(PP_MEM_ERROR);
            pin_index_cwin(ix);
            j->cwin_lo = ix->cwin_off[(size_t)first_chunk];
            j->cwin_bytes = ix->cwin_off[(size_t)(first_chunk + n)] - j->cwin_lo;
            wdescs.resize((size_t)n);
            for (int k = 0; k < n; k++) {
                const int c = first_chunk + k;
                ChunkDesc &w = wdescs[(size_t)k];
                w = ChunkDesc{};
                w.in_bit = (ix->cwin_off[(size_t)c] - j->cwin_lo) * 8u + 16u;  // past the 2-byte zlib header
                w.in_limit = ix->cwin_off[(size_t)c + 1] - j->cwin_lo;
                w.slot_off = j->descs[(size_t)k].slot_off;                       // the chunk's lead area
                w.lead_src = ppinf::kLeadInPlace;
                w.lead_len = 0;
                w.out_len = PP_WINSIZE;
                j->descs[(size_t)k].lead_src = ppinf::kLeadInPlace;
            }
        }
        // parse tiles: chunk k's combined memory starts prefix_len bytes before its output
        std::vector<uint32_t> tile_base((size_t)n + 1, 0);
        {
            const uint64_t tb = parse_tile_bytes();
            uint64_t acc = 0;
            for (int k = 0; k < n; k++) {
                const ChunkDesc &d = j->descs[(size_t)k];
                const uint64_t head = (d.slot_off + d.lead_len - d.prefix_len) & 15u;
                tile_base[(size_t)k] = (uint32_t)acc;
                const uint64_t nt = (head + d.prefix_len + d.out_len + tb - 1) / tb;
                j->max_tiles = std::max<uint32_t>(j->max_tiles, (uint32_t)std::min<uint64_t>(nt, 0xffffffffu));
                acc += nt;
            }
            if (acc > 0xfffffff0ull) return fail(PP_E_ARG);
            tile_base[(size_t)n] = (uint32_t)acc;
            j->total_tiles = (uint32_t)acc;
        }

        CK(pool_alloc(&j->d_slots, j->slots_bytes, ctx->stream));
        CK(pool_alloc(&j->d_descs, sizeof(ChunkDesc) * (size_t)std::max(n, 1), ctx->stream));
        CK(pool_alloc(&j->d_results, sizeof(ChunkResult) * (size_t)std::max(n, 1), ctx->stream));
        CK(pool_alloc(&j->d_pdesc, sizeof(ParseDesc) * (size_t)std::max(n, 1), ctx->stream));
        CK(pool_alloc(&j->d_pout, sizeof(ParseOut) * (size_t)std::max(n, 1), ctx->stream));
        CK(pool_alloc(&j->d_totals, sizeof(ScanTotals), ctx->stream));
        CK(pool_alloc(&j->d_exact, sizeof(int64_t) * (size_t)std::max(n, 1), ctx->stream));
        CK(pool_alloc(&j->d_tile_base, sizeof(uint32_t) * ((size_t)n + 1), ctx->stream));
        CK(pool_alloc(&j->d_parse_work, sizeof(unsigned long long) * ((size_t)j->total_tiles + 1), ctx->stream));
        CK(cudaMemcpyAsync(j->d_tile_base, tile_base.data(), sizeof(uint32_t) * ((size_t)n + 1), cudaMemcpyHostToDevice,
                           ctx->stream));
        {
            // one pinned + mapped allocation for every small host mirror (each cudaHostAlloc is a driver call)
            const size_t n1 = (size_t)std::max(n, 1);
            const uint64_t kPiece = 8ull << 20;   // pipelined upload: pieces of 8 MB
            j->gate.total = j->comp_copy;
            j->n_marks = (int)((j->comp_copy + kPiece - 1) / kPiece) + 2;
            auto up = [](size_t v) { return (v + 63) & ~(size_t)63; };
            const size_t o_res = 0, o_pd = o_res + up(sizeof(ChunkResult) * n1), o_po = o_pd + up(sizeof(ParseDesc) * n1),
                         o_tot = o_po + up(sizeof(ParseOut) * n1), o_wres = o_tot + up(sizeof(ScanTotals)),
                         o_marks = o_wres + up(sizeof(ChunkResult) * n1),
                         o_done = o_marks + up(sizeof(unsigned long long) * ((size_t)j->n_marks + 2)),
                         total = o_done + up(sizeof(uint32_t) * n1);
            CK(ctx_arena_get(ctx, total, &j->h_arena, &j->h_arena_cap));
            j->arena_owner = ctx;
            memset(j->h_arena, 0, total);
            j->h_results = (ChunkResult *)(j->h_arena + o_res);
            j->h_pdesc = (ParseDesc *)(j->h_arena + o_pd);
            j->h_pout = (ParseOut *)(j->h_arena + o_po);
            j->h_totals = (ScanTotals *)(j->h_arena + o_tot);
            j->h_wresults = (ChunkResult *)(j->h_arena + o_wres);
            j->h_marks = (unsigned long long *)(j->h_arena + o_marks);
            j->h_done = (uint32_t *)(j->h_arena + o_done);
        }
        if (!j->zero_copy) {
            CK(pool_alloc(&j->d_comp, j->comp_alloc, ctx->stream));
            CK(cudaMemsetAsync(j->d_comp, 0, j->comp_alloc, ctx->stream));
            if (j->compact) {
                CK(pool_alloc(&j->d_cwin, j->cwin_bytes + 4096, ctx->stream));
                CK(cudaMemsetAsync(j->d_cwin, 0, j->cwin_bytes + 4096, ctx->stream));
            } else {
                CK(pool_alloc(&j->d_lead, j->lead_bytes, ctx->stream));
            }
        }
        if (j->compact) {
            CK(pool_alloc(&j->d_wdescs, sizeof(ChunkDesc) * (size_t)n, ctx->stream));
            CK(pool_alloc(&j->d_wresults, sizeof(ChunkResult) * (size_t)n, ctx->stream));
            CK(cudaMemcpyAsync(j->d_wdescs, wdescs.data(), sizeof(ChunkDesc) * (size_t)n, cudaMemcpyHostToDevice,
                               ctx->stream));
            CK(cudaMemsetAsync(j->d_wresults, 0, sizeof(ChunkResult) * (size_t)n, ctx->stream));
        }
        if (j->pipeline) {
            // the marks (bytes in place after every piece) are filled in by pp_job_upload; until the
            // first upload nothing waits
            CK(pool_alloc(&j->d_avail, sizeof(unsigned long long), ctx->stream));
            j->h_marks[0] = ~0ull;
            CK(cudaMemcpyAsync(j->d_avail, &j->h_marks[0], sizeof(unsigned long long), cudaMemcpyHostToDevice,
                               ctx->stream));
        }
        CK(cudaEventCreateWithFlags(&j->ev_reset, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&j->ev_lead, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&j->ev_exec_done, cudaEventDisableTiming));

        if (j->compact) {
            // nothing: the leads come out of the compact blob
        } else if (j->lead_direct) {
            pin_index_windows(ix);
        } else {
            CK(cudaHostAlloc(&j->h_lead, j->lead_bytes, cudaHostAllocDefault));
            for (int k = 0; k < n; k++) {
                const int c = first_chunk + k;
                const ChunkDesc &d = j->descs[(size_t)k];
                uint8_t *dst = j->h_lead + d.lead_src;
                const int32_t ol = ix->off_len[(size_t)c];
                if (ol <= PP_WINSIZE) {
                    memcpy(dst, ix->window(c), PP_WINSIZE);
                } else {
                    const uint32_t padn = d.lead_len - (uint32_t)ol;
                    memset(dst, 0, padn);
                    memcpy(dst + padn, ix->offset(c), (size_t)ol);
                }
            }
        }
        // records: start from >= 32 bytes per record; execute() grows the arrays if a corpus
        // has shorter records (the scan kernel reports the exact need)
        int rc = job_alloc_lines(j, scanned / 32 + n + 16);
        if (rc != PP_OK) return fail(rc);
        // exact-count overrides: -1 = none; 0 = forced through the exact parser
        std::vector<int64_t> ex((size_t)std::max(n, 1), -1);
        for (int k = 0; k < n; k++)
            if (forced_exact[(size_t)k]) ex[(size_t)k] = 0;
        CK(cudaMemcpyAsync(j->d_exact, ex.data(), sizeof(int64_t) * (size_t)std::max(n, 1), cudaMemcpyHostToDevice,
                           ctx->stream));
        CK(cudaMemcpyAsync(j->d_descs, j->descs.data(), sizeof(ChunkDesc) * (size_t)std::max(n, 1),
                           cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemsetAsync(j->d_pout, 0, sizeof(ParseOut) * (size_t)std::max(n, 1), ctx->stream));
        CK(cudaMemsetAsync(j->d_results, 0, sizeof(ChunkResult) * (size_t)std::max(n, 1), ctx->stream));
        for (auto &e : j->ev) CK(cudaEventCreate(&e));
        CK(cudaStreamSynchronize(ctx->stream));
    } catch (...) {
        return fail(PP_MEM_ERROR);
    }
    return PP_OK;
}

extern "C" {

int pp_job_create(pp_ctx *ctx, const pp_index *ix, size_t gz_len, int32_t first_chunk, int32_t n_chunks,
                  uint32_t flags, pp_job **out)
{
    if (!ctx || !ix || !out) return PP_E_ARG;
    *out = nullptr;
    const int32_t nchunks_total = ix->count() - 1;
    if (n_chunks < 0) n_chunks = nchunks_total - first_chunk;
    if (first_chunk < 0 || n_chunks < 0 || first_chunk + n_chunks > nchunks_total) return PP_E_ARG;
    pp_job *j = nullptr;
    int rc;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        CK(cudaSetDevice(ctx->device));
        j = new (std::nothrow) pp_job();
        if (!j) return PP_MEM_ERROR;
        j->ctx = nullptr;  // set on success; pp_job_free must not lock the mutex we hold
        j->st_alloc = ctx->stream;
        j->device = ctx->device;
        j->ix = ix;
        j->first = first_chunk;
        j->n = n_chunks;
        j->flags = flags;
        j->zero_copy = (flags & PP_JOB_ZEROCOPY) != 0;
        j->pipeline = (flags & PP_JOB_PIPELINE) != 0 && !j->zero_copy;
        j->compact = (flags & PP_JOB_COMPACT_WINDOWS) != 0;
        rc = job_create_inner(j, ctx, ix, gz_len, first_chunk, n_chunks);
    }
    if (rc != PP_OK) {
        pp_job_free(j);
        return rc;
    }
    j->info.first_chunk = first_chunk;
    j->info.n_chunks = n_chunks;
    j->ctx = ctx;
    *out = j;
    return PP_OK;
}

int pp_job_file_range(const pp_job *j, int64_t *file_offset, int64_t *length)
{
    if (!j) return PP_E_ARG;
    if (file_offset) *file_offset = (int64_t)j->comp_file_lo;
    if (length) *length = (int64_t)j->comp_copy;
    return PP_OK;
}

int pp_job_upload_range(pp_job *j, const uint8_t *range, int64_t range_file_offset, int64_t range_len)
{
    if (!j || range_file_offset < 0 || range_len < 0) return PP_E_ARG;
    if (j->comp_copy &&
        (!range || (uint64_t)range_file_offset > j->comp_file_lo ||
         (uint64_t)range_file_offset + (uint64_t)range_len < j->comp_file_lo + j->comp_copy))
        return PP_E_ARG;  // the buffer does not cover the bytes the job reads
    // the job addresses the file as gz[offset]: hand it the base that puts `range` at its file offset
    // (only [comp_file_lo, comp_file_lo + comp_copy) is ever touched)
    return pp_job_upload(j, range ? range - range_file_offset : nullptr);
}

int pp_job_upload(pp_job *j, const uint8_t *gz)
{
    if (!j || !j->ctx || (!gz && j->comp_copy)) return PP_E_ARG;
    std::lock_guard<std::mutex> lk(j->ctx->mu);
    CK(cudaSetDevice(j->ctx->device));
    cudaStream_t st = j->ctx->stream;
    CK(cudaEventRecord(j->ev[0], st));
    int64_t h2d = 0;
    if (j->zero_copy) {
        // kernels read the caller's pinned buffer through its device alias: nothing to copy
        void *dp = nullptr;
        CK(cudaHostGetDevicePointer(&dp, const_cast<uint8_t *>(gz + j->comp_file_lo), 0));
        j->zc_comp = (const uint8_t *)dp;
    } else if (j->n > 0 && !j->pipeline) {
        CK(cudaMemcpyAsync(j->d_comp, gz + j->comp_file_lo, j->comp_copy, cudaMemcpyHostToDevice, st));
        h2d += (int64_t)j->comp_copy;
        if (j->compact) {
            CK(cudaMemcpyAsync(j->d_cwin, j->ix->cwin.data() + j->cwin_lo, j->cwin_bytes, cudaMemcpyHostToDevice, st));
            h2d += (int64_t)j->cwin_bytes;
        } else {
            const uint8_t *src = j->lead_direct ? j->ix->window(j->first) : j->h_lead;
            CK(cudaMemcpyAsync(j->d_lead, src, j->lead_bytes, cudaMemcpyHostToDevice, st));
            h2d += (int64_t)j->lead_bytes;
        }
    } else if (j->n > 0) {
        // pipelined: the copies go to the copy stream and pp_job_execute's kernel overlaps them.
        // The buffers may still be read by the previous execute: order the copies behind it.
        cudaStream_t cs = j->ctx->copy_stream;
        if (j->exec_pending) CK(cudaStreamWaitEvent(cs, j->ev_exec_done, 0));
        int nm = 0;
        j->h_marks[nm] = 0;   // reset: nothing in place
        CK(cudaMemcpyAsync(j->d_avail, &j->h_marks[nm++], sizeof(unsigned long long), cudaMemcpyHostToDevice, cs));
        CK(cudaEventRecord(j->ev_reset, cs));
        if (j->compact) {
            CK(cudaMemcpyAsync(j->d_cwin, j->ix->cwin.data() + j->cwin_lo, j->cwin_bytes, cudaMemcpyHostToDevice, cs));
            h2d += (int64_t)j->cwin_bytes;
        } else {
            const uint8_t *src = j->lead_direct ? j->ix->window(j->first) : j->h_lead;
            CK(cudaMemcpyAsync(j->d_lead, src, j->lead_bytes, cudaMemcpyHostToDevice, cs));
            h2d += (int64_t)j->lead_bytes;
        }
        CK(cudaEventRecord(j->ev_lead, cs));
        const uint8_t *src0 = gz + j->comp_file_lo;
        const uint64_t kPiece = 8ull << 20;
        for (uint64_t off = 0; off < j->comp_copy; off += kPiece) {
            const uint64_t len = std::min<uint64_t>(kPiece, j->comp_copy - off);
            CK(cudaMemcpyAsync(j->d_comp + off, src0 + off, len, cudaMemcpyHostToDevice, cs));
            j->h_marks[nm] = off + len >= j->comp_copy ? ~0ull : off + len;
            CK(cudaMemcpyAsync(j->d_avail, &j->h_marks[nm++], sizeof(unsigned long long), cudaMemcpyHostToDevice, cs));
        }
        h2d += (int64_t)j->comp_copy;
        // the kernels must not start before the mark was reset and the leads are in place
        CK(cudaStreamWaitEvent(st, j->ev_reset, 0));
        CK(cudaStreamWaitEvent(st, j->ev_lead, 0));
    }
    CK(cudaEventRecord(j->ev[1], st));
    j->info.h2d_bytes = h2d;
    return PP_OK;
}

static int job_parse_stage(pp_job *j, cudaStream_t st, bool with_pout_flags, int *launches)
{
    const int n = j->n;
    CK(launch_exact_count(j->d_slots, j->d_descs, j->d_results, with_pout_flags ? j->d_pout : nullptr, n, j->d_exact, st));
    CK(launch_scan(j->d_descs, j->d_results, j->d_exact, n, (j->flags & PP_JOB_STRICT) ? 1u : 0u, j->rec_cap, j->d_pdesc,
                   j->d_pout, j->d_totals, st));
    CK(cudaEventRecord(j->ev[4], st));
    CK(launch_parse(j->d_slots, j->d_pdesc, n, j->d_tile_base, j->total_tiles, j->max_tiles, j->d_lines, j->rec_cap,
                    j->d_pout, j->d_totals, j->d_parse_work, j->ctx->sm_count, j->ctx->parse_per_sm, st));
    CK(cudaEventRecord(j->ev[5], st));
    CK(launch_exact_emit(j->d_slots, j->d_pdesc, n, j->d_lines, j->rec_cap, j->d_pout, j->d_totals, st));
    *launches += n > 0 ? 4 : 1;
    return PP_OK;
}

static int job_execute_locked(pp_job *j, bool stream_done)
{
    CK(cudaSetDevice(j->ctx->device));
    cudaStream_t st = j->ctx->stream;
    int launches = 0;
    CK(cudaEventRecord(j->ev[2], st));
    const uint8_t *comp = j->zero_copy ? j->zc_comp : j->d_comp;
    const uint8_t *lead = j->d_lead;
    const uint8_t *cwin = j->d_cwin;
    uint64_t comp_bytes = j->comp_alloc;
    if (j->zero_copy) {
        if (!comp) return PP_E_ARG;
        void *dp = nullptr;
        if (j->compact) {
            if (!j->ix->cwin_pinned) return PP_E_CUDA;  // the compact blob could not be pinned
            CK(cudaHostGetDevicePointer(&dp, const_cast<uint8_t *>(j->ix->cwin.data()), 0));
            cwin = (const uint8_t *)dp + j->cwin_lo;
        } else {
            const uint8_t *hl = j->lead_direct ? j->ix->window(j->first) : j->h_lead;
            CK(cudaHostGetDevicePointer(&dp, const_cast<uint8_t *>(hl), 0));
            lead = (const uint8_t *)dp;
        }
        // the TRUE extent of what the job may read: nothing behind the caller's buffer is touched, so
        // any pinned buffer works (pp_host_alloc or pp_host_register); the kernel zero-fills past it
        comp_bytes = j->comp_copy;
    }
    if (j->compact) {
        // pre-pass: every checkpoint window is a small independent zlib stream; the same kernel inflates
        // them into the slots' lead areas (chunk descriptors then say "history already in place")
        CK(launch_inflate(j->d_wdescs, j->n, cwin, j->cwin_bytes, j->d_slots, j->d_slots, j->d_wresults,
                          j->ctx->inflate_cfg(j->n), st));
        launches += j->n > 0 ? 1 : 0;
        lead = j->d_slots;  // unused
    }
    InflateSync sy;
    if (j->pipeline) {
        sy.gate = j->gate;
        sy.gate.mark = j->d_avail;
    }
    if (stream_done) {
        void *dp = nullptr;
        CK(cudaHostGetDevicePointer(&dp, j->h_done, 0));
        sy.done = (uint32_t *)dp;
    }
    sy.pull = j->zero_copy && !getenv("PPB200_NO_PULL_REUSE");  // the input crosses PCIe: staged bytes are re-used
    CK(launch_inflate(j->d_descs, j->n, comp, comp_bytes, j->d_slots, lead, j->d_results, j->ctx->inflate_cfg(j->n),
                      st, sy));
    launches += j->n > 0 ? 1 : 0;  // (the chunk-counter memset is not a kernel)
    CK(cudaEventRecord(j->ev[3], st));
    CK(cudaEventRecord(j->ev_exec_done, st));
    j->exec_pending = true;
    int rc = job_parse_stage(j, st, false, &launches);
    if (rc != PP_OK) return rc;
    j->info.launches = launches;
    j->have_results = false;
    return PP_OK;
}

int pp_job_execute(pp_job *j)
{
    if (!j || !j->ctx) return PP_E_ARG;
    std::lock_guard<std::mutex> lk(j->ctx->mu);
    return job_execute_locked(j, false);
}

int pp_job_execute_to_host(pp_job *j, uint8_t *dst, int64_t cap)
{
    if (!j || !j->ctx || !dst) return PP_E_ARG;
    int64_t need = 0;
    for (int k = 0; k < j->n; k++) need += j->descs[(size_t)k].out_len;
    if (cap < need) return PP_BUF_ERROR;
    std::lock_guard<std::mutex> lk(j->ctx->mu);
    CK(cudaSetDevice(j->ctx->device));
    for (int k = 0; k < j->n; k++) ((volatile uint32_t *)j->h_done)[k] = 0;
    int rc = job_execute_locked(j, true);
    if (rc != PP_OK) return rc;
    // chunks are handed out, and so finish, in file order: copy each one out as soon as its flag is up
    cudaStream_t ds = j->ctx->d2h_stream;
    int64_t pos = 0;
    const auto t0 = std::chrono::steady_clock::now();
    for (int k = 0; k < j->n; k++) {
        const ChunkDesc &d = j->descs[(size_t)k];
        unsigned spins = 0;
        while (((volatile uint32_t *)j->h_done)[k] == 0) {
            if ((++spins & 0xfffu) == 0) {
                // a kernel that died never raises the flag: notice, and never spin forever
                const cudaError_t q = cudaStreamQuery(j->ctx->stream);
                if (q != cudaSuccess && q != cudaErrorNotReady) { CK(q); }
                if (q == cudaSuccess && ((volatile uint32_t *)j->h_done)[k] == 0) return PP_E_CUDA;
                if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120)) return PP_E_CUDA;
                std::this_thread::yield();
            }
        }
        if (d.out_len)
            CK(cudaMemcpyAsync(dst + pos, j->d_slots + d.slot_off + d.lead_len, d.out_len, cudaMemcpyDeviceToHost, ds));
        pos += d.out_len;
    }
    CK(cudaStreamSynchronize(ds));
    return PP_OK;
}

int pp_job_download(pp_job *j)
{
    if (!j || !j->ctx) return PP_E_ARG;
    std::lock_guard<std::mutex> lk(j->ctx->mu);
    CK(cudaSetDevice(j->ctx->device));
    cudaStream_t st = j->ctx->stream;
    const size_t n = (size_t)std::max(j->n, 1);
    // download, and when the results ask for another parse pass (record arrays too small, or a chunk
    // the fast parser flagged for the exact parser) run it and download AGAIN: the last action is
    // always a download, and a job that does not settle is an error, never stale results
    bool settled = false;
    for (int attempt = 0; attempt < 5 && !settled; attempt++) {
        CK(cudaEventRecord(j->ev[6], st));
        CK(cudaMemcpyAsync(j->h_results, j->d_results, sizeof(ChunkResult) * n, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(j->h_pdesc, j->d_pdesc, sizeof(ParseDesc) * n, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(j->h_pout, j->d_pout, sizeof(ParseOut) * n, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(j->h_totals, j->d_totals, sizeof(ScanTotals), cudaMemcpyDeviceToHost, st));
        CK(cudaEventRecord(j->ev[7], st));
        CK(cudaStreamSynchronize(st));
        bool redo = false;
        if (j->h_totals->overflow) {
            // more records than the arrays hold: grow to the exact need and parse again.  (Device
            // pointers handed out earlier by pp_job_device_ptrs are invalidated by this; callers take
            // them after pp_job_download, see ppb200.h.)
            int rc = job_alloc_lines(j, j->h_totals->total_records + 16);
            if (rc != PP_OK) return rc;
            redo = true;
        } else {
            for (int k = 0; k < j->n; k++)
                if ((j->h_pout[k].flags & 1u) && !j->h_pdesc[k].exact) { redo = true; break; }
        }
        if (!redo) { settled = true; break; }
        int launches = 0;
        int rc = job_parse_stage(j, st, true, &launches);
        if (rc != PP_OK) return rc;
        j->info.launches += launches;
    }
    if (!settled) {
        fprintf(stderr, "ppb200: parse stage did not settle after 4 re-runs\n");
        return PP_E_CUDA;
    }
    pp_job_info &I = j->info;
    I.total_records = j->h_totals->total_records;
    I.total_bytes = j->h_totals->total_bytes;
    I.scanned_bytes = j->h_totals->scanned_bytes;
    I.status = j->h_totals->first_status;
    I.exact_chunks = j->h_totals->exact_chunks;
    I.compressed_bytes = (int64_t)j->comp_copy;
    I.d2h_bytes = (int64_t)((sizeof(ChunkResult) + sizeof(ParseDesc) + sizeof(ParseOut)) * n + sizeof(ScanTotals));
    auto ms = [&](float *dst, int a, int b) {
        if (cudaEventElapsedTime(dst, j->ev[a], j->ev[b]) != cudaSuccess) {
            *dst = 0.f;  // an event not recorded yet (e.g. no upload) is not an error of the job
            cudaGetLastError();
        }
    };
    ms(&I.upload_ms, 0, 1);
    ms(&I.inflate_ms, 2, 3);
    ms(&I.scan_ms, 3, 4);
    ms(&I.parse_ms, 4, 5);
    ms(&I.download_ms, 6, 7);
    j->have_results = true;
    return PP_OK;
}

int pp_job_info_get(const pp_job *j, pp_job_info *out)
{
    if (!j || !out) return PP_E_ARG;
    *out = j->info;
    return PP_OK;
}

int pp_job_chunk_info(const pp_job *j, int32_t k, pp_chunk_info *out)
{
    if (!j || !out || k < 0 || k >= j->n || !j->have_results) return PP_E_ARG;
    const ChunkResult &r = j->h_results[k];
    const ParseDesc &p = j->h_pdesc[k];
    const ParseOut &o = j->h_pout[k];
    out->status = r.status;
    out->prefix_len = (int32_t)j->descs[(size_t)k].prefix_len;
    out->inflated = r.produced;
    out->records = p.rec_count;
    out->record_base = p.rec_base;
    out->parse_end = o.parse_end;
    out->flags = p.exact ? 1u : 0u;
    return PP_OK;
}

int pp_job_fetch_line_starts(pp_job *j, uint32_t *l0, uint32_t *l1, uint32_t *l2, uint32_t *l3)
{
    if (!j || !j->ctx || !j->have_results) return PP_E_ARG;
    std::lock_guard<std::mutex> lk(j->ctx->mu);
    CK(cudaSetDevice(j->ctx->device));
    const size_t bytes = (size_t)j->info.total_records * sizeof(uint32_t);
    uint32_t *dst[4] = {l0, l1, l2, l3};
    for (int f = 0; f < 4; f++)
        if (dst[f] && bytes)
            CK(cudaMemcpyAsync(dst[f], j->d_lines + (size_t)f * (size_t)j->rec_cap, bytes, cudaMemcpyDeviceToHost,
                               j->ctx->stream));
    CK(cudaStreamSynchronize(j->ctx->stream));
    return PP_OK;
}

int pp_job_fetch_chunk(pp_job *j, int32_t k, uint8_t *dst, int64_t cap)
{
    if (!j || !j->ctx || !j->have_results || k < 0 || k >= j->n || !dst) return PP_E_ARG;
    const int64_t nbytes = j->h_results[k].produced;
    if (cap < nbytes) return PP_BUF_ERROR;
    std::lock_guard<std::mutex> lk(j->ctx->mu);
    CK(cudaSetDevice(j->ctx->device));
    const ChunkDesc &d = j->descs[(size_t)k];
    if (nbytes) CK(cudaMemcpyAsync(dst, j->d_slots + d.slot_off + d.lead_len, (size_t)nbytes, cudaMemcpyDeviceToHost, j->ctx->stream));
    CK(cudaStreamSynchronize(j->ctx->stream));
    return PP_OK;
}

int pp_job_fetch_bytes(pp_job *j, uint8_t *dst, int64_t cap)
{
    if (!j || !j->ctx || !j->have_results || !dst) return PP_E_ARG;
    if (cap < j->info.total_bytes) return PP_BUF_ERROR;
    std::lock_guard<std::mutex> lk(j->ctx->mu);
    CK(cudaSetDevice(j->ctx->device));
    int64_t pos = 0;
    for (int k = 0; k < j->n; k++) {
        const ChunkDesc &d = j->descs[(size_t)k];
        const int64_t nbytes = j->h_results[k].produced;
        if (nbytes) CK(cudaMemcpyAsync(dst + pos, j->d_slots + d.slot_off + d.lead_len, (size_t)nbytes, cudaMemcpyDeviceToHost, j->ctx->stream));
        pos += nbytes;
    }
    CK(cudaStreamSynchronize(j->ctx->stream));
    return PP_OK;
}

int pp_job_base_histogram(pp_job *j, uint64_t counts[256])
{
    if (!j || !j->ctx || !counts || !j->have_results) return PP_E_ARG;
    std::lock_guard<std::mutex> lk(j->ctx->mu);
    CK(cudaSetDevice(j->ctx->device));
    DevBuf d;
    CK(d.alloc(256 * sizeof(unsigned long long)));
    CK(launch_base_histogram(j->d_slots, j->d_pdesc, j->n, j->d_lines, j->rec_cap, d.as<unsigned long long>(),
                             j->ctx->sm_count, j->ctx->stream));
    CK(cudaMemcpyAsync(counts, d.p, 256 * sizeof(uint64_t), cudaMemcpyDeviceToHost, j->ctx->stream));
    CK(cudaStreamSynchronize(j->ctx->stream));
    return PP_OK;
}

int pp_job_count_pattern(pp_job *j, const uint8_t *pattern, int32_t pattern_len, uint64_t *count)
{
    if (!j || !j->ctx || !count || !j->have_results || pattern_len < 0 || (pattern_len && !pattern))
        return PP_E_ARG;
    if (pattern_len == 0) {  // string.Contains("") is true for every record
        *count = (uint64_t)j->h_totals->total_records;
        return PP_OK;
    }
    std::lock_guard<std::mutex> lk(j->ctx->mu);
    CK(cudaSetDevice(j->ctx->device));
    DevBuf d;
    CK(d.alloc((size_t)pattern_len + sizeof(unsigned long long)));
    uint8_t *d_pat = d.as<uint8_t>() + sizeof(unsigned long long);
    CK(cudaMemcpyAsync(d_pat, pattern, (size_t)pattern_len, cudaMemcpyHostToDevice, j->ctx->stream));
    CK(launch_pattern_count(j->d_slots, j->d_pdesc, j->n, j->d_lines, j->rec_cap, d_pat, pattern_len,
                            d.as<unsigned long long>(), j->ctx->sm_count, j->ctx->stream));
    CK(cudaMemcpyAsync(count, d.p, sizeof(uint64_t), cudaMemcpyDeviceToHost, j->ctx->stream));
    CK(cudaStreamSynchronize(j->ctx->stream));
    return PP_OK;
}

int pp_job_digests(pp_job *j, uint64_t *bytes_digest, uint64_t *fields_digest)
{
    if (!j || !j->ctx || !j->have_results) return PP_E_ARG;
    if (j->n <= 0) return PP_OK;
    std::lock_guard<std::mutex> lk(j->ctx->mu);
    CK(cudaSetDevice(j->ctx->device));
    DevBuf d;
    const size_t words = (size_t)j->n * 2;
    CK(d.alloc(words * sizeof(unsigned long long)));
    CK(launch_digests(j->d_slots, j->d_descs, j->d_results, j->d_pdesc, j->d_pout, j->n, j->d_lines, j->rec_cap,
                      d.as<unsigned long long>(), j->ctx->stream));
    std::vector<uint64_t> h(words);
    CK(cudaMemcpyAsync(h.data(), d.p, words * sizeof(uint64_t), cudaMemcpyDeviceToHost, j->ctx->stream));
    CK(cudaStreamSynchronize(j->ctx->stream));
    for (int k = 0; k < j->n; k++) {
        if (bytes_digest) bytes_digest[k] = h[2 * (size_t)k];
        if (fields_digest) fields_digest[k] = h[2 * (size_t)k + 1];
    }
    return PP_OK;
}

int pp_job_device_ptrs(const pp_job *j, const uint8_t **slots, const uint64_t **chunk_data_off, const uint32_t **l0,
                       const uint32_t **l1, const uint32_t **l2, const uint32_t **l3)
{
    if (!j) return PP_E_ARG;
    if (slots) *slots = j->d_slots;
    // ParseDesc is {u64 data_off, ...} with a 32-byte stride; the first field is the offset
    if (chunk_data_off) *chunk_data_off = reinterpret_cast<const uint64_t *>(j->d_pdesc);
    const uint32_t *b = j->d_lines;
    if (l0) *l0 = b;
    if (l1) *l1 = b + (size_t)j->rec_cap;
    if (l2) *l2 = b + 2 * (size_t)j->rec_cap;
    if (l3) *l3 = b + 3 * (size_t)j->rec_cap;
    return PP_OK;
}

int pp_decompress_all(pp_ctx *ctx, const pp_index *ix, const uint8_t *gz, size_t gz_len, int32_t first_chunk,
                      int32_t n_chunks, uint32_t flags, pp_job **out)
{
    if (!out) return PP_E_ARG;
    pp_job *j = nullptr;
    int rc = pp_job_create(ctx, ix, gz_len, first_chunk, n_chunks, flags, &j);
    if (rc != PP_OK) return rc;
    rc = pp_job_upload(j, gz);
    if (rc == PP_OK) rc = pp_job_execute(j);
    if (rc == PP_OK) rc = pp_job_download(j);
    if (rc != PP_OK) {
        pp_job_free(j);
        return rc;
    }
    *out = j;
    return j->info.status;
}


// ------------------------------------------------------------ cached contexts
// The one-call entry points over several GPUs / paired files need a context (stream, 600 MB of token
// scratch, occupancy queries) per GPU and per concurrent job; opening one costs tens of milliseconds,
// more than the decode of 10 M reads.  They borrow contexts from a process-wide cache instead and give
// them back when the handle is freed; pp_release_cached_contexts() closes the idle ones.
static std::mutex g_ctx_cache_mu;
static std::vector<pp_ctx *> g_ctx_cache;

static int ctx_acquire(int32_t device, pp_ctx **out)
{
    {
        std::lock_guard<std::mutex> lk(g_ctx_cache_mu);
        for (size_t i = 0; i < g_ctx_cache.size(); i++)
            if (g_ctx_cache[i]->device == device) {
                *out = g_ctx_cache[i];
                g_ctx_cache.erase(g_ctx_cache.begin() + (long)i);
                return PP_OK;
            }
    }
    return pp_open(device, out);
}

static void ctx_release(pp_ctx *c)
{
    if (!c) return;
    std::lock_guard<std::mutex> lk(g_ctx_cache_mu);
    g_ctx_cache.push_back(c);
}

void pp_release_cached_contexts(void)
{
    std::vector<pp_ctx *> idle;
    {
        std::lock_guard<std::mutex> lk(g_ctx_cache_mu);
        idle.swap(g_ctx_cache);
    }
    for (pp_ctx *c : idle) pp_close(c);
}

// ------------------------------------------------------------ multi-GPU DecompressAll
//
// Index chunks are independent (chunk k needs index[k], index[k+1] and the file bytes
// [Input_k - 1, Input_{k+1}), LazyFileReader.cs:53-69), so DecompressAll over several GPUs is a
// partition of the chunk list into contiguous ranges of near-equal COMPRESSED size, one range per
// GPU, each GPU touching only its byte range and its points' windows.  No collective: the only
// cross-GPU quantity is the global ordinal of a range's first record, an exclusive prefix sum over
// `n_parts` integers done here on the host.

int pp_partition_chunks(const pp_index *ix, int32_t parts, int32_t *first_chunk, int32_t *n_chunks)
{
    if (!ix || parts <= 0 || !first_chunk || !n_chunks) return PP_E_ARG;
    const int32_t n = std::max(ix->count() - 1, 0);
    if (n == 0) {
        for (int32_t r = 0; r < parts; r++) first_chunk[r] = n_chunks[r] = 0;
        return PP_OK;
    }
    const int64_t in0 = ix->input[0], total = ix->input[(size_t)n] - in0;
    int32_t prev = 0;
    for (int32_t r = 0; r < parts; r++) {
        int32_t cut = n;
        if (r + 1 < parts) {
            // first point whose Input reaches this rank's share of the compressed bytes; never backwards
            const int64_t target = in0 + (int64_t)((__int128)total * (r + 1) / parts);
            cut = (int32_t)(std::lower_bound(ix->input.begin(), ix->input.begin() + n + 1, target) - ix->input.begin());
            cut = std::min(std::max(cut, prev), n);
        }
        first_chunk[r] = prev;
        n_chunks[r] = cut - prev;
        prev = cut;
    }
    return PP_OK;
}

struct pp_multi {
    std::vector<pp_ctx *> ctxs;
    std::vector<pp_job *> jobs;
    std::vector<int64_t> record_base;
    std::vector<int> rc;
    pp_multi_info info{};
};

void pp_multi_free(pp_multi *m)
{
    if (!m) return;
    for (pp_job *j : m->jobs) pp_job_free(j);
    for (pp_ctx *c : m->ctxs) ctx_release(c);
    delete m;
}

int pp_decompress_all_multi(const int32_t *devices, int32_t n_devices, const pp_index *ix, const uint8_t *gz,
                            size_t gz_len, uint32_t flags, pp_multi **out)
{
    if (!devices || n_devices <= 0 || !ix || !out || (!gz && gz_len)) return PP_E_ARG;
    *out = nullptr;
    pp_multi *m = new (std::nothrow) pp_multi();
    if (!m) return PP_MEM_ERROR;
    int rc = PP_OK;
    try {
        m->ctxs.assign((size_t)n_devices, nullptr);
        m->jobs.assign((size_t)n_devices, nullptr);
        m->record_base.assign((size_t)n_devices, 0);
        m->rc.assign((size_t)n_devices, PP_OK);
        std::vector<int32_t> first((size_t)n_devices), cnt((size_t)n_devices);
        rc = pp_partition_chunks(ix, n_devices, first.data(), cnt.data());
        if (rc == PP_OK && (flags & PP_JOB_COMPACT_WINDOWS) && !index_build_compact_windows(ix)) rc = PP_MEM_ERROR;
        if (rc == PP_OK) {
            // pin the shared host buffers ONCE, before the per-GPU threads start (cudaHostRegister is
            // portable across the contexts of one process)
            pin_index_windows(ix);
            if (flags & PP_JOB_COMPACT_WINDOWS) pin_index_cwin(ix);
            // one host thread per GPU: context, plan, upload, kernels, download
            std::vector<std::thread> th;
            for (int32_t r = 0; r < n_devices; r++)
                th.emplace_back([&, r]() {
                    int e = ctx_acquire(devices[r], &m->ctxs[(size_t)r]);
                    if (e == PP_OK)
                        e = pp_decompress_all(m->ctxs[(size_t)r], ix, gz, gz_len, first[(size_t)r], cnt[(size_t)r], flags,
                                              &m->jobs[(size_t)r]);
                    m->rc[(size_t)r] = e;
                });
            for (auto &t : th) t.join();
            int64_t base = 0;
            m->info.n_parts = n_devices;
            for (int32_t r = 0; r < n_devices; r++) {
                const int e = m->rc[(size_t)r];
                if (e < 0 && !(m->jobs[(size_t)r])) { rc = e; break; }   // an API failure (no job to report from)
                if (e < 0 && m->info.status == 0) m->info.status = e;      // a chunk's ZResult: the job exists
                m->record_base[(size_t)r] = base;
                const pp_job_info &I = m->jobs[(size_t)r]->info;
                base += I.total_records;
                m->info.total_records += I.total_records;
                m->info.total_bytes += I.total_bytes;
                m->info.compressed_bytes += I.compressed_bytes;
                m->info.n_chunks += I.n_chunks;
            }
        }
    } catch (...) {
        rc = PP_MEM_ERROR;
    }
    if (rc != PP_OK) {
        pp_multi_free(m);
        return rc;
    }
    *out = m;
    return m->info.status;
}

int pp_multi_info_get(const pp_multi *m, pp_multi_info *out)
{
    if (!m || !out) return PP_E_ARG;
    *out = m->info;
    return PP_OK;
}

int pp_multi_part(const pp_multi *m, int32_t part, pp_job **job, int32_t *device, int64_t *record_base)
{
    if (!m || part < 0 || part >= (int32_t)m->jobs.size()) return PP_E_ARG;
    if (job) *job = m->jobs[(size_t)part];
    if (device) *device = m->ctxs[(size_t)part] ? m->ctxs[(size_t)part]->device : -1;
    if (record_base) *record_base = m->record_base[(size_t)part];
    return PP_OK;
}

// ------------------------------------------------------------ paired-end R1/R2
//
// The reference names paired files only in its assignment text (README.md:9 — "chunks with identical
// record counts"); its index cannot promise that: checkpoints sit on deflate block ends, which fall
// at different records in R1 and R2.  What CAN be kept identical is the ORDINAL of a record: with the
// H1 duplicates dropped (PP_JOB_STRICT) record r of R1 and record r of R2 are mates.  So a paired
// DecompressAll is: R1 and R2 decoded as two DecompressAll's whose chunk lists are partitioned over the
// same GPUs; per part the global ordinal ranges of its R1 and R2 records are compared and the few R2
// chunks that hold mates of the part's R1 records but landed on a neighbour are decoded here as well
// ("top-up" jobs).  Every R1 record's mate is then resident on the same GPU, and pp_pair_locate maps an
// ordinal to (job, record index) — the "identical chunk record counts" of the assignment, expressed as
// a record-range map.

struct pp_pair {
    int32_t n_parts = 0;
    std::vector<pp_ctx *> ctx1, ctx2;
    std::vector<pp_job *> r1;                  // per part
    std::vector<std::vector<pp_job *>> r2;     // per part: jobs ordered by ordinal (top-up left, main, top-up right)
    std::vector<int64_t> r1_base;              // global ordinal of the part's first R1 record
    std::vector<std::vector<int64_t>> r2_base; // same for every R2 job of the part
    pp_pair_info info{};
};

void pp_pair_free(pp_pair *p)
{
    if (!p) return;
    for (pp_job *j : p->r1) pp_job_free(j);
    for (auto &v : p->r2)
        for (pp_job *j : v) pp_job_free(j);
    for (pp_ctx *c : p->ctx1) ctx_release(c);
    for (pp_ctx *c : p->ctx2) ctx_release(c);
    delete p;
}

// chunk range [first, first+n) of `ix` (whose chunks' first-record ordinals are `chunk_base`, n+1 entries
// with the total at the end) that holds ordinals [lo, hi)
static void chunks_for_ordinals(const std::vector<int64_t> &chunk_base, int64_t lo, int64_t hi, int32_t *first, int32_t *n)
{
    const int32_t nch = (int32_t)chunk_base.size() - 1;
    int32_t a = (int32_t)(std::upper_bound(chunk_base.begin(), chunk_base.end(), lo) - chunk_base.begin()) - 1;
    a = std::min(std::max(a, 0), nch);
    int32_t b = (int32_t)(std::lower_bound(chunk_base.begin(), chunk_base.end(), hi) - chunk_base.begin());
    b = std::min(std::max(b, a), nch);
    *first = a;
    *n = b - a;
}

int pp_pair_decompress_all(const int32_t *devices, int32_t n_devices, const pp_index *ix1, const uint8_t *gz1,
                           size_t gz1_len, const pp_index *ix2, const uint8_t *gz2, size_t gz2_len, uint32_t flags,
                           pp_pair **out)
{
    if (!devices || n_devices <= 0 || !ix1 || !ix2 || !out || (!gz1 && gz1_len) || (!gz2 && gz2_len)) return PP_E_ARG;
    *out = nullptr;
    flags |= PP_JOB_STRICT;  // mates are paired by ordinal: the H1 duplicate must not shift one file against the other
    pp_pair *p = new (std::nothrow) pp_pair();
    if (!p) return PP_MEM_ERROR;
    int rc = PP_OK;
    try {
        const size_t P = (size_t)n_devices;
        p->n_parts = n_devices;
        p->ctx1.assign(P, nullptr);
        p->ctx2.assign(P, nullptr);
        p->r1.assign(P, nullptr);
        p->r2.assign(P, {});
        p->r1_base.assign(P, 0);
        p->r2_base.assign(P, {});
        std::vector<int32_t> f1(P), n1(P), f2(P), n2(P);
        rc = pp_partition_chunks(ix1, n_devices, f1.data(), n1.data());
        if (rc == PP_OK) rc = pp_partition_chunks(ix2, n_devices, f2.data(), n2.data());
        if (rc == PP_OK && (flags & PP_JOB_COMPACT_WINDOWS) &&
            (!index_build_compact_windows(ix1) || !index_build_compact_windows(ix2)))
            rc = PP_MEM_ERROR;
        if (rc != PP_OK) throw rc;
        pin_index_windows(ix1);
        pin_index_windows(ix2);
        if (flags & PP_JOB_COMPACT_WINDOWS) { pin_index_cwin(ix1); pin_index_cwin(ix2); }
        // R2 parts are widened by one chunk on either side: checkpoints fall at different records in the two files,
        // so the mates of a part's first and last R1 records usually sit in the neighbour's boundary chunk — decoding
        // that chunk here from the start (two chunks among hundreds) spares a second, serial round of top-up jobs
        const int32_t nch2 = ix2->count() - 1;
        std::vector<int32_t> x2(P), m2(P);
        for (size_t g = 0; g < P; g++) {
            x2[g] = n2[g] > 0 ? std::max(f2[g] - 1, 0) : f2[g];
            m2[g] = n2[g] > 0 ? std::min(f2[g] + n2[g] + 1, nch2) - x2[g] : 0;
            p->info.topup_chunks += m2[g] - n2[g];
        }
        // phase 1: every GPU decodes its part of R1 and of R2 concurrently (two contexts = two streams:
        // the second job's CTAs fill the SMs the first one's last, thin wave leaves idle)
        std::vector<pp_job *> main2(P, nullptr);
        std::vector<int> e1(P, PP_OK), e2(P, PP_OK);
        {
            std::vector<std::thread> th;
            for (size_t g = 0; g < P; g++) {
                th.emplace_back([&, g]() {
                    int e = ctx_acquire(devices[g], &p->ctx1[g]);
                    if (e == PP_OK) e = pp_decompress_all(p->ctx1[g], ix1, gz1, gz1_len, f1[g], n1[g], flags, &p->r1[g]);
                    e1[g] = e;
                });
                th.emplace_back([&, g]() {
                    int e = ctx_acquire(devices[g], &p->ctx2[g]);
                    if (e == PP_OK) e = pp_decompress_all(p->ctx2[g], ix2, gz2, gz2_len, x2[g], m2[g], flags, &main2[g]);
                    e2[g] = e;
                });
            }
            for (auto &t : th) t.join();
        }
        for (size_t g = 0; g < P; g++) {
            if (!p->r1[g] || !main2[g]) {   // an API failure: no job to report from
                for (pp_job *j : main2) pp_job_free(j);
                throw (e1[g] < 0 ? e1[g] : e2[g] < 0 ? e2[g] : PP_E_CUDA);
            }
            if (e1[g] < 0 && p->info.status == 0) p->info.status = e1[g];
            if (e2[g] < 0 && p->info.status == 0) p->info.status = e2[g];
        }
        // global ordinals: per part of R1, and per chunk of R2 (a chunk's record count does not depend on the job
        // that decoded it: under PP_JOB_STRICT the H1 duplicate is a property of the checkpoint)
        std::vector<int64_t> b1(P + 1, 0);
        for (size_t g = 0; g < P; g++) b1[g + 1] = b1[g] + p->r1[g]->info.total_records;
        std::vector<int64_t> cb2((size_t)nch2 + 1, 0);   // first-record ordinal of every R2 chunk, total at the end
        for (size_t g = 0; g < P; g++) {
            const pp_job *j = main2[g];
            for (int32_t c = f2[g]; c < f2[g] + n2[g]; c++) {   // the part's own chunks
                const int k = c - x2[g];
                const int64_t hi = k + 1 < j->n ? j->h_pdesc[k + 1].rec_base : j->info.total_records;
                cb2[(size_t)c + 1] = hi - j->h_pdesc[k].rec_base;
            }
        }
        for (int32_t c = 0; c < nch2; c++) cb2[(size_t)c + 1] += cb2[(size_t)c];
        std::vector<int64_t> lo2(P), hi2(P);               // ordinals the widened R2 job of part g holds
        for (size_t g = 0; g < P; g++) {
            lo2[g] = cb2[(size_t)x2[g]];
            hi2[g] = cb2[(size_t)(x2[g] + m2[g])];
        }
        p->info.records_r1 = b1[P];
        p->info.records_r2 = cb2[(size_t)nch2];
        p->info.pairs = std::min(b1[P], cb2[(size_t)nch2]);
        // phase 2: mates of this part's R1 records that a neighbour's R2 part holds
        std::vector<std::thread> th;
        std::vector<int> e3(2 * P, PP_OK);
        std::vector<pp_job *> left(P, nullptr), right(P, nullptr);
        std::vector<int64_t> lbase(P, 0), rbase(P, 0);
        for (size_t g = 0; g < P; g++) {
            const int64_t a1 = b1[g], z1 = std::min(b1[g + 1], cb2[(size_t)nch2]), a2 = lo2[g], z2 = hi2[g];
            struct Gap { int64_t lo, hi; pp_job **dst; int64_t *base; int slot; };
            const Gap gaps[2] = {{a1, std::min(z1, a2), &left[g], &lbase[g], (int)(2 * g)},
                                 {std::max(a1, z2), z1, &right[g], &rbase[g], (int)(2 * g + 1)}};
            for (const Gap &gp : gaps) {
                if (gp.lo >= gp.hi) continue;
                int32_t cf = 0, cn = 0;
                chunks_for_ordinals(cb2, gp.lo, gp.hi, &cf, &cn);
                if (cn <= 0) continue;
                *gp.base = cb2[(size_t)cf];
                p->info.topup_chunks += cn;
                pp_job **dst = gp.dst;
                const int slot = gp.slot;
                th.emplace_back([&, g, cf, cn, dst, slot]() {
                    e3[(size_t)slot] = pp_decompress_all(p->ctx2[g], ix2, gz2, gz2_len, cf, cn, flags, dst);
                });
            }
        }
        for (auto &t : th) t.join();
        for (size_t g = 0; g < P; g++) {
            p->r1_base[g] = b1[g];
            if (left[g]) { p->r2[g].push_back(left[g]); p->r2_base[g].push_back(lbase[g]); }
            p->r2[g].push_back(main2[g]);
            p->r2_base[g].push_back(lo2[g]);
            if (right[g]) { p->r2[g].push_back(right[g]); p->r2_base[g].push_back(rbase[g]); }
        }
        for (int e : e3)
            if (e < 0 && e > -100 && p->info.status == 0) p->info.status = e;
            else if (e <= -100) throw e;
        p->info.n_parts = n_devices;
    } catch (int e) {
        rc = e;
    } catch (...) {
        rc = PP_MEM_ERROR;
    }
    if (rc != PP_OK) {
        pp_pair_free(p);
        return rc;
    }
    *out = p;
    return p->info.status;
}

int pp_pair_info_get(const pp_pair *p, pp_pair_info *out)
{
    if (!p || !out) return PP_E_ARG;
    *out = p->info;
    return PP_OK;
}

int pp_pair_part(const pp_pair *p, int32_t part, pp_job **r1, int64_t *r1_base, int32_t *n_r2_jobs)
{
    if (!p || part < 0 || part >= p->n_parts) return PP_E_ARG;
    if (r1) *r1 = p->r1[(size_t)part];
    if (r1_base) *r1_base = p->r1_base[(size_t)part];
    if (n_r2_jobs) *n_r2_jobs = (int32_t)p->r2[(size_t)part].size();
    return PP_OK;
}

int pp_pair_part_r2(const pp_pair *p, int32_t part, int32_t which, pp_job **r2, int64_t *r2_base)
{
    if (!p || part < 0 || part >= p->n_parts || which < 0 || which >= (int32_t)p->r2[(size_t)part].size()) return PP_E_ARG;
    if (r2) *r2 = p->r2[(size_t)part][(size_t)which];
    if (r2_base) *r2_base = p->r2_base[(size_t)part][(size_t)which];
    return PP_OK;
}

int pp_pair_locate(const pp_pair *p, int32_t part, int64_t ordinal, int32_t *which_r2, int64_t *record_index)
{
    if (!p || part < 0 || part >= p->n_parts) return PP_E_ARG;
    const auto &jobs = p->r2[(size_t)part];
    const auto &base = p->r2_base[(size_t)part];
    for (size_t i = 0; i < jobs.size(); i++) {
        const int64_t n = jobs[i]->info.total_records;
        if (ordinal >= base[i] && ordinal < base[i] + n) {
            if (which_r2) *which_r2 = (int32_t)i;
            if (record_index) *record_index = ordinal - base[i];
            return PP_OK;
        }
    }
    return PP_E_ARG;  // the mate is not on this part (or R2 holds fewer records)
}

// ------------------------------------------------------------ single-call entry points


int64_t pp_extract(pp_ctx *ctx, const uint8_t *fileBuffer, int64_t fileBufferLen, const pp_index *ix,
                   int32_t from_point, uint8_t *buf, int64_t buf_len)
{
    if (!ctx || !fileBuffer || !ix || !buf || fileBufferLen < 0) return PP_E_ARG;
    if (from_point < 0 || from_point + 1 >= ix->count()) return PP_E_ARG;
    const int64_t len64 = ix->output[(size_t)from_point + 1] - ix->output[(size_t)from_point];
    if ((int32_t)len64 < 0) return 0;  // Core.cs:145
    if (len64 > buf_len || len64 > 0x7fffffffLL) return PP_BUF_ERROR;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int32_t bits = ix->bits[(size_t)from_point];
    // fileBuffer[0] is file byte from.Input-1 (LazyFileReader.cs:68): the stream starts
    // `bits` bits before fileBuffer[1] (Core.cs:151-157)
    ChunkDesc d{};
    d.in_bit = 8u - (uint64_t)bits;
    d.in_limit = (uint64_t)fileBufferLen;
    d.slot_off = 0;
    d.lead_src = 0;
    d.lead_len = PP_WINSIZE;
    d.out_len = (uint32_t)len64;
    const uint64_t comp_alloc = align_up((uint64_t)fileBufferLen, kTile) + 2 * kTile;
    const uint64_t slot_bytes = align_up((uint64_t)PP_WINSIZE + d.out_len + 1, 128);
    DevBuf comp, lead, slot, desc, res;
    CK(comp.alloc(comp_alloc));
    CK(lead.alloc(PP_WINSIZE));
    CK(slot.alloc(slot_bytes));
    CK(desc.alloc(sizeof d));
    CK(res.alloc(sizeof(ChunkResult)));
    CK(cudaMemsetAsync(comp.p, 0, comp_alloc, st));
    CK(cudaMemcpyAsync(comp.p, fileBuffer, (size_t)fileBufferLen, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(lead.p, ix->window(from_point), PP_WINSIZE, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(desc.p, &d, sizeof d, cudaMemcpyHostToDevice, st));
    if (launch_inflate(desc.as<ChunkDesc>(), 1, comp.as<uint8_t>(), comp_alloc, slot.as<uint8_t>(), lead.as<uint8_t>(),
                       res.as<ChunkResult>(), ctx->inflate_cfg(1), st) != cudaSuccess)
        return PP_E_CUDA;
    ChunkResult r{};
    CK(cudaMemcpyAsync(&r, res.p, sizeof r, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (r.status < 0) return r.status;  // the reference throws ZException(code), Core.cs:178-179
    if (r.produced) CK(cudaMemcpyAsync(buf, slot.as<uint8_t>() + PP_WINSIZE, r.produced, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return (int64_t)r.produced;
}

int64_t pp_parse(pp_ctx *ctx, const uint8_t *prepend, int64_t prepend_len, const uint8_t *rest, int64_t rest_len,
                 uint32_t *line_starts, int64_t cap, uint32_t *parse_end)
{
    if (!ctx || prepend_len < 0 || rest_len < 0 || (prepend_len && !prepend) || (rest_len && !rest)) return PP_E_ARG;
    if (prepend_len + rest_len > 0x7fffffffLL) return PP_E_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ChunkDesc d{};
    d.lead_len = (uint32_t)align_up((uint64_t)prepend_len, 16);
    d.out_len = (uint32_t)rest_len;
    d.prefix_len = (uint32_t)prepend_len;
    bool nul = false;
    count_prefix(prepend, (int32_t)prepend_len, &d.prefix_nl, &nul);
    const uint64_t slot_bytes = align_up((uint64_t)d.lead_len + d.out_len + 1, 128);
    const int64_t rec_cap = (prepend_len + rest_len) / 4 + 16;
    DevBuf slot, desc, res, pdesc, pout, totals, exact, lines, tbase, work;
    const uint64_t head0 = ((uint64_t)d.lead_len - d.prefix_len) & 15u;
    const uint32_t tiles0 = (uint32_t)((head0 + d.prefix_len + d.out_len + parse_tile_bytes() - 1) / parse_tile_bytes());
    const uint32_t tb_host[2] = {0u, tiles0};
    CK(tbase.alloc(sizeof tb_host));
    CK(work.alloc(sizeof(unsigned long long) * ((size_t)tiles0 + 1)));
    CK(cudaMemcpyAsync(tbase.p, tb_host, sizeof tb_host, cudaMemcpyHostToDevice, st));
    CK(slot.alloc(slot_bytes));
    CK(desc.alloc(sizeof d));
    CK(res.alloc(sizeof(ChunkResult)));
    CK(pdesc.alloc(sizeof(ParseDesc)));
    CK(pout.alloc(sizeof(ParseOut)));
    CK(totals.alloc(sizeof(ScanTotals)));
    CK(exact.alloc(sizeof(int64_t)));
    CK(lines.alloc((size_t)rec_cap * 4 * sizeof(uint32_t)));
    CK(cudaMemsetAsync(slot.p, 0, slot_bytes, st));
    CK(cudaMemsetAsync(pout.p, 0, sizeof(ParseOut), st));
    if (prepend_len)
        CK(cudaMemcpyAsync(slot.as<uint8_t>() + d.lead_len - prepend_len, prepend, (size_t)prepend_len,
                           cudaMemcpyHostToDevice, st));
    if (rest_len) CK(cudaMemcpyAsync(slot.as<uint8_t>() + d.lead_len, rest, (size_t)rest_len, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(desc.p, &d, sizeof d, cudaMemcpyHostToDevice, st));
    const int64_t ex0 = nul ? 0 : -1;
    CK(cudaMemcpyAsync(exact.p, &ex0, sizeof ex0, cudaMemcpyHostToDevice, st));
    if (launch_bytes_stats(slot.as<uint8_t>(), desc.as<ChunkDesc>(), res.as<ChunkResult>(), 1, st) != cudaSuccess)
        return PP_E_CUDA;
    ParseOut o{};
    ParseDesc p{};
    for (int pass = 0; pass < 2; pass++) {
        if (launch_exact_count(slot.as<uint8_t>(), desc.as<ChunkDesc>(), res.as<ChunkResult>(),
                               pass ? pout.as<ParseOut>() : nullptr, 1, exact.as<int64_t>(), st) != cudaSuccess ||
            launch_scan(desc.as<ChunkDesc>(), res.as<ChunkResult>(), exact.as<int64_t>(), 1, 0, rec_cap,
                        pdesc.as<ParseDesc>(), pout.as<ParseOut>(), totals.as<ScanTotals>(), st) != cudaSuccess ||
            launch_parse(slot.as<uint8_t>(), pdesc.as<ParseDesc>(), 1, tbase.as<uint32_t>(), tiles0, tiles0,
                         lines.as<uint32_t>(), rec_cap, pout.as<ParseOut>(), totals.as<ScanTotals>(),
                         work.as<unsigned long long>(), ctx->sm_count, ctx->parse_per_sm, st) != cudaSuccess ||
            launch_exact_emit(slot.as<uint8_t>(), pdesc.as<ParseDesc>(), 1, lines.as<uint32_t>(), rec_cap,
                              pout.as<ParseOut>(), totals.as<ScanTotals>(), st) != cudaSuccess)
            return PP_E_CUDA;
        CK(cudaMemcpyAsync(&o, pout.p, sizeof o, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&p, pdesc.p, sizeof p, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (!((o.flags & 1u) && !p.exact)) break;
    }
    const int64_t nrec = p.rec_count;
    if (parse_end) *parse_end = o.parse_end;
    if (line_starts && cap > 0 && nrec > 0) {
        // interleave the four arrays into line_starts[4r + f] on the host side of the copy
        const int64_t m = std::min(nrec, cap);
        std::vector<uint32_t> tmp((size_t)m * 4);
        for (int f = 0; f < 4; f++)
            CK(cudaMemcpyAsync(tmp.data() + (size_t)f * (size_t)m, lines.as<uint32_t>() + (size_t)f * (size_t)rec_cap,
                               (size_t)m * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        for (int64_t r = 0; r < m; r++)
            for (int f = 0; f < 4; f++) line_starts[4 * r + f] = tmp[(size_t)f * (size_t)m + (size_t)r];
    }
    return nrec;
}

}  // extern "C"

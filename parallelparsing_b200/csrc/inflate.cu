// Kernel 1 — checkpoint inflate.  One CTA per index chunk at a time (CTAs pull chunk
// numbers from a global counter), all threads of the CTA decoding sub-sequences of the
// same deflate block; the decoder itself lives in inflate_core.cuh.  Replaces
// Core.ExtractDeflateIndex + zlib inflate (Decompressor/Core.cs:133-192,
// Interop/PlatformInterop.cs:9-34).
#include "kernels.cuh"

namespace pp {

template <bool DUAL, bool PULL>
__device__ __forceinline__ void inflate_kernel_body(const ChunkDesc *__restrict__ descs, int n, const uint8_t *__restrict__ comp,
                                                    uint64_t comp_bytes, uint8_t *slots, const uint8_t *__restrict__ lead,
                                                    ChunkResult *__restrict__ results, uint32_t *scratch, size_t scratch_words,
                                                    int *next_chunk, uint32_t comp_shift, const InflateSync &sy,
                                                    ppinf::DualOut dual)
{
    extern __shared__ __align__(128) uint8_t pp_smem_raw[];
    ppinf::Sm sm;
    ppinf::sm_carve(sm, pp_smem_raw, (int)blockDim.x);
    if (threadIdx.x == 0) {
        ppinf::mbar_init(sm.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t *map = scratch + (size_t)blockIdx.x * scratch_words;
    uint32_t stage_phase = 0;
    for (;;) {
        if (threadIdx.x == 0) sm.u[16] = (uint32_t)atomicAdd(next_chunk, 1);
        __syncthreads();
        const int k = (int)sm.u[16];
        __syncthreads();
        if (k >= n) break;
        ppinf::ChunkDesc d = descs[k];
        // `comp` was aligned down to 16 bytes for the bulk copies: the chunk's bits sit comp_shift bytes further in
        d.in_bit += 8ull * comp_shift;
        d.in_limit += comp_shift;
        ppinf::inflate_chunk<DUAL, PULL>(sm, d, comp, comp_bytes, slots, lead, map, results[k], stage_phase, &sy.gate, dual);
        if (sy.done && threadIdx.x == 0) {
            // streamed download: tell the host (mapped pinned memory) that this chunk's bytes are final
            __threadfence_system();
            *((volatile uint32_t *)sy.done + k) = 1u;
        }
    }
}

__global__ void __launch_bounds__(ppinf::kMaxThreads, 1)
    pp_inflate_kernel(const ChunkDesc *__restrict__ descs, int n, const uint8_t *__restrict__ comp, uint64_t comp_bytes,
                      uint8_t *slots, const uint8_t *__restrict__ lead, ChunkResult *__restrict__ results,
                      uint32_t *scratch, size_t scratch_words, int *next_chunk, uint32_t comp_shift, InflateSync sy)
{
    inflate_kernel_body<false, false>(descs, n, comp, comp_bytes, slots, lead, results, scratch, scratch_words, next_chunk,
                                      comp_shift, sy, ppinf::DualOut{0, 0});
}

// Pull mode: `comp` is pinned host memory and the PCIe link is the limit; a window re-uses the bytes the window
// before it already brought over (stage_window_reuse).
__global__ void __launch_bounds__(ppinf::kMaxThreads, 1)
    pp_inflate_pull_kernel(const ChunkDesc *__restrict__ descs, int n, const uint8_t *__restrict__ comp, uint64_t comp_bytes,
                           uint8_t *slots, const uint8_t *__restrict__ lead, ChunkResult *__restrict__ results,
                           uint32_t *scratch, size_t scratch_words, int *next_chunk, uint32_t comp_shift, InflateSync sy)
{
    inflate_kernel_body<false, true>(descs, n, comp, comp_bytes, slots, lead, results, scratch, scratch_words, next_chunk,
                                     comp_shift, sy, ppinf::DualOut{0, 0});
}

// GPU CreateIndex: every chunk decoded once and resolved twice, against two histories (createindex.cu).
__global__ void __launch_bounds__(ppinf::kMaxThreads, 1)
    pp_inflate_dual_kernel(const ChunkDesc *__restrict__ descs, int n, const uint8_t *__restrict__ comp, uint64_t comp_bytes,
                           uint8_t *slots, const uint8_t *__restrict__ lead, ChunkResult *__restrict__ results,
                           uint32_t *scratch, size_t scratch_words, int *next_chunk, uint32_t comp_shift, InflateSync sy,
                           ppinf::DualOut dual)
{
    inflate_kernel_body<true, false>(descs, n, comp, comp_bytes, slots, lead, results, scratch, scratch_words, next_chunk,
                                     comp_shift, sy, dual);
}

// Debug aid: cycles thread 0 of every CTA spent per phase since the last call (and reset).
extern "C" int pp_internal_phase_cycles(unsigned long long *out, int n)
{
    unsigned long long h[ppinf::PH_COUNT];
    if (cudaMemcpyFromSymbol(h, ppinf::g_phase_cycles, sizeof h) != cudaSuccess) return -1;
    for (int i = 0; i < n && i < ppinf::PH_COUNT; i++) out[i] = h[i];
    unsigned long long z[ppinf::PH_COUNT] = {};
    cudaMemcpyToSymbol(ppinf::g_phase_cycles, z, sizeof z);
    return ppinf::PH_COUNT;
}

// The dynamic shared memory opt-in is a property of the function in the device's context, shared by
// every pp_ctx on that device: it is set ONCE per context, to the largest configuration, and never
// changed per launch (two contexts launching different CTA sizes would otherwise race on it).
cudaError_t inflate_set_max_smem(int threads)
{
    cudaError_t e = cudaFuncSetAttribute(pp_inflate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)ppinf::sm_bytes_for(threads));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(pp_inflate_dual_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)ppinf::sm_bytes_for(threads));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(pp_inflate_pull_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)ppinf::sm_bytes_for(threads));
}

int inflate_max_ctas_per_sm(int threads)
{
    int nb = 0;
    const size_t smem = ppinf::sm_bytes_for(threads);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, pp_inflate_kernel, threads, smem) != cudaSuccess) return 0;
    return nb;
}

size_t inflate_scratch_bytes(int threads, int grid)
{
    return (size_t)grid * ppinf::scratch_words_for(threads) * sizeof(uint32_t);
}

cudaError_t launch_inflate(const ChunkDesc *descs, int n, const uint8_t *comp, uint64_t comp_bytes, uint8_t *slots,
                           const uint8_t *lead, ChunkResult *results, const InflateLaunch &cfg, cudaStream_t st,
                           InflateSync sy)
{
    if (n <= 0) return cudaSuccess;
    // TMA bulk copies need 16-byte aligned global addresses: align the base down and shift the bit cursors
    const uint32_t comp_shift = (uint32_t)((uintptr_t)comp & 15u);
    comp -= comp_shift;
    comp_bytes += comp_shift;
    sy.gate.shift = comp_shift;
    cudaError_t e = cudaMemsetAsync(cfg.counter, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
    const size_t smem = ppinf::sm_bytes_for(cfg.threads);
    const int grid = n < cfg.grid ? n : cfg.grid;
    if (sy.pull)
        pp_inflate_pull_kernel<<<grid, cfg.threads, smem, st>>>(descs, n, comp, comp_bytes, slots, lead, results, cfg.map,
                                                                 ppinf::scratch_words_for(cfg.threads), cfg.counter, comp_shift, sy);
    else
        pp_inflate_kernel<<<grid, cfg.threads, smem, st>>>(descs, n, comp, comp_bytes, slots, lead, results, cfg.map,
                                                            ppinf::scratch_words_for(cfg.threads), cfg.counter, comp_shift, sy);
    return cudaGetLastError();
}

// One decode, two outputs: chunk k also lands slot_delta bytes further on in `slots`, resolved against the
// history lead_delta bytes further on in `lead` (same geometry and scratch as launch_inflate).
cudaError_t launch_inflate_dual(const ChunkDesc *descs, int n, const uint8_t *comp, uint64_t comp_bytes, uint8_t *slots,
                                const uint8_t *lead, ChunkResult *results, const InflateLaunch &cfg, uint64_t slot_delta,
                                uint64_t lead_delta, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    if ((slot_delta & 127u) || (lead_delta & 15u)) return cudaErrorInvalidValue;
    const uint32_t comp_shift = (uint32_t)((uintptr_t)comp & 15u);
    comp -= comp_shift;
    comp_bytes += comp_shift;
    InflateSync sy;
    sy.gate.shift = comp_shift;
    cudaError_t e = cudaMemsetAsync(cfg.counter, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
    const size_t smem = ppinf::sm_bytes_for(cfg.threads);
    const int grid = n < cfg.grid ? n : cfg.grid;
    pp_inflate_dual_kernel<<<grid, cfg.threads, smem, st>>>(descs, n, comp, comp_bytes, slots, lead, results, cfg.map,
                                                             ppinf::scratch_words_for(cfg.threads), cfg.counter, comp_shift, sy,
                                                             ppinf::DualOut{slot_delta, lead_delta});
    return cudaGetLastError();
}

}  // namespace pp

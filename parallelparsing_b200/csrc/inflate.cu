// Kernel 1 — checkpoint inflate.  One CTA (one warp) per index chunk; the decoder
// itself lives in inflate_core.cuh.  Replaces Core.ExtractDeflateIndex + zlib
// inflate (Decompressor/Core.cs:133-192, Interop/PlatformInterop.cs:9-34).
#include "kernels.cuh"

namespace pp {

__global__ void __launch_bounds__(32) pp_inflate_kernel(const ChunkDesc *__restrict__ descs, int n,
                                                        const uint8_t *__restrict__ comp, uint64_t comp_bytes,
                                                        uint8_t *slots, const uint8_t *__restrict__ lead,
                                                        ChunkResult *__restrict__ results)
{
    ppinf::Smem &sm = ppinf::g_sm;
    const int k = (int)blockIdx.x;
    if (k >= n) return;
    if (threadIdx.x == 0) {
        for (int s = 0; s < ppinf::kStages; s++) ppinf::mbar_init(&sm.bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    ppinf::inflate_chunk(descs[k], comp, comp_bytes, slots, lead, results[k]);
}

cudaError_t launch_inflate(const ChunkDesc *descs, int n, const uint8_t *comp, uint64_t comp_bytes, uint8_t *slots,
                           const uint8_t *lead, ChunkResult *results, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    pp_inflate_kernel<<<n, 32, 0, st>>>(descs, n, comp, comp_bytes, slots, lead, results);
    return cudaGetLastError();
}

}  // namespace pp

// How fast can SMs pull pinned host memory over PCIe with TMA bulk copies?  (The pull mode's ceiling.)
// Every CTA streams its own contiguous slice of a pinned host buffer into shared memory with
// cp.async.bulk, `depth` copies of `piece` bytes in flight per CTA.  Prints GB/s per configuration.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pull_probe pull_probe.cu && ./pull_probe
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int DEPTH>
__global__ void __launch_bounds__(128) pull_kernel(const uint8_t *src, size_t per_cta, uint32_t piece, unsigned long long *sink)
{
    extern __shared__ __align__(128) uint8_t buf[];
    __shared__ __align__(8) unsigned long long bar[DEPTH];
    const uint8_t *mine = src + (size_t)blockIdx.x * per_cta;
    const uint32_t n = (uint32_t)(per_cta / piece);
    if (threadIdx.x == 0) {
        for (int d = 0; d < DEPTH; d++) asm volatile("mbarrier.init.shared.b64 [%0], 1;" ::"r"(smem_u32(&bar[d])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned long long acc = 0;
    if (threadIdx.x == 0) {
        uint32_t phase[DEPTH];
        for (int d = 0; d < DEPTH; d++) phase[d] = 0;
        for (uint32_t i = 0; i < n + DEPTH; i++) {
            const int d = (int)(i % DEPTH);
            if (i >= DEPTH) {  // wait for copy i - DEPTH
                uint32_t ok = 0;
                while (!ok)
                    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                                 : "=r"(ok) : "r"(smem_u32(&bar[d])), "r"(phase[d]) : "memory");
                phase[d] ^= 1u;
                acc += buf[(size_t)d * piece];
            }
            if (i < n) {
                asm volatile("mbarrier.arrive.expect_tx.shared.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[d])), "r"(piece) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(buf + (size_t)d * piece)), "l"(mine + (size_t)i * piece), "r"(piece), "r"(smem_u32(&bar[d])) : "memory");
            }
        }
        atomicAdd(sink, acc);
    }
}

template <int DEPTH> static float run(const uint8_t *d_src, size_t total, int grid, uint32_t piece, unsigned long long *sink)
{
    const size_t per = total / (size_t)grid / piece * piece;
    cudaFuncSetAttribute(pull_kernel<DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(DEPTH * piece));
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 1e9f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(a);
        pull_kernel<DEPTH><<<grid, 128, DEPTH * piece>>>(d_src, per, piece, sink);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return (float)((double)per * grid / best / 1e6);
}

int main()
{
    const size_t total = (size_t)1 << 30;
    uint8_t *h = nullptr;
    if (cudaHostAlloc(&h, total, cudaHostAllocMapped) != cudaSuccess) return 1;
    for (size_t i = 0; i < total; i += 4096) h[i] = (uint8_t)i;
    unsigned long long *sink;
    cudaMalloc(&sink, 8);
    cudaMemset(sink, 0, 8);
    uint8_t *dcopy;
    cudaMalloc(&dcopy, total);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaMemcpy(dcopy, h, total, cudaMemcpyHostToDevice);
    cudaEventRecord(a);
    cudaMemcpyAsync(dcopy, h, total, cudaMemcpyHostToDevice);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    printf("copy engine H2D: %.1f GB/s\n", total / ms / 1e6);
    for (int grid : {148, 296, 592})
        for (uint32_t piece : {4096u, 16384u, 32768u}) {
            printf("grid %4d piece %6u: depth1 %.1f  depth2 %.1f  depth4 %.1f GB/s\n", grid, piece, run<1>(h, total, grid, piece, sink),
                   run<2>(h, total, grid, piece, sink), run<4>(h, total, grid, piece, sink));
        }
    printf("from HBM, grid 296 piece 32768 depth2: %.1f GB/s\n", run<2>(dcopy, total, 296, 32768u, sink));
    return 0;
}

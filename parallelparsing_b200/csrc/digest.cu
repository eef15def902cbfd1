// Per-chunk integrity digests computed where the data lives (pp_job_digests, include/ppb200.h).
//
// Parity at the BASELINE sizes (10 M reads = 3.9 GB inflated, 977 chunks) cannot bring every byte and
// every record back over PCIe for a comparison, so the GPU folds each chunk's inflated bytes
// (Core.ExtractDeflateIndex's output, Decompressor/Core.cs:133-192) and each chunk's records (the nine
// integers of Parsing.Parse, Decompressor/Parsing.cs:20-39) into two 64-bit sums of position-keyed
// terms.  The sums are order sensitive (every term is multiplied by a hash of its position), need no
// sequential chain (so all threads work) and are restated in the test oracle over ITS bytes/records.
#include "kernels.cuh"

namespace pp {

__device__ __forceinline__ unsigned long long mix64(unsigned long long x)
{
    unsigned long long z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ unsigned long long block_sum_u64(unsigned long long v, unsigned long long *s_warp)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    const int lane = (int)threadIdx.x & 31, warp = (int)threadIdx.x >> 5;
    if (lane == 0) s_warp[warp] = v;
    __syncthreads();
    unsigned long long r = 0;
    if (threadIdx.x == 0)
        for (int w = 0; w < ((int)blockDim.x + 31) / 32; w++) r += s_warp[w];
    return r;  // valid in thread 0
}

constexpr int kDigestThreads = 256;
constexpr int kDigestSplit = 8;  // CTAs per chunk

// bytes: mix(n) + sum_i mix(i) * (W_i + 1), W_i = little-endian u64 of bytes [8i, 8i+8), zero padded
__global__ void __launch_bounds__(kDigestThreads) pp_digest_bytes_kernel(const uint8_t *__restrict__ slots,
                                                                         const ChunkDesc *__restrict__ descs,
                                                                         const ChunkResult *__restrict__ results,
                                                                         int n, unsigned long long *__restrict__ out)
{
    __shared__ unsigned long long s_warp[kDigestThreads / 32];
    const int k = (int)blockIdx.x / kDigestSplit, part = (int)blockIdx.x % kDigestSplit;
    if (k >= n) return;
    const ChunkDesc d = descs[k];
    const ChunkResult r = results[k];
    const uint64_t nbytes = r.status == 0 ? r.produced : 0u;
    // the output starts 16-byte aligned: slot_off is a multiple of 128, lead_len of 16
    const uint8_t *p = slots + d.slot_off + d.lead_len;
    const uint64_t nvec = nbytes / 16;
    unsigned long long acc = 0;
    for (uint64_t v = (uint64_t)part * kDigestThreads + threadIdx.x; v < nvec; v += (uint64_t)kDigestSplit * kDigestThreads) {
        const uint4 q = *reinterpret_cast<const uint4 *>(p + 16 * v);
        const unsigned long long w0 = (unsigned long long)q.x | ((unsigned long long)q.y << 32);
        const unsigned long long w1 = (unsigned long long)q.z | ((unsigned long long)q.w << 32);
        acc += mix64(2 * v) * (w0 + 1) + mix64(2 * v + 1) * (w1 + 1);
    }
    if (part == 0 && threadIdx.x == 0) {
        acc += mix64(nbytes);
        // tail: fewer than 16 bytes, as one or two zero-padded words
        const uint64_t done = nvec * 16;
        for (uint64_t wi = done / 8; wi * 8 < nbytes; wi++) {
            unsigned long long w = 0;
            for (int b = 0; b < 8; b++)
                if (wi * 8 + (uint64_t)b < nbytes) w |= (unsigned long long)p[wi * 8 + b] << (8 * b);
            acc += mix64(wi) * (w + 1);
        }
    }
    const unsigned long long tot = block_sum_u64(acc, s_warp);
    if (threadIdx.x == 0 && tot) atomicAdd(&out[2 * k], tot);
}

// fields: sum_r sum_{f<9} mix(9r+f) * (field + 1): start, idnFrom, idnLen, seqFrom, seqLen, plsFrom,
// plsLen, qltFrom, qltLen from the SoA line starts (see pp_parse in ppb200.h for the mapping)
__global__ void __launch_bounds__(kDigestThreads) pp_digest_fields_kernel(const ParseDesc *__restrict__ pdesc,
                                                                          const ParseOut *__restrict__ pout, int n,
                                                                          const uint32_t *__restrict__ lines,
                                                                          int64_t stride,
                                                                          unsigned long long *__restrict__ out)
{
    __shared__ unsigned long long s_warp[kDigestThreads / 32];
    const int k = (int)blockIdx.x / kDigestSplit, part = (int)blockIdx.x % kDigestSplit;
    if (k >= n) return;
    const ParseDesc pd = pdesc[k];
    const uint32_t *l0 = lines + pd.rec_base, *l1 = l0 + stride, *l2 = l1 + stride, *l3 = l2 + stride;
    const long long pend = (long long)pout[k].parse_end;
    unsigned long long acc = 0;
    for (uint32_t r = (uint32_t)part * kDigestThreads + threadIdx.x; r < pd.rec_count; r += kDigestSplit * kDigestThreads) {
        const long long a = l0[r], b = l1[r], c = l2[r], e = l3[r];
        const long long nx = r + 1u < pd.rec_count ? (long long)l0[r + 1] : pend;
        const long long f[9] = {a + 1, a + 1, b - a - 2, b, c - b - 1, c + 1, e - c - 2, e, nx - e - 1};
#pragma unroll
        for (int i = 0; i < 9; i++) acc += mix64(9ull * r + (unsigned)i) * ((unsigned long long)f[i] + 1ull);
    }
    const unsigned long long tot = block_sum_u64(acc, s_warp);
    if (threadIdx.x == 0 && tot) atomicAdd(&out[2 * k + 1], tot);
}

// out: 2 words per chunk (bytes digest, fields digest), zeroed here
cudaError_t launch_digests(const uint8_t *slots, const ChunkDesc *descs, const ChunkResult *results,
                           const ParseDesc *pdesc, const ParseOut *pout, int n, const uint32_t *lines,
                           int64_t line_stride, unsigned long long *out, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(out, 0, (size_t)n * 2 * sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    pp_digest_bytes_kernel<<<n * kDigestSplit, kDigestThreads, 0, st>>>(slots, descs, results, n, out);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    pp_digest_fields_kernel<<<n * kDigestSplit, kDigestThreads, 0, st>>>(pdesc, pout, n, lines, line_stride, out);
    return cudaGetLastError();
}

}  // namespace pp

// Device-side interfaces shared by inflate.cu, parse.cu and runtime.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "inflate_core.cuh"

namespace ppinf {
struct BlockRec;  // blockscan_core.cuh
}

namespace pp {

using ppinf::ChunkDesc;
using ppinf::ChunkResult;

// Per-chunk input of the parse stage (written by the record-base scan kernel).
struct ParseDesc {
    uint64_t data_off;   // byte offset of the chunk's combined memory (from.offset ++ inflated) in the slots buffer
    int64_t rec_base;    // index of the chunk's first record in the line-start arrays
    uint32_t total;      // |from.offset| + bytes inflated
    uint32_t rec_count;  // records the chunk yields
    uint32_t skip;       // leading records dropped (PP_JOB_STRICT, quirk H1)
    uint32_t exact;      // 1: route through the exact (sequential, quirk-exact) parser
};

// Per-chunk output of the parse stage.
struct ParseOut {
    uint32_t parse_end;  // combined-memory index just past the last record
    uint32_t flags;      // bit0: anomaly found by the fast parser (needs the exact parser)
    uint32_t newlines;   // '\n' seen by the fast parser (cross-check against the inflate count)
    uint32_t records;    // records emitted
};

struct ScanTotals {
    int64_t total_records;
    int64_t total_bytes;
    int64_t scanned_bytes;
    int32_t overflow;   // total_records > capacity: nothing was parsed
    int32_t first_status;
    int32_t exact_chunks;
    int32_t pad;
};

// kernel launchers (each returns the cudaGetLastError() of its launch)
// Launch geometry + scratch of the inflate kernel (owned by the device context).
struct InflateLaunch {
    int threads = 0;         // CTA size (multiple of 32, <= 1024)
    int grid = 0;            // resident CTAs (SMs x CTAs per SM)
    uint32_t *map = nullptr; // token/index scratch: grid x scratch_words_for(threads) words
    int *counter = nullptr;  // chunk counter the CTAs pull work from
};
// Optional host <-> kernel hand-shakes of the inflate kernel.
struct InflateSync {
    ppinf::ByteGate gate = {nullptr, 0, 0};  // pipelined upload: which bytes are in place (mark == null: all)
    uint32_t *done = nullptr;                             // mapped pinned host memory: done[k] = 1 when chunk k's bytes are final
    bool pull = false;                                    // comp is pinned host memory: the kernel variant that re-uses staged bytes
};
int inflate_max_ctas_per_sm(int threads);
cudaError_t inflate_set_max_smem(int threads);
size_t inflate_scratch_bytes(int threads, int grid);
cudaError_t launch_inflate(const ChunkDesc *descs, int n, const uint8_t *comp, uint64_t comp_bytes, uint8_t *slots,
                           const uint8_t *lead, ChunkResult *results, const InflateLaunch &cfg, cudaStream_t st,
                           InflateSync sy = InflateSync());
cudaError_t launch_inflate_dual(const ChunkDesc *descs, int n, const uint8_t *comp, uint64_t comp_bytes, uint8_t *slots,
                                const uint8_t *lead, ChunkResult *results, const InflateLaunch &cfg, uint64_t slot_delta,
                                uint64_t lead_delta, cudaStream_t st);
cudaError_t launch_bytes_stats(const uint8_t *slots, const ChunkDesc *descs, ChunkResult *results, int n,
                               cudaStream_t st);
cudaError_t launch_scan(const ChunkDesc *descs, const ChunkResult *results, const int64_t *exact_counts, int n,
                        uint32_t strict, int64_t capacity, ParseDesc *pdesc, ParseOut *pout, ScanTotals *totals,
                        cudaStream_t st);
uint32_t parse_tile_bytes();
cudaError_t launch_parse(const uint8_t *slots, const ParseDesc *pdesc, int n, const uint32_t *tile_base,
                         uint32_t total_tiles, uint32_t max_tiles, uint32_t *lines, int64_t line_stride,
                         ParseOut *pout, const ScanTotals *totals, unsigned long long *work, int sm_count,
                         int per_sm, cudaStream_t st);
int parse_max_ctas_per_sm();
cudaError_t launch_base_histogram(const uint8_t *slots, const ParseDesc *pdesc, int n, const uint32_t *lines,
                                  int64_t line_stride, unsigned long long *counts, int sm_count, cudaStream_t st);
cudaError_t launch_pattern_count(const uint8_t *slots, const ParseDesc *pdesc, int n, const uint32_t *lines,
                                 int64_t line_stride, const uint8_t *pattern, int plen, unsigned long long *count,
                                 int sm_count, cudaStream_t st);
cudaError_t launch_exact_count(const uint8_t *slots, const ChunkDesc *descs, const ChunkResult *results,
                               const ParseOut *pout, int n, int64_t *exact_counts, cudaStream_t st);
cudaError_t launch_exact_emit(const uint8_t *slots, const ParseDesc *pdesc, int n, uint32_t *lines,
                              int64_t line_stride, ParseOut *pout, const ScanTotals *totals, cudaStream_t st);
cudaError_t launch_digests(const uint8_t *slots, const ChunkDesc *descs, const ChunkResult *results,
                           const ParseDesc *pdesc, const ParseOut *pout, int n, const uint32_t *lines,
                           int64_t line_stride, unsigned long long *out, cudaStream_t st);


// blockscan.cu: the block scan on a stream already resident in device memory (used by pp_scan_blocks and
// pp_index_create_gpu; h_gz: the host copy, may be null), and the length of a gzip member header (0: not gzip).
int scan_blocks_resident(int device, int sm_count, cudaStream_t st, const uint8_t *d_comp, const uint8_t *h_gz, size_t gz_len,
                         size_t hdr, int64_t segment_bytes, std::vector<ppinf::BlockRec> &chain, uint64_t &land, uint64_t &total_out,
                         float &ms_total, int &npass);
size_t gzip_member_header_len(const uint8_t *gz, size_t n);

}  // namespace pp

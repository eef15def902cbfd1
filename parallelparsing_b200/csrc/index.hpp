// Host-side checkpoint index (Common/Index.cs:5-82) in structure-of-arrays form.
//
// The reference keeps a List<Point> with one 32 KB byte[] per point; here the
// windows of all points live in ONE contiguous allocation (count x 32768) so the
// whole set can be pinned once and shipped to the GPU with a single copy.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "ppb200.h"

struct pp_index {
    std::vector<int64_t> output;   // Point.Output
    std::vector<int64_t> input;    // Point.Input
    std::vector<int32_t> bits;     // Point.Bits
    std::vector<int64_t> off_pos;  // start of Point.offset inside `offsets`
    std::vector<int32_t> off_len;  // Point.offset.Length
    std::vector<uint8_t> offsets;  // all Point.offset arrays back to back
    uint8_t *windows = nullptr;    // count x PP_WINSIZE, page aligned
    size_t win_cap = 0;            // capacity in points
    int32_t chunk_max_bytes = 0;   // Index.ChunkMaxBytes
    mutable void *pinned_base = nullptr;  // set by the runtime when `windows` is cudaHostRegister'ed
    mutable size_t pinned_bytes = 0;
    // Compact form of the windows (zlib streams, as IndexIO version 1 stores them), built on first
    // use by a job with PP_JOB_COMPACT_WINDOWS: what crosses PCIe instead of 32 KB per checkpoint.
    // cwin_off has count+1 entries; every stream starts 16-byte aligned.
    mutable std::mutex cw_mu;
    mutable std::vector<uint8_t> cwin;
    mutable std::vector<uint64_t> cwin_off;
    mutable int32_t cwin_points = -1;      // points covered by cwin (-1: not built)
    mutable void *cwin_pinned = nullptr;   // set by the runtime when cwin is cudaHostRegister'ed

    int32_t count() const { return (int32_t)output.size(); }
    const uint8_t *window(int32_t i) const { return windows + (size_t)i * PP_WINSIZE; }
    const uint8_t *offset(int32_t i) const { return offsets.data() + off_pos[(size_t)i]; }
    uint8_t *append_window(bool zero = true);  // returns the (zeroed) window slot of the point being added
    void reserve_windows(size_t points);  // capacity for that many points in one allocation
    ~pp_index();
};

// Runtime hook: called before `windows` is reallocated or freed so a pinned
// registration can be dropped (implemented in runtime.cu).
extern "C" void pp_internal_unpin_index(const pp_index *ix);
extern "C" void pp_internal_unpin_cwin(const pp_index *ix);
// Build (once; thread safe) the compact windows.  Returns false when out of memory.
bool index_build_compact_windows(const pp_index *ix);

// ---- GPU CreateIndex (createindex.cu): the host half of Core.BuildDeflateIndex over per-block statistics ----
// One deflate block as the device reports it: where it starts (bit, output offset) and what its output
// holds in '@' bytes (Core.cs:86 counts every byte 0x40 as a record start).
struct CiBlockStat {
    uint64_t bit;      // first bit of the block header = the Z_BLOCK stop in front of the block
    uint64_t out;      // output offset of the block's first byte
    uint32_t ats;      // '@' bytes in the block's output
    uint32_t first;    // first / last '@', relative to the block's first byte (0xffffffff: none)
    uint32_t last;
    uint32_t maxgap;   // largest distance between two consecutive '@' inside the block
};
// One checkpoint to take: Point.Bits / Input / Output, and where its `offset` bytes start in the output
// (off_from == output: no offset).
struct CiPointPlan {
    int32_t bits;
    int64_t input, output, off_from;
};
// Core.cs:98-125 over the block list: which stops become points.  PP_E_RECORD_TOO_LONG as Core.cs:93 would
// throw (unless PP_INDEX_LIFT_RECORD_CAP).  total_in = the member's length in bytes (header .. trailer).
int index_plan_points(const CiBlockStat *b, size_t nb, uint64_t total_out, uint64_t total_in, uint32_t chunksize,
                      uint32_t flags, std::vector<CiPointPlan> &plan);
// Size `ix` for the planned points (scalars filled in, windows zeroed, offsets sized); the caller then
// writes the bytes into ix->windows / ix->offsets.
void index_from_plan(pp_index *ix, const std::vector<CiPointPlan> &plan);

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
    uint8_t *append_window();  // returns the (zeroed) window slot of the point being added
    ~pp_index();
};

// Runtime hook: called before `windows` is reallocated or freed so a pinned
// registration can be dropped (implemented in runtime.cu).
extern "C" void pp_internal_unpin_index(const pp_index *ix);
extern "C" void pp_internal_unpin_cwin(const pp_index *ix);
// Build (once; thread safe) the compact windows.  Returns false when out of memory.
bool index_build_compact_windows(const pp_index *ix);

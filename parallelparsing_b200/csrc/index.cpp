// Host side of the C ABI: Index, IndexIO and CreateIndex.
//
// CreateIndex stays on the CPU exactly as in the reference, whose
// Core.BuildDeflateIndex is a serial zlib inflate(Z_BLOCK) scan
// (Decompressor/Core.cs:14-131) bound through [DllImport("libz")]
// (Interop/PlatformInterop.cs:9-34).  This file makes the same zlib calls but
// is organised by produced SPAN instead of by byte: '@' bytes are counted and
// the last one located per span, and the partial record before a checkpoint
// (Point.offset) is kept in a growable buffer that is reset at each '@'.
#include "index.hpp"

#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <memory>
#include <new>
#include <thread>

static const int32_t kChunk = 16384;  // Common/Constants.cs:12

pp_index::~pp_index()
{
    pp_internal_unpin_cwin(this);
    if (windows) {
        pp_internal_unpin_index(this);
        free(windows);
    }
}

// zlib-compress every window (level 6, the streams IndexIO version 1 stores) on all host cores.
bool index_build_compact_windows(const pp_index *ix)
{
    std::lock_guard<std::mutex> lk(ix->cw_mu);
    const int32_t n = ix->count();
    if (ix->cwin_points == n) return true;
    pp_internal_unpin_cwin(ix);
    try {
        const size_t bound = (compressBound(PP_WINSIZE) + 15) & ~(size_t)15;
        std::vector<uint8_t> tmp((size_t)n * bound);
        std::vector<uint32_t> len((size_t)n, 0);
        std::atomic<int32_t> next{0};
        std::atomic<bool> ok{true};
        auto work = [&]() {
            for (;;) {
                const int32_t i = next.fetch_add(1);
                if (i >= n) break;
                uLongf m = (uLongf)bound;
                if (compress2(tmp.data() + (size_t)i * bound, &m, ix->window(i), PP_WINSIZE, 6) != Z_OK) ok = false;
                len[(size_t)i] = (uint32_t)m;
            }
        };
        unsigned nt = std::thread::hardware_concurrency();
        nt = nt ? (nt > 32 ? 32 : nt) : 1;
        if ((unsigned)n < nt) nt = n > 0 ? (unsigned)n : 1;
        std::vector<std::thread> th;
        for (unsigned t = 1; t < nt; t++) th.emplace_back(work);
        work();
        for (auto &t : th) t.join();
        if (!ok) return false;
        ix->cwin_off.assign((size_t)n + 1, 0);
        uint64_t acc = 0;
        for (int32_t i = 0; i < n; i++) {
            ix->cwin_off[(size_t)i] = acc;
            acc += ((uint64_t)len[(size_t)i] + 15u) & ~(uint64_t)15;
        }
        ix->cwin_off[(size_t)n] = acc;
        ix->cwin.assign((size_t)acc + 4096, 0);
        for (int32_t i = 0; i < n; i++)
            memcpy(ix->cwin.data() + ix->cwin_off[(size_t)i], tmp.data() + (size_t)i * bound, len[(size_t)i]);
        ix->cwin_points = n;
    } catch (...) {
        return false;
    }
    return true;
}

// First touch of a large fresh allocation, by several threads: the kernel maps the pages while the device is still
// busy, instead of one fault after the other under the copy that fills them (32 MB of windows at 10 M reads:
// 3-15 ms single-threaded, 320 MB at 100 M reads: 100 ms).
static void prefault(uint8_t *p, size_t bytes)
{
    if (bytes < (8u << 20)) return;
    const unsigned nt = std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
    const size_t per = ((bytes + nt - 1) / nt + 4095) & ~(size_t)4095;
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++)
        th.emplace_back([=]() {
            const size_t lo = (size_t)t * per, hi = std::min(bytes, lo + per);
            for (size_t o = lo; o < hi; o += 4096) ((volatile uint8_t *)p)[o] = 0;
        });
    for (auto &x : th) x.join();
}

void pp_index::reserve_windows(size_t points)
{
    if (points <= win_cap) return;
    void *p = nullptr;
    if (posix_memalign(&p, 4096, points * (size_t)PP_WINSIZE) != 0) throw std::bad_alloc();
    prefault((uint8_t *)p, points * (size_t)PP_WINSIZE);
    if (windows) {
        pp_internal_unpin_index(this);
        memcpy(p, windows, std::min(win_cap, output.size()) * (size_t)PP_WINSIZE);
        free(windows);
    }
    windows = (uint8_t *)p;
    win_cap = points;
}

uint8_t *pp_index::append_window(bool zero)
{
    cwin_points = -1;  // the compact windows no longer cover every point
    size_t n = output.size();  // caller has already pushed the scalar fields
    if (n > win_cap) {
        size_t cap = win_cap ? win_cap * 2 : 64;
        while (cap < n) cap *= 2;
        void *p = nullptr;
        if (posix_memalign(&p, 4096, cap * (size_t)PP_WINSIZE) != 0) throw std::bad_alloc();
        if (windows) {
            pp_internal_unpin_index(this);
            memcpy(p, windows, win_cap * (size_t)PP_WINSIZE);
            free(windows);
        }
        windows = (uint8_t *)p;
        win_cap = cap;
    }
    uint8_t *w = windows + (n - 1) * (size_t)PP_WINSIZE;
    if (zero) memset(w, 0, PP_WINSIZE);
    return w;
}

// Index.AddPoint — Common/Index.cs:24-48.  `window` is the 32 KB circular
// inflate window, `left` the free space behind the write position; the stored
// Window is the window un-rotated into stream order (:42-46).
static void add_point(pp_index *ix, int32_t bits, int64_t input, int64_t output, uint32_t left,
                      const uint8_t *window, const uint8_t *offset, int32_t offset_len)
{
    if (ix->count() == 0) {
        ix->chunk_max_bytes = (int32_t)output;  // :27-30
    } else {
        int32_t sz = (int32_t)output - (int32_t)ix->output.back();  // :33 (int arithmetic)
        if (sz > ix->chunk_max_bytes) ix->chunk_max_bytes = sz;
    }
    ix->output.push_back(output);
    ix->input.push_back(input);
    ix->bits.push_back(bits);
    ix->off_pos.push_back((int64_t)ix->offsets.size());
    ix->off_len.push_back(offset_len > 0 ? offset_len : 0);
    if (offset_len > 0) ix->offsets.insert(ix->offsets.end(), offset, offset + offset_len);
    uint8_t *w = ix->append_window();
    if (window) {
        if (left > PP_WINSIZE) left = PP_WINSIZE;
        if (left != 0) memcpy(w, window + PP_WINSIZE - left, left);
        if (left < PP_WINSIZE) memcpy(w + left, window, PP_WINSIZE - left);
    }
}

int index_plan_points(const CiBlockStat *b, size_t nb, uint64_t total_out, uint64_t total_in, uint32_t chunksize,
                      uint32_t flags, std::vector<CiPointPlan> &plan)
{
    const bool lift = (flags & PP_INDEX_LIFT_RECORD_CAP) != 0;
    const int64_t threshold = (int64_t)(uint32_t)(chunksize - 8u);  // Core.cs:105, uint arithmetic
    int64_t records = 0;
    uint64_t run_start = 0;  // where offsetBeforePoint starts: the last '@' so far, or the stream's first byte
    plan.clear();
    for (size_t i = 0; i < nb; i++) {
        // the stop in front of block i: the output of blocks < i has been seen (Core.cs:78-95)
        if (i) {
            const CiBlockStat &p = b[i - 1];
            if (p.ats) {
                if (!lift && ((p.out + p.first) - run_start > (uint64_t)PP_WINSIZE || p.maxgap > (uint32_t)PP_WINSIZE))
                    return PP_E_RECORD_TOO_LONG;  // Core.cs:93
                run_start = p.out + p.last;
                records += p.ats;
            }
            if (!lift && b[i].out - run_start > (uint64_t)PP_WINSIZE) return PP_E_RECORD_TOO_LONG;
        }
        const int32_t bits = (int32_t)((8u - (uint32_t)(b[i].bit & 7u)) & 7u);
        const int64_t input = (int64_t)((b[i].bit + 7u) >> 3);
        if (b[i].out == 0) {  // Core.cs:99-102
            plan.push_back({bits, input, 0, 0});
        } else if (records > threshold) {  // Core.cs:105-109
            plan.push_back({bits, input, (int64_t)b[i].out, (int64_t)run_start});
            records = 0;
        }
    }
    if (nb) {
        const CiBlockStat &p = b[nb - 1];
        if (p.ats) {
            if (!lift && ((p.out + p.first) - run_start > (uint64_t)PP_WINSIZE || p.maxgap > (uint32_t)PP_WINSIZE))
                return PP_E_RECORD_TOO_LONG;
            run_start = p.out + p.last;
        }
    }
    if (!lift && total_out - run_start > (uint64_t)PP_WINSIZE) return PP_E_RECORD_TOO_LONG;
    plan.push_back({0, (int64_t)total_in, (int64_t)total_out, (int64_t)total_out});  // Core.cs:114-125, Z_STREAM_END
    return PP_OK;
}

void index_from_plan(pp_index *ix, const std::vector<CiPointPlan> &plan)
{
    ix->reserve_windows(ix->output.size() + plan.size());
    size_t off_total = ix->offsets.size();
    for (const CiPointPlan &p : plan) {   // add_point's bookkeeping (Index.cs:24-48) without touching the window bytes:
        const int64_t n = p.output - p.off_from;  // the caller overwrites every one of them
        if (ix->count() == 0) {
            ix->chunk_max_bytes = (int32_t)p.output;
        } else {
            const int32_t sz = (int32_t)p.output - (int32_t)ix->output.back();
            if (sz > ix->chunk_max_bytes) ix->chunk_max_bytes = sz;
        }
        ix->output.push_back(p.output);
        ix->input.push_back(p.input);
        ix->bits.push_back(p.bits);
        ix->off_pos.push_back((int64_t)off_total);
        ix->off_len.push_back((int32_t)n);
        off_total += (size_t)n;
        ix->append_window(false);
    }
    ix->offsets.resize(off_total);
}

extern "C" {

int pp_abi_version(void) { return PP_ABI_VERSION; }

const char *pp_strerror(int code)
{
    switch (code) {
        case PP_OK: return "ok";
        case PP_STREAM_END: return "stream end";
        case PP_NEED_DICT: return "need dictionary";
        case PP_ERRNO: return "file error";
        case PP_STREAM_ERROR: return "stream error";
        case PP_DATA_ERROR: return "data error";
        case PP_MEM_ERROR: return "insufficient memory";
        case PP_BUF_ERROR: return "buffer error";
        case PP_VERSION_ERROR: return "incompatible version";
        case PP_E_CUDA: return "CUDA error";
        case PP_E_NO_DEVICE: return "no sm_100 CUDA device (this library has no CPU fallback)";
        case PP_E_ARG: return "bad argument";
        case PP_E_IO: return "I/O error";
        case PP_E_RECORD_TOO_LONG: return "record longer than 32768 bytes (reference: IndexOutOfRangeException)";
        case PP_E_FORMAT: return "malformed index file";
        case PP_E_UNSUPPORTED: return "input not supported by this entry point";
        default: return "unknown error";
    }
}

int pp_index_new(pp_index **out)
{
    if (!out) return PP_E_ARG;
    *out = new (std::nothrow) pp_index();
    return *out ? PP_OK : PP_MEM_ERROR;
}

void pp_index_free(pp_index *ix) { delete ix; }

int pp_index_add_point(pp_index *ix, int32_t bits, int64_t input, int64_t output, uint32_t left,
                       const uint8_t *window, const uint8_t *offset, int32_t offset_len)
{
    if (!ix || (offset_len > 0 && !offset)) return PP_E_ARG;
    try {
        add_point(ix, bits, input, output, left, window, offset, offset_len);
    } catch (...) {
        return PP_MEM_ERROR;
    }
    return PP_OK;
}

int pp_index_add(pp_index *ix, int32_t bits, int64_t input, int64_t output, const uint8_t *window,
                 const uint8_t *offset, int32_t offset_len)
{
    if (!ix || (offset_len > 0 && !offset)) return PP_E_ARG;
    const int32_t keep = ix->chunk_max_bytes;
    try {
        add_point(ix, bits, input, output, 0, window, offset, offset_len);
    } catch (...) {
        return PP_MEM_ERROR;
    }
    ix->chunk_max_bytes = keep;
    return PP_OK;
}

// Test hook (not part of the ABI): index_plan_points over flat arrays.  stats = nb x {ats, first, last, maxgap},
// plan = cap x {bits, input, output, off_from}.
int pp_internal_plan_points(const uint64_t *bit, const uint64_t *out, const uint32_t *stats, int64_t nb, uint64_t total_out,
                            uint64_t total_in, uint32_t chunksize, uint32_t flags, int64_t *plan, int64_t cap, int64_t *count)
{
    if (nb < 0 || !count) return PP_E_ARG;
    try {
        std::vector<CiBlockStat> b((size_t)nb);
        for (int64_t i = 0; i < nb; i++)
            b[(size_t)i] = {bit[i], out[i], stats[4 * i], stats[4 * i + 1], stats[4 * i + 2], stats[4 * i + 3]};
        std::vector<CiPointPlan> pl;
        const int rc = index_plan_points(b.data(), (size_t)nb, total_out, total_in, chunksize, flags, pl);
        if (rc != PP_OK) return rc;
        *count = (int64_t)pl.size();
        for (int64_t i = 0; i < (int64_t)pl.size() && i < cap; i++) {
            plan[4 * i] = pl[(size_t)i].bits;
            plan[4 * i + 1] = pl[(size_t)i].input;
            plan[4 * i + 2] = pl[(size_t)i].output;
            plan[4 * i + 3] = pl[(size_t)i].off_from;
        }
        return PP_OK;
    } catch (...) {
        return PP_MEM_ERROR;
    }
}

int32_t pp_index_count(const pp_index *ix) { return ix ? ix->count() : 0; }
int32_t pp_index_chunk_max_bytes(const pp_index *ix) { return ix ? ix->chunk_max_bytes : 0; }

int pp_index_point(const pp_index *ix, int32_t i, pp_point *out)
{
    if (!ix || !out || i < 0 || i >= ix->count()) return PP_E_ARG;
    out->output = ix->output[(size_t)i];
    out->input = ix->input[(size_t)i];
    out->bits = ix->bits[(size_t)i];
    out->offset_len = ix->off_len[(size_t)i];
    out->window = ix->window(i);
    out->offset = ix->offset(i);
    return PP_OK;
}

// CreateIndex.  See the file header; statement references are to Decompressor/Core.cs.
int pp_index_create(const uint8_t *gz, size_t gz_len, uint32_t chunksize, uint32_t flags, pp_index **out)
{
    if (!out || (!gz && gz_len)) return PP_E_ARG;
    *out = nullptr;
    std::unique_ptr<pp_index> ix(new (std::nothrow) pp_index());
    if (!ix) return PP_MEM_ERROR;

    z_stream strm;
    memset(&strm, 0, sizeof strm);
    int ret = inflateInit2(&strm, 47);  // :30 automatic gzip decoding
    if (ret != Z_OK) return ret;

    std::vector<uint8_t> window(PP_WINSIZE, 0);
    std::vector<uint8_t> partial;  // bytes since the last '@' (offsetBeforePoint, :23,86-94)
    partial.reserve(PP_WINSIZE);
    const bool lift = (flags & PP_INDEX_LIFT_RECORD_CAP) != 0;
    // `recordCounter > chunksize - 8` compares int with uint, i.e. as long (:105);
    // chunksize < 8 wraps and never triggers.
    const int64_t threshold = (int64_t)(uint32_t)(chunksize - 8u);
    int64_t records = 0;
    int64_t totin = 0, totout = 0;
    size_t fed = 0;
    int rc = PP_OK;
    bool done = false;

    try {
        while (!done) {
            // :41 file.Read(input, 0, CHUNK) — feed at most CHUNK bytes at a time so that
            // inflate stops at the same places the reference's does.
            size_t n = gz_len - fed < (size_t)kChunk ? gz_len - fed : (size_t)kChunk;
            if (n == 0) { rc = PP_DATA_ERROR; break; }  // :42-45
            strm.next_in = const_cast<Bytef *>(gz + fed);
            strm.avail_in = (uInt)n;
            fed += n;
            do {
                if (strm.avail_out == 0) {  // :52-56
                    strm.avail_out = PP_WINSIZE;
                    strm.next_out = window.data();
                }
                const uint32_t in_before = strm.avail_in, out_before = strm.avail_out;
                ret = inflate(&strm, Z_BLOCK);  // :64
                totin += in_before - strm.avail_in;
                totout += out_before - strm.avail_out;
                if (ret == Z_NEED_DICT || ret == Z_MEM_ERROR || ret == Z_DATA_ERROR || ret == Z_STREAM_ERROR ||
                    ret == Z_BUF_ERROR || ret == Z_VERSION_ERROR) {  // :68-74
                    rc = ret;
                    done = true;
                    break;
                }
                // :78-95 on the span just produced
                const uint8_t *span = window.data() + (PP_WINSIZE - out_before);
                const size_t span_len = out_before - strm.avail_out;
                if (span_len) {
                    size_t ats = 0;
                    for (size_t i = 0; i < span_len; i++) ats += (span[i] == 64);
                    size_t keep_from = 0;
                    // :93 indexes offsetBeforePoint byte by byte: the reference throws as soon as the running
                    // partial record needs a 32769th byte, even when an '@' follows later in the same span.
                    // Inside a span only the stretch before its first '@' can get there (a span is <= 32768 B).
                    if (!lift) {
                        const uint8_t *first = ats ? (const uint8_t *)memchr(span, 64, span_len) : nullptr;
                        const size_t head = first ? (size_t)(first - span) : span_len;
                        if (partial.size() + head > (size_t)PP_WINSIZE) {
                            rc = PP_E_RECORD_TOO_LONG;
                            done = true;
                            break;
                        }
                    }
                    if (ats) {
                        records += (int64_t)ats;
                        const uint8_t *last = (const uint8_t *)memrchr(span, 64, span_len);
                        keep_from = (size_t)(last - span);
                        partial.clear();
                    }
                    partial.insert(partial.end(), span + keep_from, span + span_len);
                    if (!lift && partial.size() > (size_t)PP_WINSIZE) {  // :93 would index past the array
                        rc = PP_E_RECORD_TOO_LONG;
                        done = true;
                        break;
                    }
                }
                if ((strm.data_type & 128) && !(strm.data_type & 64)) {  // :98 end of a non-final block
                    if (totout == 0) {
                        add_point(ix.get(), strm.data_type & 7, totin, totout, strm.avail_out, window.data(),
                                  nullptr, 0);  // :101-102
                    } else if (records > threshold) {  // :105-109
                        add_point(ix.get(), strm.data_type & 7, totin, totout, strm.avail_out, window.data(),
                                  partial.data(), (int32_t)partial.size());
                        records = 0;
                    }
                }
                if (ret == Z_STREAM_END) {  // :114-125
                    if (strm.avail_in != 0 || fed != gz_len) {
                        ret = inflateReset(&strm);
                        if (ret != Z_OK) { rc = ret; done = true; break; }
                        continue;
                    }
                    add_point(ix.get(), strm.data_type & 7, totin, totout, strm.avail_out, window.data(), nullptr, 0);
                    done = true;
                    break;
                }
            } while (strm.avail_in != 0);
        }
    } catch (...) {
        rc = PP_MEM_ERROR;
    }
    inflateEnd(&strm);
    if (rc != PP_OK) return rc;
    *out = ix.release();
    return PP_OK;
}

static int read_file(const char *path, std::vector<uint8_t> &buf)
{
    FILE *f = fopen(path, "rb");
    if (!f) return PP_E_IO;
    if (fseek(f, 0, SEEK_END) != 0) { fclose(f); return PP_E_IO; }
    long n = ftell(f);
    if (n < 0) { fclose(f); return PP_E_IO; }
    fseek(f, 0, SEEK_SET);
    buf.resize((size_t)n);
    size_t got = n ? fread(buf.data(), 1, (size_t)n, f) : 0;
    fclose(f);
    return got == (size_t)n ? PP_OK : PP_E_IO;
}

int pp_index_create_file(const char *gz_path, uint32_t chunksize, uint32_t flags, pp_index **out)
{
    if (!gz_path || !out) return PP_E_ARG;
    std::vector<uint8_t> buf;
    try {
        int rc = read_file(gz_path, buf);
        if (rc != PP_OK) return rc;
    } catch (...) {
        return PP_MEM_ERROR;
    }
    return pp_index_create(buf.data(), buf.size(), chunksize, flags, out);
}

// IndexIO.Serialize — Common/IndexIO.cs:7-27.  Layout (little endian, the code
// not the stale comment at :5-6): int32 0 | int32 ChunkMaxBytes | int32 Count |
// Count x { int64 Output | int64 Input | int32 Bits | int32 winLen | win | int32 offLen | off }.
//
// version 1 (extension, pp_index_serialize_v1): the leading int32 — reserved, written as 0 and
// ignored on read by the reference (IndexIO.cs:12,34) — is 1, and every window is stored
// zlib-compressed: `winLen` is the compressed length, the bytes inflate to exactly 32768.  The
// windows are the bulk of an index (32 KB per point; 27.6 % of the .gz size at chunk 1000), and
// they are FASTQ text.  Files the reference must read stay version 0.
static int serialize_index(const pp_index *ix, const char *path, int version)
{
    if (!ix || !path) return PP_E_ARG;
    FILE *f = fopen(path, "wb");
    if (!f) return PP_E_IO;
    bool ok = true;
    auto put32 = [&](int32_t v) { ok = ok && fwrite(&v, 4, 1, f) == 1; };
    auto put64 = [&](int64_t v) { ok = ok && fwrite(&v, 8, 1, f) == 1; };
    std::vector<uint8_t> packed;
    try {
        if (version == 1) packed.resize(compressBound(PP_WINSIZE));
    } catch (...) {
        fclose(f);
        return PP_MEM_ERROR;
    }
    put32(version);
    put32(ix->chunk_max_bytes);
    put32(ix->count());
    for (int32_t i = 0; i < ix->count() && ok; i++) {
        put64(ix->output[(size_t)i]);
        put64(ix->input[(size_t)i]);
        put32(ix->bits[(size_t)i]);
        if (version == 1) {
            uLongf n = (uLongf)packed.size();
            ok = ok && compress2(packed.data(), &n, ix->window(i), PP_WINSIZE, 6) == Z_OK;
            put32((int32_t)n);
            ok = ok && fwrite(packed.data(), 1, (size_t)n, f) == (size_t)n;
        } else {
            put32(PP_WINSIZE);
            ok = ok && fwrite(ix->window(i), 1, PP_WINSIZE, f) == PP_WINSIZE;
        }
        const int32_t ol = ix->off_len[(size_t)i];
        put32(ol);
        if (ol) ok = ok && fwrite(ix->offset(i), 1, (size_t)ol, f) == (size_t)ol;
    }
    ok = (fclose(f) == 0) && ok;
    return ok ? PP_OK : PP_E_IO;
}
int pp_index_serialize(const pp_index *ix, const char *path) { return serialize_index(ix, path, 0); }
int pp_index_serialize_v1(const pp_index *ix, const char *path) { return serialize_index(ix, path, 1); }

// IndexIO.Deserialize — Common/IndexIO.cs:29-53.  As in the reference the stored
// ChunkMaxBytes is read and discarded (:35,52): the new Index(points) has 0.
int pp_index_deserialize(const char *path, pp_index **out)
{
    if (!path || !out) return PP_E_ARG;
    *out = nullptr;
    FILE *f = fopen(path, "rb");
    if (!f) return PP_E_IO;
    std::unique_ptr<pp_index> ix(new (std::nothrow) pp_index());
    if (!ix) { fclose(f); return PP_MEM_ERROR; }
    int rc = PP_OK;
    int32_t hdr[3];
    if (fread(hdr, 4, 3, f) != 3 || hdr[2] < 0) rc = PP_E_FORMAT;
    try {
        const bool v1 = rc == PP_OK && hdr[0] == 1;  // any other value: version 0, as the reference reads it
        std::vector<uint8_t> win(PP_WINSIZE), off, packed(v1 ? compressBound(PP_WINSIZE) : 0);
        for (int32_t i = 0; rc == PP_OK && i < hdr[2]; i++) {
            int64_t output, input;
            int32_t bits, winlen, ol;
            if (fread(&output, 8, 1, f) != 1 || fread(&input, 8, 1, f) != 1 || fread(&bits, 4, 1, f) != 1 ||
                fread(&winlen, 4, 1, f) != 1) {
                rc = PP_E_FORMAT;
                break;
            }
            if (v1) {
                uLongf n = PP_WINSIZE;
                if (winlen <= 0 || (size_t)winlen > packed.size() ||
                    fread(packed.data(), 1, (size_t)winlen, f) != (size_t)winlen ||
                    uncompress(win.data(), &n, packed.data(), (uLong)winlen) != Z_OK || n != PP_WINSIZE) {
                    rc = PP_E_FORMAT;
                    break;
                }
            } else if (winlen != PP_WINSIZE || fread(win.data(), 1, PP_WINSIZE, f) != PP_WINSIZE) {
                rc = PP_E_FORMAT;
                break;
            }
            if (fread(&ol, 4, 1, f) != 1 || ol < 0) {
                rc = PP_E_FORMAT;
                break;
            }
            off.resize((size_t)ol);
            if (ol && fread(off.data(), 1, (size_t)ol, f) != (size_t)ol) { rc = PP_E_FORMAT; break; }
            // `new Point(output, input, bits, window, offset)` keeps the window as stored:
            // left = 0 makes add_point copy it unrotated.
            const int32_t keep = ix->chunk_max_bytes;
            add_point(ix.get(), bits, input, output, 0, win.data(), off.data(), ol);
            ix->chunk_max_bytes = keep;  // Deserialize never sets ChunkMaxBytes
        }
    } catch (...) {
        rc = PP_MEM_ERROR;
    }
    fclose(f);
    if (rc != PP_OK) return rc;
    ix->chunk_max_bytes = 0;
    *out = ix.release();
    return PP_OK;
}

}  // extern "C"

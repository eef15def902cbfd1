/*
 * ppb200.h — C ABI of libppb200.so: the B200-native checkpointed gzip-FASTQ
 * decode path (CreateIndex / Decompress(checkpoint) / DecompressAll /
 * Serialize / Deserialize) of Quantumzhao/ParallelParsing.
 *
 * This is the drop-in boundary.  The reference reaches its only native code
 * (system zlib) through a static class of [DllImport] externs
 * (Interop/PlatformInterop.cs:6-35) wrapped by the Compat facade
 * (Interop/Conventions.cs:129-196); this library is bound the same way (see
 * INTEGRATION.md for the C# stub).  Conventions kept from that boundary:
 *   - only blittable types cross it: pointers, sizes, int32/int64;
 *   - the caller owns every data buffer, native code owns only opaque handles
 *     released by an explicit *_free / pp_close (Conventions.cs:121-126);
 *   - every call returns a ZResult-compatible int (Conventions.cs:9-20):
 *     0 OK, 1 STREAM_END, 2 NEED_DICT, negatives are errors; the C# side turns
 *     negatives into ZException(code) as Core.cs:33,74,149,179 do today;
 *   - no callbacks, no errno, no exceptions across the boundary;
 *   - calls on one pp_ctx are serialised internally, so they may be issued from
 *     thread-pool tasks as BatchedFASTQ.cs:62 does (README.md:50).
 *
 * There is NO CPU fallback: every decode/parse entry point runs hand-written
 * sm_100a CUDA kernels and fails with PP_E_NO_DEVICE / PP_E_CUDA otherwise.
 * Index creation and (de)serialisation are host code, as in the reference
 * (Core.BuildDeflateIndex is a serial zlib Z_BLOCK scan, Core.cs:14-131).
 */
#ifndef PPB200_H
#define PPB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PP_ABI_VERSION 1
#define PP_WINSIZE 32768 /* Common/Constants.cs:9 */

/* ZResult (Interop/Conventions.cs:9-20) */
#define PP_OK 0
#define PP_STREAM_END 1
#define PP_NEED_DICT 2
#define PP_ERRNO (-1)
#define PP_STREAM_ERROR (-2)
#define PP_DATA_ERROR (-3)
#define PP_MEM_ERROR (-4)
#define PP_BUF_ERROR (-5)
#define PP_VERSION_ERROR (-6)
/* new codes, outside zlib's range */
#define PP_E_CUDA (-100)            /* a CUDA runtime call or kernel failed */
#define PP_E_NO_DEVICE (-101)       /* no usable sm_100 device: there is no CPU fallback */
#define PP_E_ARG (-102)             /* bad argument */
#define PP_E_IO (-103)              /* file open/read/write failed */
#define PP_E_RECORD_TOO_LONG (-104) /* Core.cs:93 IndexOutOfRangeException: record > 32768 B */
#define PP_E_FORMAT (-105)          /* malformed IndexIO file */
#define PP_E_UNSUPPORTED (-106)     /* input outside what this entry point handles (see the entry point) */

typedef struct pp_index pp_index; /* Common/Index.cs:5  Index  */
typedef struct pp_ctx pp_ctx;     /* one GPU + stream + scratch */
typedef struct pp_job pp_job;     /* one DecompressAll plan over a chunk range */

/* Common/Index.cs:51-82  Point (a view; pointers stay valid until the index is freed/modified) */
typedef struct pp_point {
    int64_t output;        /* Point.Output: offset in uncompressed data          */
    int64_t input;         /* Point.Input : offset in the file of first full byte */
    int32_t bits;          /* Point.Bits  : 0, or 1-7 bits taken from byte input-1 */
    int32_t offset_len;    /* Point.offset.Length                                 */
    const uint8_t *window; /* Point.Window: preceding 32768 uncompressed bytes    */
    const uint8_t *offset; /* Point.offset: partial record before the point       */
} pp_point;

int pp_abi_version(void);
const char *pp_strerror(int code);

/* ------------------------------------------------------------------ Index */

/* flags for pp_index_create */
#define PP_INDEX_LIFT_RECORD_CAP 1u /* extension: do not fail on records > 32768 B */

/* CreateIndex — Core.BuildDeflateIndex(FileStream, uint chunksize), Decompressor/Core.cs:14-131. */
int pp_index_create(const uint8_t *gz, size_t gz_len, uint32_t chunksize, uint32_t flags, pp_index **out);
int pp_index_create_file(const char *gz_path, uint32_t chunksize, uint32_t flags, pp_index **out);
/* new Index() — Common/Index.cs:11 */
int pp_index_new(pp_index **out);
/* Index.AddPoint(bits, input, output, left, window, offset) — Common/Index.cs:24-48 */
int pp_index_add_point(pp_index *ix, int32_t bits, int64_t input, int64_t output, uint32_t left,
                       const uint8_t *window, const uint8_t *offset, int32_t offset_len);
/* Index.Add(Point) — Common/Index.cs:22: append a finished point unchanged (ChunkMaxBytes untouched) */
int pp_index_add(pp_index *ix, int32_t bits, int64_t input, int64_t output, const uint8_t *window,
                 const uint8_t *offset, int32_t offset_len);
/* IndexIO.Serialize / Deserialize — Common/IndexIO.cs:7-27 / :29-53 (same bytes on disk) */
int pp_index_serialize(const pp_index *ix, const char *path);
int pp_index_deserialize(const char *path, pp_index **out);
/*
 * Extension: IndexIO file version 1.  The leading int32 of the file — reserved, written as 0 and
 * ignored on read by the reference (Common/IndexIO.cs:12,34) — is 1 and every 32 KB window is stored
 * zlib-compressed (`winLen` = compressed length).  pp_index_deserialize reads both versions; files
 * the reference itself must read have to stay version 0 (pp_index_serialize).
 */
int pp_index_serialize_v1(const pp_index *ix, const char *path);
/* Index.Count, Index.ChunkMaxBytes, Index[i] — Common/Index.cs:9,20,21 */
int32_t pp_index_count(const pp_index *ix);
int32_t pp_index_chunk_max_bytes(const pp_index *ix);
int pp_index_point(const pp_index *ix, int32_t i, pp_point *out);
void pp_index_free(pp_index *ix);

/* ----------------------------------------------------------------- Device */


/* Open CUDA device `device`.  Fails with PP_E_NO_DEVICE when there is no GPU.  Free every job created
 * on a context before closing it (a job's device memory is returned to the pool on its context's stream). */
int pp_open(int32_t device, pp_ctx **out);
void pp_close(pp_ctx *ctx);
/* Pinned host memory for the compressed file (cudaHostAlloc / cudaHostRegister). */
int pp_host_alloc(size_t bytes, void **out);
void pp_host_free(void *p);
int pp_host_register(void *p, size_t bytes);
void pp_host_unregister(void *p);

/* -------------------------------------------------- Decompress(checkpoint) */

/*
 * Core.ExtractDeflateIndex(fileBuffer, from, to, buf) — Decompressor/Core.cs:133-192.
 * `fileBuffer` is the byte range LazyFileReader hands over: file bytes
 * [from.Input-1, to.Input) (Decompressor/LazyFileReader.cs:63-69) where
 * from = index[from_point], to = index[from_point+1].  Inflates into `buf`
 * (capacity >= to.Output-from.Output).  Returns bytes produced, or a negative code.
 */
int64_t pp_extract(pp_ctx *ctx, const uint8_t *fileBuffer, int64_t fileBufferLen, const pp_index *ix,
                   int32_t from_point, uint8_t *buf, int64_t buf_len);

/* ------------------------------------------------------------ Parsing.Parse */

/*
 * Parsing.Parse(new CombinedMemory(prepend, rest)) — Decompressor/Parsing.cs:11-117.
 * `rest` is the data proper; the reference's zero tail (the rented array is
 * pow-2 sized, BatchedFASTQ.cs:65-68) is implied: parsing stops at the first
 * NUL or at the end of `rest`, and a trailing partial record is dropped.
 * For each record r < cap writes line_starts[4r..4r+3]: combined-memory index
 * of the first byte of the id line ('@'), sequence line, '+' line and quality
 * line; *parse_end receives the index just past the last record.  In the
 * reference's terms: start=idnFrom=l0+1, idnLen=l1-l0-2, seqFrom=l1,
 * seqLen=l2-l1-1, plsFrom=l2+1, plsLen=l3-l2-2, qltFrom=l3, qltLen=next_l0-l3-1.
 * Returns the record count (may exceed cap) or a negative code.
 */
int64_t pp_parse(pp_ctx *ctx, const uint8_t *prepend, int64_t prepend_len, const uint8_t *rest, int64_t rest_len,
                 uint32_t *line_starts, int64_t cap, uint32_t *parse_end);

/* ------------------------------------------------------------ DecompressAll */

/* flags for pp_job_create */
#define PP_JOB_STRICT 1u      /* extension: drop the duplicate record of quirk H1 */
/* kernels pull compressed bytes/windows straight from pinned host memory.  `gz` handed to
 * pp_job_upload must then be device-accessible pinned memory (pp_host_alloc, or any buffer passed
 * through pp_host_register); exactly gz_len bytes are readable, nothing behind them is touched. */
#define PP_JOB_ZEROCOPY 2u
/* staged mode only: pp_job_upload queues the compressed range in pieces on a copy stream and returns;
 * pp_job_execute's inflate kernel starts at once and each chunk waits (on the device) only until ITS
 * bytes are in place, so the PCIe copy hides under the decode (LazyFileReader's read-ahead,
 * LazyFileReader.cs:77-97, expressed on the device).  `gz` must stay valid until pp_job_download. */
#define PP_JOB_PIPELINE 4u
/* the checkpoint windows cross PCIe zlib-compressed (the form IndexIO version 1 stores,
 * pp_index_serialize_v1) and are inflated on the GPU into place by a pre-pass of the inflate kernel:
 * ~8 KB instead of 32 KB per checkpoint (the windows are +27 % of the bytes moved at chunk 1000). */
#define PP_JOB_COMPACT_WINDOWS 8u

typedef struct pp_job_info {
    int32_t first_chunk, n_chunks;
    int64_t total_records;   /* records over all chunks (canonical order: chunk, then in-chunk) */
    int64_t total_bytes;     /* inflated bytes over all chunks                                    */
    int64_t scanned_bytes;   /* sum over chunks of |offset_k| + inflated_k (bytes the parser reads) */
    int64_t compressed_bytes;/* compressed bytes consumed                                          */
    int64_t h2d_bytes;       /* bytes copied host->device by pp_job_upload                         */
    int64_t d2h_bytes;       /* bytes copied device->host by pp_job_download                       */
    int32_t status;          /* first non-zero chunk status, else 0                                */
    int32_t exact_chunks;    /* chunks that went through the exact (quirk-exact) parser            */
    float upload_ms, inflate_ms, scan_ms, parse_ms, download_ms; /* CUDA-event times of the last run */
    int32_t launches;        /* kernels launched by the last pp_job_execute                        */
} pp_job_info;

typedef struct pp_chunk_info {
    int32_t status;        /* 0 or negative ZResult of this chunk's inflate                      */
    int32_t prefix_len;    /* |from.offset|                                                      */
    int64_t inflated;      /* bytes produced (Core.ExtractDeflateIndex return value)             */
    int64_t records;       /* records Parsing.Parse yields for the chunk                         */
    int64_t record_base;   /* index of the chunk's first record in the line_start arrays         */
    uint32_t parse_end;    /* combined-memory index just past the chunk's last record            */
    uint32_t flags;        /* bit0: went through the exact parser                                */
} pp_chunk_info;

/*
 * DecompressAll — the BatchedFASTQ enumeration (Decompressor/BatchedFASTQ.cs:54-98)
 * over chunks [first_chunk, first_chunk+n_chunks) of the index (chunk k =
 * (index[k], index[k+1]), LazyFileReader.cs:53-61), split into phases so a
 * caller can keep the plan and inputs resident:
 *   create  : plan device layout, allocate device + pinned staging (no timing relevance)
 *   upload  : H2D of the compressed byte range and the checkpoint windows
 *   execute : inflate kernel -> record-base scan -> parse kernel (all device resident)
 *   download: D2H of the per-chunk results (status, counts)
 * `gz` is the whole .gz file in host memory (pinned for full H2D speed).
 * n_chunks < 0 means "to the last chunk".
 */
int pp_job_create(pp_ctx *ctx, const pp_index *ix, size_t gz_len, int32_t first_chunk, int32_t n_chunks,
                  uint32_t flags, pp_job **out);
int pp_job_upload(pp_job *job, const uint8_t *gz);
/*
 * The bytes of the .gz file a job reads — [*file_offset, *file_offset + *length), the union of its
 * chunks' LazyFileReader ranges (LazyFileReader.cs:63-69) — and an upload that takes just those
 * bytes: `range` holds file bytes [range_file_offset, range_file_offset + range_len) and must cover
 * them.  A process that decodes one partition of a large file (one rank per GPU) needs to hold only
 * its own byte range in (pinned) host memory.
 */
int pp_job_file_range(const pp_job *job, int64_t *file_offset, int64_t *length);
int pp_job_upload_range(pp_job *job, const uint8_t *range, int64_t range_file_offset, int64_t range_len);
int pp_job_execute(pp_job *job);
/*
 * pp_job_execute that also delivers every chunk's inflated bytes to HOST memory, concatenated in
 * chunk order (`dst` capacity >= sum of to.Output-from.Output; pinned memory for full speed), while
 * the decode is still running: the kernel raises a per-chunk flag in mapped host memory and the call
 * queues that chunk's device-to-host copy at once, so the PCIe download overlaps the kernels — what
 * a host consumer of FastqRecords (Parsing.cs:41-49 copies every record's bytes) needs.  Returns when
 * all bytes are in `dst`; follow with pp_job_download for counts / line starts.
 */
int pp_job_execute_to_host(pp_job *job, uint8_t *dst, int64_t cap);
int pp_job_download(pp_job *job);
int pp_job_info_get(const pp_job *job, pp_job_info *out);
int pp_job_chunk_info(const pp_job *job, int32_t chunk /* relative to first_chunk */, pp_chunk_info *out);
/* Copy results to host memory.  line_starts: four arrays of total_records u32 each. */
int pp_job_fetch_line_starts(pp_job *job, uint32_t *l0, uint32_t *l1, uint32_t *l2, uint32_t *l3);
/* Inflated bytes of one chunk (dst capacity >= inflated). */
int pp_job_fetch_chunk(pp_job *job, int32_t chunk, uint8_t *dst, int64_t cap);
/* Inflated bytes of all chunks, concatenated (dst capacity >= total_bytes). */
int pp_job_fetch_bytes(pp_job *job, uint8_t *dst, int64_t cap);
/* Device pointers for on-device consumers.  Take them AFTER pp_job_download: the download may
 * re-allocate the line-start arrays when a corpus has more records than first provisioned.
 * Valid until the next pp_job_download / pp_job_free. */
int pp_job_device_ptrs(const pp_job *job, const uint8_t **slots, const uint64_t **chunk_data_off,
                       const uint32_t **l0, const uint32_t **l1, const uint32_t **l2, const uint32_t **l3);
/*
 * On-device consumer (what the reference's driver and benchmark do with the records:
 * records.Count() and counting 'A's, Decompressor/Program.cs:51-52, Benchmark/Naive.cs:158-178):
 * histogram of the bytes of every record's SEQUENCE line, computed on the GPU from the
 * structure-of-arrays line starts.  Only the 256 counters cross PCIe.  Needs pp_job_download first.
 */
int pp_job_base_histogram(pp_job *job, uint64_t counts[256]);
/*
 * On-device consumer for the reference benchmark's pattern search (Benchmark/Naive.cs:167-180:
 * `if (record.Sequence.Contains(pattern)) count++`, ordinal comparison): the number of records
 * whose SEQUENCE line contains `pattern` (any length; an empty pattern matches every record, as
 * string.Contains("") does).  Only the pattern and one counter cross PCIe.  Needs pp_job_download first.
 */
int pp_job_count_pattern(pp_job *job, const uint8_t *pattern, int32_t pattern_len, uint64_t *count);
/*
 * Per-chunk integrity digests computed on the GPU (n_chunks entries each; either pointer may be NULL),
 * so that a whole DecompressAll can be checked against an independent implementation with two
 * integers per chunk crossing PCIe instead of every byte and record.  With
 *   mix(x) = splitmix64 finaliser: z = x + 0x9E3779B97F4A7C15; z = (z ^ z>>30) * 0xBF58476D1CE4E5B9;
 *            z = (z ^ z>>27) * 0x94D049BB133111EB; z ^ z>>31
 * and all arithmetic modulo 2^64:
 *   bytes_digest[k]  = mix(n) + sum_i mix(i) * (W_i + 1) over the n bytes Core.ExtractDeflateIndex
 *                      produced for chunk k (Core.cs:133-192), W_i = little-endian u64 of bytes
 *                      [8i, 8i+8), zero padded;
 *   fields_digest[k] = sum_r sum_{f<9} mix(9r + f) * (field_{r,f} + 1) over the chunk's records r and
 *                      the nine integers Parsing.Parse computes per record (Parsing.cs:20-39: start,
 *                      idnFrom, idnLen, seqFrom, seqLen, plsFrom, plsLen, qltFrom, qltLen).
 * Needs pp_job_download first.
 */
int pp_job_digests(pp_job *job, uint64_t *bytes_digest, uint64_t *fields_digest);
void pp_job_free(pp_job *job);

/* One-call DecompressAll: create + upload + execute + download.  Free with pp_job_free. */
int pp_decompress_all(pp_ctx *ctx, const pp_index *ix, const uint8_t *gz, size_t gz_len, int32_t first_chunk,
                      int32_t n_chunks, uint32_t flags, pp_job **out);

/* ------------------------------------------------ GPU-assisted CreateIndex, first slice */

/*
 * The first bit and the output offset of every deflate block of the gzip member `gz` — the stops
 * Core.BuildDeflateIndex gets from a serial inflate(Z_BLOCK) pass (Decompressor/Core.cs:64) and the only
 * places where Core.cs:98-109 may drop a checkpoint — found on the GPU: the compressed stream is cut into
 * segments of `segment_bytes` (<= 0: one segment per resident CTA, 128 KiB .. 8 MiB), every segment searches for its first block header
 * speculatively and walks its blocks with the inflate kernel's Huffman passes (no history needed),
 * and the segments are stitched on the host (a seam that does not close is re-walked; `passes` counts
 * the kernel launches).  start_bits[i] = 8*Input - Bits of a checkpoint taken at block i (Common/Index.cs
 * Point.Input/Bits), out_offsets[i] = its Point.Output.  *count receives the number of blocks (PP_BUF_ERROR
 * when > cap), *end_bit the bit after the final block, *total_out the stream's length.  The rest of
 * CreateIndex ('@' counting, checkpoint choice, windows) is still host code (pp_index_create).
 */
int pp_scan_blocks(pp_ctx *ctx, const uint8_t *gz, size_t gz_len, int64_t segment_bytes, int64_t *start_bits,
                   int64_t *out_offsets, int64_t cap, int64_t *count, int64_t *end_bit, int64_t *total_out,
                   float *kernel_ms, int32_t *passes);

/* ------------------------------------------------------------- CreateIndex on the GPU */

typedef struct pp_create_stats {  /* where pp_index_create_gpu's time went (CUDA events, ms) */
    float h2d_ms;          /* the file to the device                                        */
    float scan_ms;         /* block scan incl. stitching on the host (scan_kernel_ms: kernels only) */
    float scan_kernel_ms;
    float plan_ms;         /* segment planning, allocations                                  */
    float inflate_ms;      /* the inflate kernel over every segment, twice                   */
    float chain_ms;        /* window behind every segment                                    */
    float resolve_ms;      /* dictionary-derived bytes replaced                              */
    float count_crc_ms;    /* '@' statistics per block, CRC-32                               */
    float gather_ms;       /* host: point selection; device: windows + offsets; D2H          */
    float total_ms;
    int64_t blocks;        /* deflate blocks = Z_BLOCK stops                                 */
    int64_t total_out;     /* inflated length                                                */
    int32_t segments;      /* decode segments                                                */
    int32_t scan_passes;
    int32_t points;
    int32_t pad;
} pp_create_stats;

/*
 * Core.BuildDeflateIndex (Decompressor/Core.cs:14-131) on the GPU: the same index, point for point and
 * byte for byte, as pp_index_create — which stays the general path — without the serial inflate pass.
 * The block scan (pp_scan_blocks) finds the Z_BLOCK stops; runs of blocks are inflated by the inflate
 * kernel twice, each time with a dictionary that encodes its own positions instead of data, which tells
 * for every output byte whether it is final or which byte of the preceding 32 KB it copies; the windows
 * are then chained through the segments and the bytes resolved; '@' statistics per block (Core.cs:86)
 * feed the reference's own checkpoint rule (Core.cs:98-125), and the windows / offsets of the chosen
 * points are gathered on the device.  ISIZE and CRC-32 of the trailer are verified (zlib: -3).
 * gz: one complete gzip member; more members or trailing bytes, or a stream of pathologically
 * small blocks (< 64 compressed bytes on average): PP_E_UNSUPPORTED, use pp_index_create.  PP_E_RECORD_TOO_LONG / PP_INDEX_LIFT_RECORD_CAP as pp_index_create.  A stream that is
 * both damaged and holds an over-long record may report the other of the two errors.  Device memory:
 * about twice the inflated size; PP_MEM_ERROR when the device cannot provide it (pp_index_create then is
 * the way).  stats may be NULL.
 */
int pp_index_create_gpu(pp_ctx *ctx, const uint8_t *gz, size_t gz_len, uint32_t chunksize, uint32_t flags,
                        pp_index **out, pp_create_stats *stats);

/* ------------------------------------------------------ DecompressAll on several GPUs */

typedef struct pp_multi pp_multi; /* one DecompressAll spread over several GPUs */

typedef struct pp_multi_info {
    int32_t n_parts;          /* GPUs used (= entries of `devices`)                           */
    int32_t n_chunks;         /* chunks over all parts                                        */
    int64_t total_records;    /* records over all parts, canonical order                      */
    int64_t total_bytes;      /* inflated bytes over all parts                                */
    int64_t compressed_bytes; /* compressed bytes consumed over all parts                     */
    int32_t status;           /* first non-zero chunk status in part order, else 0            */
    int32_t pad;
} pp_multi_info;

/*
 * The split the multi-GPU DecompressAll uses, also exported for callers that run one process per
 * GPU: chunks 0..Count-2 cut into `parts` contiguous ranges of near-equal COMPRESSED size (Input
 * deltas; LazyFileReader.cs:53-69 defines a chunk's byte range).  Ranges are disjoint, ordered,
 * cover every chunk, and may be empty when there are fewer chunks than parts.
 */
int pp_partition_chunks(const pp_index *ix, int32_t parts, int32_t *first_chunk, int32_t *n_chunks);

/*
 * DecompressAll (BatchedFASTQ.cs:54-98) over the GPUs `devices[0..n_devices)` of this box: the chunk
 * list is partitioned with pp_partition_chunks, every GPU gets its own context and host thread and
 * touches only ITS compressed byte range and ITS checkpoint windows; nothing is exchanged between
 * GPUs (chunks are independent), the global ordinal of a part's first record is a host-side prefix
 * sum.  `flags` are pp_job_create's.  Returns 0, a chunk's negative ZResult (the handle is still
 * returned: per-part jobs tell which chunk), or an API error (no handle).
 */
int pp_decompress_all_multi(const int32_t *devices, int32_t n_devices, const pp_index *ix, const uint8_t *gz,
                            size_t gz_len, uint32_t flags, pp_multi **out);
/* pp_decompress_all_multi and pp_pair_decompress_all borrow their per-GPU contexts (stream, token scratch)
 * from a process-wide cache and return them when the handle is freed; this closes the idle ones. */
void pp_release_cached_contexts(void);
int pp_multi_info_get(const pp_multi *m, pp_multi_info *out);
/* Part `part`: its job (owned by the handle: do not free), its device, the global ordinal of its first record. */
int pp_multi_part(const pp_multi *m, int32_t part, pp_job **job, int32_t *device, int64_t *record_base);
void pp_multi_free(pp_multi *m);

/* --------------------------------------------------------------- paired-end R1 / R2 */

typedef struct pp_pair pp_pair;

typedef struct pp_pair_info {
    int32_t n_parts;        /* GPUs used                                                          */
    int32_t topup_chunks;   /* R2 chunks decoded a second time so that every mate is co-resident   */
    int64_t records_r1, records_r2;
    int64_t pairs;          /* min(records_r1, records_r2): record r of R1 and record r of R2      */
    int32_t status;         /* first non-zero chunk status, else 0                                 */
    int32_t pad;
} pp_pair_info;

/*
 * Paired DecompressAll (README.md:9: R1/R2 "chunks with identical record counts"; the reference has
 * no code for it).  Checkpoints sit on deflate block ends, so the two files cannot cut their chunks at
 * the same records; what is identical is a record's ORDINAL once the H1 duplicates are dropped
 * (PP_JOB_STRICT is forced).  Both chunk lists are partitioned over `devices`, each GPU decodes its part
 * of R1 and of R2 concurrently, and then the few R2 chunks that hold mates of its R1 records but belong
 * to a neighbour's part are decoded there too, so every R1 record of a part has its mate on the same
 * GPU.  Per part: one R1 job and 1-3 R2 jobs in ordinal order, each with the global ordinal of its first
 * record; pp_pair_locate maps an ordinal to (R2 job, record index) — the record-range map that stands in
 * for "identical chunk record counts".  Handles are owned by the pp_pair.
 */
int pp_pair_decompress_all(const int32_t *devices, int32_t n_devices, const pp_index *ix1, const uint8_t *gz1,
                           size_t gz1_len, const pp_index *ix2, const uint8_t *gz2, size_t gz2_len, uint32_t flags,
                           pp_pair **out);
int pp_pair_info_get(const pp_pair *p, pp_pair_info *out);
int pp_pair_part(const pp_pair *p, int32_t part, pp_job **r1, int64_t *r1_base, int32_t *n_r2_jobs);
int pp_pair_part_r2(const pp_pair *p, int32_t part, int32_t which, pp_job **r2, int64_t *r2_base);
int pp_pair_locate(const pp_pair *p, int32_t part, int64_t ordinal, int32_t *which_r2, int64_t *record_index);
void pp_pair_free(pp_pair *p);

#ifdef __cplusplus
}
#endif
#endif /* PPB200_H */

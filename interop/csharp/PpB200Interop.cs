// Source-only binding of libppb200.so for the reference's Interop project (not compiled in this
// repository: the build image has no .NET).  Mirrors Interop/PlatformInterop.cs:6-35; the C ABI is
// include/ppb200.h.  Drop this file next to PlatformInterop.cs.
using System.Runtime.InteropServices;
using System.Security;

namespace ParallelParsing.Interop;

[SuppressUnmanagedCodeSecurity]
internal static class LibPpB200
{
    const string L = "ppb200";   // libppb200.so next to the executable / on LD_LIBRARY_PATH

    [DllImport(L)] public static extern int pp_abi_version();
    [DllImport(L)] public static extern IntPtr pp_strerror(int code);

    // Index  (Common/Index.cs, Common/IndexIO.cs, Core.BuildDeflateIndex)
    [DllImport(L)] public static unsafe extern int pp_index_create(byte* gz, nuint gzLen, uint chunksize, uint flags, out IntPtr index);
    [DllImport(L, CharSet = CharSet.Ansi)] public static extern int pp_index_create_file(string gzPath, uint chunksize, uint flags, out IntPtr index);
    [DllImport(L)] public static extern int pp_index_new(out IntPtr index);
    [DllImport(L)] public static unsafe extern int pp_index_add(IntPtr index, int bits, long input, long output, byte* window, byte* offset, int offsetLen);
    [DllImport(L)] public static unsafe extern int pp_index_add_point(IntPtr index, int bits, long input, long output, uint left, byte* window, byte* offset, int offsetLen);
    [DllImport(L, CharSet = CharSet.Ansi)] public static extern int pp_index_serialize(IntPtr index, string path);
    [DllImport(L, CharSet = CharSet.Ansi)] public static extern int pp_index_deserialize(string path, out IntPtr index);
    [DllImport(L, CharSet = CharSet.Ansi)] public static extern int pp_index_serialize_v1(IntPtr index, string path);   // extension: file version 1, compressed windows
    [DllImport(L)] public static extern int pp_index_count(IntPtr index);
    [DllImport(L)] public static extern int pp_index_chunk_max_bytes(IntPtr index);
    [DllImport(L)] public static extern int pp_index_point(IntPtr index, int i, out pp_point p);
    [DllImport(L)] public static extern void pp_index_free(IntPtr index);

    // Device
    [DllImport(L)] public static extern int pp_open(int device, out IntPtr ctx);
    [DllImport(L)] public static extern void pp_close(IntPtr ctx);
    [DllImport(L)] public static extern int pp_host_alloc(nuint bytes, out IntPtr p);
    [DllImport(L)] public static extern void pp_host_free(IntPtr p);
    [DllImport(L)] public static unsafe extern int pp_host_register(void* p, nuint bytes);
    [DllImport(L)] public static unsafe extern void pp_host_unregister(void* p);

    // Decompress(checkpoint) and Parse  (Core.ExtractDeflateIndex, Parsing.Parse)
    [DllImport(L)] public static unsafe extern long pp_extract(IntPtr ctx, byte* fileBuffer, long fileBufferLen, IntPtr index, int fromPoint, byte* buf, long bufLen);
    [DllImport(L)] public static unsafe extern long pp_parse(IntPtr ctx, byte* prepend, long prependLen, byte* rest, long restLen, uint* lineStarts, long cap, out uint parseEnd);

    // DecompressAll  (BatchedFASTQ)
    [DllImport(L)] public static extern int pp_job_create(IntPtr ctx, IntPtr index, nuint gzLen, int firstChunk, int nChunks, uint flags, out IntPtr job);
    [DllImport(L)] public static unsafe extern int pp_job_upload(IntPtr job, byte* gz);
    [DllImport(L)] public static extern int pp_job_execute(IntPtr job);
    [DllImport(L)] public static extern int pp_job_download(IntPtr job);
    [DllImport(L)] public static extern int pp_job_info_get(IntPtr job, out pp_job_info info);
    [DllImport(L)] public static extern int pp_job_chunk_info(IntPtr job, int chunk, out pp_chunk_info info);
    [DllImport(L)] public static unsafe extern int pp_job_fetch_line_starts(IntPtr job, uint* l0, uint* l1, uint* l2, uint* l3);
    [DllImport(L)] public static unsafe extern int pp_job_fetch_chunk(IntPtr job, int chunk, byte* dst, long cap);
    [DllImport(L)] public static unsafe extern int pp_job_fetch_bytes(IntPtr job, byte* dst, long cap);
    [DllImport(L)] public static extern int pp_job_device_ptrs(IntPtr job, out IntPtr slots, out IntPtr chunkDataOff, out IntPtr l0, out IntPtr l1, out IntPtr l2, out IntPtr l3);
    [DllImport(L)] public static unsafe extern int pp_job_base_histogram(IntPtr job, ulong* counts256);
    [DllImport(L)] public static unsafe extern int pp_job_count_pattern(IntPtr job, byte* pattern, int patternLen, out ulong count);
    [DllImport(L)] public static extern void pp_job_free(IntPtr job);
    [DllImport(L)] public static unsafe extern int pp_decompress_all(IntPtr ctx, IntPtr index, byte* gz, nuint gzLen, int firstChunk, int nChunks, uint flags, out IntPtr job);

    // round 2: partial-file uploads, streamed download, digests
    [DllImport(L)] public static extern int pp_job_file_range(IntPtr job, out long fileOffset, out long length);
    [DllImport(L)] public static unsafe extern int pp_job_upload_range(IntPtr job, byte* range, long rangeFileOffset, long rangeLen);
    [DllImport(L)] public static unsafe extern int pp_job_execute_to_host(IntPtr job, byte* dst, long cap);
    [DllImport(L)] public static unsafe extern int pp_job_digests(IntPtr job, ulong* bytesDigest, ulong* fieldsDigest);

    // DecompressAll over several GPUs (one call; partition below the ABI)
    [DllImport(L)] public static unsafe extern int pp_partition_chunks(IntPtr index, int parts, int* firstChunk, int* nChunks);
    [DllImport(L)] public static unsafe extern int pp_decompress_all_multi(int* devices, int nDevices, IntPtr index, byte* gz, nuint gzLen, uint flags, out IntPtr multi);
    [DllImport(L)] public static extern int pp_multi_info_get(IntPtr multi, out pp_multi_info info);
    [DllImport(L)] public static extern int pp_multi_part(IntPtr multi, int part, out IntPtr job, out int device, out long recordBase);
    [DllImport(L)] public static extern void pp_multi_free(IntPtr multi);
    [DllImport(L)] public static extern void pp_release_cached_contexts();

    // paired-end R1/R2 (README.md:9)
    [DllImport(L)] public static unsafe extern int pp_pair_decompress_all(int* devices, int nDevices, IntPtr index1, byte* gz1, nuint gz1Len, IntPtr index2, byte* gz2, nuint gz2Len, uint flags, out IntPtr pair);
    [DllImport(L)] public static extern int pp_pair_info_get(IntPtr pair, out pp_pair_info info);
    [DllImport(L)] public static extern int pp_pair_part(IntPtr pair, int part, out IntPtr r1Job, out long r1Base, out int nR2Jobs);
    [DllImport(L)] public static extern int pp_pair_part_r2(IntPtr pair, int part, int which, out IntPtr r2Job, out long r2Base);
    [DllImport(L)] public static extern int pp_pair_locate(IntPtr pair, int part, long ordinal, out int whichR2, out long recordIndex);
    [DllImport(L)] public static extern void pp_pair_free(IntPtr pair);

    // GPU-assisted CreateIndex, first slice: every deflate block's first bit and output offset
    [StructLayout(LayoutKind.Sequential)] public struct PpCreateStats { public float H2dMs, ScanMs, ScanKernelMs, PlanMs, InflateMs, ChainMs, ResolveMs, CountCrcMs, GatherMs, TotalMs; public long Blocks, TotalOut; public int Segments, ScanPasses, Points, Pad; }
    [DllImport(L)] public static unsafe extern int pp_index_create_gpu(IntPtr ctx, byte* gz, nuint gzLen, uint chunksize, uint flags, out IntPtr index, out PpCreateStats stats);
    [DllImport(L)] public static unsafe extern int pp_scan_blocks(IntPtr ctx, byte* gz, nuint gzLen, long segmentBytes, long* startBits, long* outOffsets, long cap, out long count, out long endBit, out long totalOut, out float kernelMs, out int passes);
}

internal static class PpJobFlags   // include/ppb200.h
{
    public const uint Strict = 1, ZeroCopy = 2, Pipeline = 4, CompactWindows = 8;
}

[StructLayout(LayoutKind.Sequential)]
internal unsafe struct pp_point        // include/ppb200.h pp_point  <->  Common/Index.cs Point
{
    public long output, input; public int bits, offset_len; public byte* window, offset;
}
[StructLayout(LayoutKind.Sequential)]
internal struct pp_job_info
{
    public int first_chunk, n_chunks; public long total_records, total_bytes, scanned_bytes, compressed_bytes, h2d_bytes, d2h_bytes;
    public int status, exact_chunks; public float upload_ms, inflate_ms, scan_ms, parse_ms, download_ms; public int launches;
}
[StructLayout(LayoutKind.Sequential)]
internal struct pp_chunk_info
{
    public int status, prefix_len; public long inflated, records, record_base; public uint parse_end, flags;
}
[StructLayout(LayoutKind.Sequential)]
internal struct pp_multi_info
{
    public int n_parts, n_chunks; public long total_records, total_bytes, compressed_bytes; public int status, pad;
}
[StructLayout(LayoutKind.Sequential)]
internal struct pp_pair_info
{
    public int n_parts, topup_chunks; public long records_r1, records_r2, pairs; public int status, pad;
}

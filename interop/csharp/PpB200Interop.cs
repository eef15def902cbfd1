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

// Source-only: a GPU-backed stand-in for Decompressor/BatchedFASTQ.cs:10-101 on top of
// PpB200Interop.cs.  Not compiled in this repository (no .NET in the build image); it shows how a
// maintainer wires libppb200.so into the reference.  Same constructor arguments and the same
// IEnumerable<FastqRecord> contract; records come out in canonical order (chunk ascending, file
// order inside a chunk) instead of the reference's nondeterministic cross-chunk order.
using System.Buffers;
using System.Collections;
using ParallelParsing.Common;
using ParallelParsing.Interop;
using Index = ParallelParsing.Common.Index;

namespace ParallelParsing;

public sealed unsafe class GpuBatchedFASTQ : IEnumerable<FastqRecord>, IDisposable
{
    private readonly IntPtr _ctx, _index, _job;
    private readonly IntPtr _gz;          // pinned copy of the .gz file (pp_host_alloc)
    private readonly Index _managedIndex;  // for Point.offset (the prefix of every chunk's combined memory)
    private readonly pp_job_info _info;

    public GpuBatchedFASTQ(string indexPath, string gzipPath, bool enableSsdOptimization, int device = 0)
        : this(IndexIO.Deserialize(indexPath), indexPath, gzipPath, device) { }

    private GpuBatchedFASTQ(Index index, string indexPath, string gzipPath, int device)
    {
        _managedIndex = index;
        Check(LibPpB200.pp_open(device, out _ctx));
        Check(LibPpB200.pp_index_deserialize(indexPath, out _index));   // same bytes IndexIO wrote
        var len = new FileInfo(gzipPath).Length;
        Check(LibPpB200.pp_host_alloc((nuint)len, out _gz));
        using (var fs = File.OpenRead(gzipPath))
            fs.ReadExactly(new Span<byte>((void*)_gz, checked((int)len)));   // > 2 GiB: read in slices
        // upload + inflate kernel + parse kernel + per-chunk results, all chunks of the index
        Check(LibPpB200.pp_decompress_all(_ctx, _index, (byte*)_gz, (nuint)len, 0, -1, 0, out _job));
        Check(LibPpB200.pp_job_info_get(_job, out _info));
    }

    /// <summary>Core.BuildDeflateIndex (Decompressor/Core.cs:14-131) on the GPU: writes the same index file
    /// IndexIO.Serialize(Core.BuildDeflateIndex(file, chunksize)) writes, ~160x faster (pp_index_create_gpu);
    /// input the GPU entry point declines (-106: more than one gzip member; -4: not enough device memory)
    /// goes through the host pass, which is the reference's own zlib loop.</summary>
    public static void CreateIndexFile(string gzipPath, string indexPath, uint chunksize, int device = 0)
    {
        Check(LibPpB200.pp_open(device, out var ctx));
        var len = new FileInfo(gzipPath).Length;
        Check(LibPpB200.pp_host_alloc((nuint)len, out var gz));
        try
        {
            using (var fs = File.OpenRead(gzipPath))
                fs.ReadExactly(new Span<byte>((void*)gz, checked((int)len)));
            int rc = LibPpB200.pp_index_create_gpu(ctx, (byte*)gz, (nuint)len, chunksize, 0, out var ix, out _);
            if (rc == -106 || rc == -4) rc = LibPpB200.pp_index_create((byte*)gz, (nuint)len, chunksize, 0, out ix);
            Check(rc);                                   // ZException where Core.cs:33,74 throws; -104 = Core.cs:93
            Check(LibPpB200.pp_index_serialize(ix, indexPath));
            LibPpB200.pp_index_free(ix);
        }
        finally
        {
            LibPpB200.pp_host_free(gz);
            LibPpB200.pp_close(ctx);
        }
    }

    public long Count => _info.total_records;   // what Benchmark/Naive.cs:158-162 measures

    public IEnumerator<FastqRecord> GetEnumerator()
    {
        var n = checked((int)_info.total_records);
        uint[] l0 = new uint[n], l1 = new uint[n], l2 = new uint[n], l3 = new uint[n];
        fixed (uint* p0 = l0, p1 = l1, p2 = l2, p3 = l3)
            Check(LibPpB200.pp_job_fetch_line_starts(_job, p0, p1, p2, p3));
        for (int k = 0; k < _info.n_chunks; k++)
        {
            Check(LibPpB200.pp_job_chunk_info(_job, k, out var c));
            if (c.records == 0) continue;
            // combined memory of the chunk = Point.offset ++ inflated bytes (Parsing.cs:72-117)
            var prefix = _managedIndex[_info.first_chunk + k].offset ?? Array.Empty<byte>();
            using var chunkOwner = MemoryPool<byte>.Shared.Rent(prefix.Length + (int)c.inflated);
            prefix.CopyTo(chunkOwner.Memory);
            using (var h = chunkOwner.Memory.Slice(prefix.Length).Pin())
                Check(LibPpB200.pp_job_fetch_chunk(_job, k, (byte*)h.Pointer, c.inflated));
            for (long r = c.record_base; r < c.record_base + c.records; r++)
            {
                // Parsing.cs:20-39 in terms of the four line starts
                int next = r + 1 < c.record_base + c.records ? (int)l0[r + 1] : (int)c.parse_end;
                int start = (int)l0[r] + 1, end = next;
                // one rented buffer PER RECORD, as Parsing.cs:41-43 does: FastqRecord.Dispose() disposes its
                // Owner (Common/FastqRecord.cs:80-83) and the reference's consumers dispose every record, so
                // records must not share an owner
                var owner = MemoryPool<byte>.Shared.Rent(end - start);
                var m = owner.Memory;
                chunkOwner.Memory.Slice(start, end - start).CopyTo(m);
                yield return new FastqRecord(owner,
                    m.Slice(0, (int)(l1[r] - l0[r]) - 2),                                  // Identifier (no '@', no '\n')
                    m.Slice((int)l1[r] - start, (int)(l2[r] - l1[r]) - 1),                  // Sequence
                    m.Slice((int)l2[r] + 1 - start, (int)(l3[r] - l2[r]) - 2),              // Other (no '+')
                    m.Slice((int)l3[r] - start, next - (int)l3[r] - 1));                    // Quality
            }
        }
    }
    IEnumerator IEnumerable.GetEnumerator() => GetEnumerator();

    public void Dispose()
    {
        LibPpB200.pp_job_free(_job);
        LibPpB200.pp_host_free(_gz);
        LibPpB200.pp_index_free(_index);
        LibPpB200.pp_close(_ctx);
    }

    private static void Check(long rc) { if (rc < 0) throw new ZException((ZResult)(int)rc); }
}

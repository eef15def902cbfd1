/*
 * pp_oracle.c — CPU ORACLE for the checkpointed gzip-FASTQ decode path of
 * Quantumzhao/ParallelParsing.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing in the product (parallelparsing_b200/, include/) may link, import or
 * call this file.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker / the
 * timed host baseline.
 *
 * What it is: a statement-by-statement C restatement of the reference's C#
 * hot path on top of the SAME third-party library the reference P/Invokes
 * (system zlib; the reference binds it with [DllImport("libz")],
 * Interop/PlatformInterop.cs:9-34, and passes version "1.2.11",
 * Common/Constants.cs:6; this image ships zlib 1.3 whose inflate output and
 * Z_BLOCK/data_type semantics are identical).
 *
 * PARITY UNPINNED: the reference holds no tests, golden vectors or fixtures
 *  and its C# cannot be built or run in this image (no
 * dotnet/mono).  The oracle is therefore pinned only by (i) calling the same
 * zlib entry points in the same order, (ii) concat(chunks) == zlib stream
 * inflate, (iii) Serialize/Deserialize byte round trips — see tests/.
 *
 * Each function cites the reference file:line it follows.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <pthread.h>
#include <zlib.h>

#define WINSIZE 32768 /* Common/Constants.cs:9  */
#define CHUNK 16384   /* Common/Constants.cs:12 */

/* ---------------------------------------------------------------- Index -- */

/* Common/Index.cs:51-82 */
typedef struct {
    int64_t output;
    int64_t input;
    int32_t bits;
    uint8_t window[WINSIZE];
    uint8_t *offset; /* may be NULL */
    int32_t offset_len;
} ora_point;

/* Common/Index.cs:5-49 */
typedef struct {
    ora_point **pts;
    int32_t count, cap;
    int32_t chunk_max_bytes;
} ora_index;

ora_index *ora_index_new(void) { return (ora_index *)calloc(1, sizeof(ora_index)); }

void ora_index_free(ora_index *ix)
{
    if (!ix) return;
    for (int i = 0; i < ix->count; i++) {
        free(ix->pts[i]->offset);
        free(ix->pts[i]);
    }
    free(ix->pts);
    free(ix);
}

static void index_add(ora_index *ix, ora_point *p)
{
    if (ix->count == ix->cap) {
        ix->cap = ix->cap ? ix->cap * 2 : 8;
        ix->pts = (ora_point **)realloc(ix->pts, sizeof(*ix->pts) * (size_t)ix->cap);
    }
    ix->pts[ix->count++] = p;
}

/* Common/Index.cs:24-48  Index.AddPoint */
static void index_add_point(ora_index *ix, int bits, int64_t input, int64_t output, uint32_t left,
                            const uint8_t *window, const uint8_t *offset, int32_t offset_len)
{
    if (ix->count == 0) {
        ix->chunk_max_bytes = (int32_t)output; /* :27-30 */
    } else {
        int32_t outputSize = (int32_t)output - (int32_t)ix->pts[ix->count - 1]->output; /* :33 */
        if (outputSize > ix->chunk_max_bytes) ix->chunk_max_bytes = outputSize;
    }
    ora_point *next = (ora_point *)calloc(1, sizeof(ora_point)); /* :39 (Window zeroed) */
    next->output = output;
    next->input = input;
    next->bits = bits;
    next->offset_len = offset_len; /* :40 */
    next->offset = (uint8_t *)malloc(offset_len > 0 ? (size_t)offset_len : 1);
    if (offset_len > 0) memcpy(next->offset, offset, (size_t)offset_len);
    if (left != 0) memcpy(next->window, window + WINSIZE - left, left); /* :42-43 */
    if (left < WINSIZE) memcpy(next->window + left, window, WINSIZE - left); /* :45-46 */
    index_add(ix, next);
}

int32_t ora_index_count(const ora_index *ix) { return ix->count; }
int32_t ora_index_chunk_max_bytes(const ora_index *ix) { return ix->chunk_max_bytes; }
int64_t ora_point_output(const ora_index *ix, int i) { return ix->pts[i]->output; }
int64_t ora_point_input(const ora_index *ix, int i) { return ix->pts[i]->input; }
int32_t ora_point_bits(const ora_index *ix, int i) { return ix->pts[i]->bits; }
const uint8_t *ora_point_window(const ora_index *ix, int i) { return ix->pts[i]->window; }
const uint8_t *ora_point_offset(const ora_index *ix, int i) { return ix->pts[i]->offset; }
int32_t ora_point_offset_len(const ora_index *ix, int i) { return ix->pts[i]->offset_len; }

/* ------------------------------------------------------- BuildDeflateIndex */

/*
 * Decompressor/Core.cs:14-131  Core.BuildDeflateIndex(FileStream, uint chunksize).
 * `gz`/`gz_len` stand in for the FileStream (file.Read of CHUNK bytes, :41).
 * lift_cap != 0 is a DOCUMENTED EXTENSION: the reference's
 * offsetBeforePoint is WINSIZE bytes (:23) and :93 throws IndexOutOfRange for a
 * record > 32768 B; with lift_cap the buffer grows instead.
 * Returns 0 or a negative ZResult (Interop/Conventions.cs:9-20); -100 stands
 * for the reference's IndexOutOfRangeException.
 */
int ora_build_index(const uint8_t *gz, size_t gz_len, uint32_t chunksize, int lift_cap, ora_index **out)
{
    z_stream strm;
    memset(&strm, 0, sizeof strm);
    ora_index *index = ora_index_new();
    uint8_t *input = (uint8_t *)malloc(CHUNK);
    uint8_t *window = (uint8_t *)calloc(1, WINSIZE);
    int recordCounter = 0;
    int prevAvailOut = 0;
    size_t obp_cap = WINSIZE;
    uint8_t *offsetBeforePoint = (uint8_t *)calloc(1, obp_cap);
    int offsetArraySize = 0;
    int ret;
    int64_t totin, totout;
    size_t file_pos = 0;
    int window_set = 0;
    int rc = 0;

    ret = inflateInit2(&strm, 47); /* :30 */
    if (ret != Z_OK) { rc = ret; goto done_noend; }

    totin = totout = 0;
    strm.avail_out = 0;
    do {
        size_t n = gz_len - file_pos < CHUNK ? gz_len - file_pos : CHUNK; /* :41 */
        memcpy(input, gz + file_pos, n);
        file_pos += n;
        strm.avail_in = (uInt)n;
        if (strm.avail_in == 0) { rc = Z_DATA_ERROR; goto done; } /* :42-45 */
        strm.next_in = input;

        do {
            if (strm.avail_out == 0) { /* :52-56 */
                strm.avail_out = WINSIZE;
                strm.next_out = window;
                window_set = 1;
            }
            totin += strm.avail_in;
            totout += strm.avail_out;
            ret = inflate(&strm, Z_BLOCK); /* :64 */
            totin -= strm.avail_in;
            totout -= strm.avail_out;
            if (ret == Z_NEED_DICT || ret == Z_MEM_ERROR || ret == Z_DATA_ERROR || ret == Z_STREAM_ERROR ||
                ret == Z_BUF_ERROR || ret == Z_VERSION_ERROR) { /* :68-74 */
                rc = ret;
                goto done;
            }

            if (window_set) { /* :76 */
                int currNextOutLength = WINSIZE; /* :79 */
                int iStartPos = prevAvailOut == 0 ? 0 : currNextOutLength - prevAvailOut; /* :80 */
                for (int i = iStartPos; i < currNextOutLength - (int)strm.avail_out; i++) { /* :82 */
                    uint8_t c = window[i];
                    if (c == 64) { /* :86-91 */
                        recordCounter++;
                        /* Array.Clear(offsetBeforePoint): contents beyond offsetArraySize are
                         * never observed (only [0..offsetArraySize) is sliced, :107), so the
                         * 32 KB memset is elided here without changing any result. */
                        offsetArraySize = 0;
                    }
                    if ((size_t)offsetArraySize >= obp_cap) { /* :93 IndexOutOfRangeException */
                        if (!lift_cap) { rc = -100; goto done; }
                        obp_cap *= 2;
                        offsetBeforePoint = (uint8_t *)realloc(offsetBeforePoint, obp_cap);
                    }
                    offsetBeforePoint[offsetArraySize] = c; /* :93 */
                    offsetArraySize++;                      /* :94 */
                }
                prevAvailOut = strm.avail_out > 0 ? (int)strm.avail_out : 0; /* :96 */

                if ((strm.data_type & 128) != 0 && (strm.data_type & 64) == 0) { /* :98 */
                    if (totout == 0) /* :101-102 */
                        index_add_point(index, strm.data_type & 7, totin, totout, strm.avail_out, window, NULL, 0);
                    else {
                        /* :105  int > uint compares as long in C# */
                        if ((int64_t)recordCounter > (int64_t)(uint32_t)(chunksize - 8u)) {
                            index_add_point(index, strm.data_type & 7, totin, totout, strm.avail_out, window,
                                            offsetBeforePoint, offsetArraySize); /* :107 */
                            recordCounter = 0;
                        }
                    }
                }
            }

            if (ret == Z_STREAM_END) { /* :114-125 */
                if (strm.avail_in != 0 || file_pos != gz_len) {
                    ret = inflateReset(&strm);
                    if (ret != Z_OK) { rc = ret; goto done; }
                    continue;
                }
                index_add_point(index, strm.data_type & 7, totin, totout, strm.avail_out, window, NULL, 0);
                break;
            }
        } while (strm.avail_in != 0); /* :127 */
    } while (ret != Z_STREAM_END); /* :128 */

done:
    inflateEnd(&strm);
done_noend:
    free(input);
    free(window);
    free(offsetBeforePoint);
    if (rc != 0) {
        ora_index_free(index);
        index = NULL;
    }
    *out = index;
    return rc;
}

/* ------------------------------------------------------------- IndexIO ---- */

/* Common/IndexIO.cs:7-27  Serialize (BinaryWriter = little endian) */
int ora_index_serialize(const ora_index *ix, const char *path)
{
    FILE *f = fopen(path, "wb");
    if (!f) return -1;
    int32_t zero = 0, winlen = WINSIZE;
    fwrite(&zero, 4, 1, f);                /* :12 */
    fwrite(&ix->chunk_max_bytes, 4, 1, f); /* :13 */
    fwrite(&ix->count, 4, 1, f);           /* :15 */
    for (int i = 0; i < ix->count; i++) {  /* :17-26 */
        const ora_point *p = ix->pts[i];
        fwrite(&p->output, 8, 1, f);
        fwrite(&p->input, 8, 1, f);
        fwrite(&p->bits, 4, 1, f);
        fwrite(&winlen, 4, 1, f);
        fwrite(p->window, 1, WINSIZE, f);
        int32_t ol = p->offset ? p->offset_len : 0;
        fwrite(&ol, 4, 1, f);
        if (ol) fwrite(p->offset, 1, (size_t)ol, f);
    }
    fclose(f);
    return 0;
}

/* Common/IndexIO.cs:29-53  Deserialize.  ChunkMaxBytes is read and discarded
 * (:35,52) so the returned index has chunk_max_bytes == 0. */
int ora_index_deserialize(const char *path, ora_index **out)
{
    FILE *f = fopen(path, "rb");
    *out = NULL;
    if (!f) return -1;
    int32_t hdr[3];
    if (fread(hdr, 4, 3, f) != 3) { fclose(f); return -1; }
    ora_index *ix = ora_index_new();
    int32_t count = hdr[2];
    for (int i = 0; i < count; i++) {
        ora_point *p = (ora_point *)calloc(1, sizeof(ora_point));
        int32_t winlen = 0, ol = 0;
        int ok = fread(&p->output, 8, 1, f) == 1 && fread(&p->input, 8, 1, f) == 1 &&
                 fread(&p->bits, 4, 1, f) == 1 && fread(&winlen, 4, 1, f) == 1;
        if (!ok || winlen != WINSIZE || fread(p->window, 1, WINSIZE, f) != WINSIZE || fread(&ol, 4, 1, f) != 1 ||
            ol < 0) {
            free(p); fclose(f); ora_index_free(ix); return -1;
        }
        p->offset_len = ol;
        p->offset = (uint8_t *)malloc(ol > 0 ? (size_t)ol : 1);
        if (ol > 0 && fread(p->offset, 1, (size_t)ol, f) != (size_t)ol) {
            free(p->offset); free(p); fclose(f); ora_index_free(ix); return -1;
        }
        index_add(ix, p);
    }
    fclose(f);
    *out = ix;
    return 0;
}

/* ---------------------------------------------------- ExtractDeflateIndex -- */

/*
 * Decompressor/Core.cs:133-192  Core.ExtractDeflateIndex(fileBuffer, from, to, buf).
 * fileBuffer is what LazyFileReader hands over: file bytes starting at
 * from.Input-1 of length to.Input-from.Input+1 (Decompressor/LazyFileReader.cs:63-69).
 * Returns bytes produced (>=0) or a negative ZResult where the reference throws.
 */
int64_t ora_extract(const uint8_t *fileBuffer, int64_t fileBufferLen, const ora_index *ix, int from_i, int to_i,
                    uint8_t *buf)
{
    const ora_point *from = ix->pts[from_i], *to = ix->pts[to_i];
    z_stream strm;
    memset(&strm, 0, sizeof strm);
    int len = (int)(to->output - from->output); /* :140 */
    int ret, value = 0;
    if (len < 0) return 0; /* :145 */

    ret = inflateInit2(&strm, -15); /* :148 */
    if (ret != Z_OK) return ret;

    int64_t posInFile = from->bits == 0 ? 1 : 0; /* :151 */
    if (from->bits != 0) {                       /* :152-157 */
        value = fileBuffer[0];
        inflatePrime(&strm, from->bits, value >> (8 - from->bits));
        posInFile++;
    }
    inflateSetDictionary(&strm, from->window, WINSIZE); /* :158 */

    strm.avail_in = 0;
    strm.avail_out = (uInt)len;
    strm.next_out = buf;
    do {
        if (strm.avail_in == 0) { /* :166-176 */
            int64_t rem = fileBufferLen - posInFile;
            value = (int)(rem < CHUNK ? rem : CHUNK);
            strm.next_in = (Bytef *)(fileBuffer + posInFile);
            strm.avail_in = (uInt)value;
            posInFile += value;
            if (value == 0) { inflateEnd(&strm); return Z_DATA_ERROR; }
        }
        ret = inflate(&strm, Z_NO_FLUSH); /* :177 */
        if (ret == Z_MEM_ERROR || ret == Z_DATA_ERROR || ret == Z_NEED_DICT) { /* :178 */
            inflateEnd(&strm);
            return ret;
        }
        if (ret == Z_STREAM_ERROR) break; /* :180-184 */
        if (ret == Z_STREAM_END) break;   /* :185 */
    } while (strm.avail_out != 0);        /* :187 */

    int64_t produced = len - (int64_t)strm.avail_out; /* :191 */
    inflateEnd(&strm);
    return produced;
}

/* ------------------------------------------------------------ Parsing ------ */

/* System.Buffers ArrayPool<byte>.Shared.Rent(n) array length as used through
 * MemoryPool<byte>.Shared.Rent (Decompressor/BatchedFASTQ.cs:65-66): next power
 * of two >= n with a minimum bucket of 16; 0 -> empty; > 2^30 -> exact. */
int64_t ora_rent_size(int64_t n)
{
    if (n <= 0) return 0;
    if (n > (1LL << 30)) return n;
    int64_t s = 16;
    while (s < n) s <<= 1;
    return s;
}

/* Decompressor/Parsing.cs:72-94  CombinedMemory indexer */
typedef struct {
    const uint8_t *prepend;
    int64_t lengthP;
    const uint8_t *rest;
    int64_t restLen;
    int64_t length;
} combined;

static inline int cm_get(const combined *m, int64_t i, int *oob)
{
    if (i < m->lengthP) return m->prepend[i];
    if (i - m->lengthP >= m->restLen) { *oob = 1; return 0; } /* Span indexer would throw */
    return m->rest[i - m->lengthP];
}

/* Decompressor/Parsing.cs:54-69  ParseLine */
static int64_t parse_line(int64_t *pos, const combined *raw, int *oob)
{
    int64_t start = *pos;
    for (;;) { /* :57-62 */
        int b = cm_get(raw, *pos, oob);
        if (*oob) return -1;
        if (b == '\n' || b == 0) break;
        else (*pos)++;
    }
    if (cm_get(raw, *pos, oob) == 0) return -1; /* :64 */
    (*pos)++;                                    /* :67 */
    return *pos - start;
}

/*
 * Decompressor/Parsing.cs:11-51  Parsing.Parse(CombinedMemory raw).
 * `rest`/`rest_len` is the WHOLE rented array (pow-2 sized, zero tail) exactly as
 * BatchedFASTQ.cs:65-68 passes it.  For every record emits 9 int64 into `recs`
 * (capacity `cap` records): start, idnFrom, idnLen, seqFrom, seqLen, plsFrom,
 * plsLen, qltFrom, qltLen — all indices into the combined memory, as in :20-39.
 * Returns the record count (counting continues past `cap`).  An index past the
 * end (the reference would throw) ends parsing and drops the
 * partial record.  If `digest` is non-NULL it receives an FNV-1a hash over the
 * bytes raw[start,end) of every record (the bytes :41-43 copies).
 */
int64_t ora_parse(const uint8_t *prepend, int64_t prepend_len, const uint8_t *rest, int64_t rest_len, int64_t *recs,
                  int64_t cap, uint64_t *digest)
{
    combined raw = {prepend, prepend ? prepend_len : 0, rest, rest_len, 0};
    raw.length = raw.lengthP + raw.restLen; /* :84 */
    int64_t n = 0;
    uint64_t h = 1469598103934665603ULL;
    int oob = 0;
    for (int64_t i = 0; i < raw.length;) { /* :13 */
        if (cm_get(&raw, i, &oob) == '\0') break; /* :16 */
        i++;                                       /* :19 skip @ (unchecked) */
        int64_t start = i;
        int64_t idnFrom = i;
        int64_t idnLen = parse_line(&i, &raw, &oob) - 1; /* :23 */
        if (idnLen < 0) break;
        int64_t seqFrom = i;
        int64_t seqLen = parse_line(&i, &raw, &oob) - 1; /* :27 */
        if (seqLen < 0) break;
        i++; /* :30 skip + (unchecked) */
        int64_t plsFrom = i;
        int64_t plsLen = parse_line(&i, &raw, &oob) - 1; /* :33 */
        if (plsLen < 0) break;
        int64_t qltFrom = i;
        int64_t qltLen = parse_line(&i, &raw, &oob) - 1; /* :37 */
        if (qltLen < 0) break;
        int64_t end = i;
        if (n < cap && recs) {
            int64_t *r = recs + 9 * n;
            r[0] = start; r[1] = idnFrom; r[2] = idnLen; r[3] = seqFrom; r[4] = seqLen;
            r[5] = plsFrom; r[6] = plsLen; r[7] = qltFrom; r[8] = qltLen;
        }
        if (digest) { /* :41-43 the per-record copy, folded into a hash */
            for (int64_t j = start; j < end; j++) {
                h ^= (uint64_t)cm_get(&raw, j, &oob);
                h *= 1099511628211ULL;
            }
        }
        n++;
    }
    if (digest) *digest = h;
    return n;
}

/* ------------------------------------------------------- DecompressAll ----- */

/*
 * One chunk of DecompressAll in canonical order:
 * LazyFileReader.cs:53-69 byte range -> Core.ExtractDeflateIndex (Core.cs:133) ->
 * Parsing.Parse over CombinedMemory(from.offset, whole rented buf)
 * (BatchedFASTQ.cs:65-68).  `buf` must hold ora_rent_size(len) bytes and is
 * cleared first (the pool buffers are returned cleared, BatchedFASTQ.cs:71-74).
 * do_copy != 0 also performs the per-record Rent+CopyTo of Parsing.cs:41-43
 * (as a malloc/memcpy/free) so that the timed baseline does the reference's work.
 */
int64_t ora_chunk(const uint8_t *gz, size_t gz_len, const ora_index *ix, int k, uint8_t *buf, int64_t *recs,
                  int64_t cap, uint64_t *digest, int64_t *produced_out, int do_copy)
{
    const ora_point *from = ix->pts[k], *to = ix->pts[k + 1];
    int64_t len_in = to->input - from->input + 1; /* LazyFileReader.cs:63 */
    int64_t pos = from->input - 1;                /* :68 */
    if (pos < 0) pos = 0;
    if (pos + len_in > (int64_t)gz_len) len_in = (int64_t)gz_len - pos; /* short read at EOF */
    int64_t len = to->output - from->output;
    int64_t rent = ora_rent_size(len);
    memset(buf, 0, (size_t)rent);
    int64_t produced = ora_extract(gz + pos, len_in, ix, k, k + 1, buf);
    if (produced_out) *produced_out = produced;
    if (produced < 0) return produced;
    int64_t n = ora_parse(from->offset, from->offset_len, buf, rent, recs, cap, digest);
    if (do_copy) {
        /* Parsing.cs:41-43: Rent(end-start) + CopyTo for every record */
        int64_t tmp[9];
        (void)tmp;
        int64_t m = n < cap ? n : cap;
        volatile uint8_t sink = 0;
        for (int64_t r = 0; r < m; r++) {
            const int64_t *q = recs + 9 * r;
            int64_t start = q[0], end = q[7] + q[8] + 1;
            uint8_t *mem = (uint8_t *)malloc((size_t)ora_rent_size(end - start));
            for (int64_t j = start; j < end; j++) {
                mem[j - start] = j < from->offset_len ? from->offset[j] : buf[j - from->offset_len];
            }
            sink ^= mem[0];
            free(mem);
        }
    }
    return n;
}

/* Thread-pool DecompressAll used as the host baseline:
 * T workers pull chunks in index order, each doing ora_chunk with the
 * per-record copy.  Returns total records; *bytes_out = total bytes inflated. */
typedef struct {
    const uint8_t *gz; size_t gz_len; const ora_index *ix;
    int first, last; int next; pthread_mutex_t mu;
    int64_t records, bytes; int err;
} pool_job;

static void *pool_worker(void *arg)
{
    pool_job *j = (pool_job *)arg;
    int64_t maxlen = 0;
    for (int k = j->first; k < j->last; k++) {
        int64_t l = j->ix->pts[k + 1]->output - j->ix->pts[k]->output;
        if (l > maxlen) maxlen = l;
    }
    uint8_t *buf = (uint8_t *)malloc((size_t)ora_rent_size(maxlen) + 16);
    int64_t cap = maxlen / 4 + 16;
    int64_t *recs = (int64_t *)malloc(sizeof(int64_t) * 9 * (size_t)cap);
    int64_t records = 0, bytes = 0;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        int k = j->next < j->last ? j->next++ : -1;
        pthread_mutex_unlock(&j->mu);
        if (k < 0) break;
        int64_t produced = 0;
        int64_t n = ora_chunk(j->gz, j->gz_len, j->ix, k, buf, recs, cap, NULL, &produced, 1);
        if (n < 0) { j->err = (int)n; break; }
        records += n;
        bytes += produced;
    }
    pthread_mutex_lock(&j->mu);
    j->records += records;
    j->bytes += bytes;
    pthread_mutex_unlock(&j->mu);
    free(buf);
    free(recs);
    return NULL;
}

int64_t ora_decompress_all_mt(const uint8_t *gz, size_t gz_len, const ora_index *ix, int first_chunk, int n_chunks,
                              int threads, int64_t *bytes_out)
{
    pool_job j;
    memset(&j, 0, sizeof j);
    j.gz = gz; j.gz_len = gz_len; j.ix = ix;
    j.first = first_chunk; j.last = first_chunk + n_chunks; j.next = first_chunk;
    pthread_mutex_init(&j.mu, NULL);
    if (threads < 1) threads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, pool_worker, &j);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th);
    pthread_mutex_destroy(&j.mu);
    if (bytes_out) *bytes_out = j.bytes;
    return j.err ? j.err : j.records;
}

/* ------------------------------------------------------ Naive serial path -- */

/*
 * SimpleDecompressor/SimpleDecompressor.cs:9-29 + SimpleDecompressor/Parsing.cs:9-49:
 * GZipStream read in 64 KB arrays, strict byte-wise parser that materialises four
 * strings per record ('\r' or '\n' ends a line, :50; a record not starting with
 * '@'/'+' throws, :24,29 -> returns -1).  Returns the record count.
 */
int64_t ora_naive_count(const uint8_t *gz, size_t gz_len, int64_t *bytes_out)
{
    z_stream strm;
    memset(&strm, 0, sizeof strm);
    if (inflateInit2(&strm, 47) != Z_OK) return -1;
    strm.next_in = (Bytef *)gz;
    strm.avail_in = (uInt)(gz_len > 0xFFFFFFFFu ? 0xFFFFFFFFu : gz_len);
    size_t fed = strm.avail_in;
    const int len = 65536; /* SimpleDecompressor.cs:19 */
    uint8_t *buffer = (uint8_t *)malloc((size_t)len);
    char *field[4];
    size_t fcap[4], flen[4];
    for (int f = 0; f < 4; f++) { fcap[f] = 1024; field[f] = (char *)malloc(fcap[f]); flen[f] = 0; }
    int line = 0, at_line_start = 1;
    int64_t records = 0, bytes = 0;
    int ret = Z_OK, bad = 0;
    while (ret != Z_STREAM_END && !bad) {
        if (strm.avail_in == 0 && fed < gz_len) {
            size_t n = gz_len - fed > 0x40000000u ? 0x40000000u : gz_len - fed;
            strm.next_in = (Bytef *)(gz + fed);
            strm.avail_in = (uInt)n;
            fed += n;
        }
        strm.next_out = buffer;
        strm.avail_out = (uInt)len;
        ret = inflate(&strm, Z_NO_FLUSH);
        if (ret != Z_OK && ret != Z_STREAM_END) { bad = 1; break; }
        int got = len - (int)strm.avail_out;
        bytes += got;
        for (int i = 0; i < got; i++) {
            uint8_t c = buffer[i];
            if (at_line_start && (line == 0 || line == 2)) { /* Parsing.cs:24,28-29 */
                if (line == 0 && c == 0) { ret = Z_STREAM_END; break; } /* :18 */
                if (c != (line == 0 ? '@' : '+')) { bad = 1; break; }
                at_line_start = 0;
                continue;
            }
            at_line_start = 0;
            if (c == '\n' || c == '\r') { /* :50 */
                /* sb.ToString(): materialise the field */
                char *s = (char *)malloc(flen[line] + 1);
                memcpy(s, field[line], flen[line]);
                s[flen[line]] = 0;
                free(s);
                flen[line] = 0;
                line++;
                at_line_start = 1;
                if (line == 4) { line = 0; records++; }
            } else {
                if (flen[line] == fcap[line]) { fcap[line] *= 2; field[line] = (char *)realloc(field[line], fcap[line]); }
                field[line][flen[line]++] = (char)c;
            }
        }
    }
    inflateEnd(&strm);
    free(buffer);
    for (int f = 0; f < 4; f++) free(field[f]);
    if (bytes_out) *bytes_out = bytes;
    return bad ? -1 : records;
}

/*
 * SimpleDecompressor/Parsing.cs:9-49 record for record — the INDEPENDENT second restatement used to
 * pin the hot-path parser: on well-formed input both parsers must cut the same four fields out of
 * the same byte stream.  `data` is the whole inflated stream followed by the zero tail the last
 * 64 KB buffer carries (SimpleDecompressor.cs:19-24).  For every record emits 8 int64 into `recs`
 * (capacity `cap` records): idnFrom, idnLen, seqFrom, seqLen, plsFrom, plsLen, qltFrom, qltLen as
 * offsets into the stream (the strings :26-31 build are data[from, from+len)).  '\r' or '\n' ends
 * a line (:48) and exactly one byte is consumed after it (:44).  Returns the record count, or -1
 * where the reference throws (:24,29: a record not starting with '@' / '+').
 */
int64_t ora_naive_records(const uint8_t *data, int64_t n, int64_t *recs, int64_t cap)
{
    int64_t pos = 0, count = 0;
    while (pos < n) {                      /* :15 !raw.IsAtEnd */
        if (data[pos] == 0) break;         /* :18 */
        if (data[pos++] != '@') return -1; /* :24 */
        int64_t f[8];
        for (int line = 0; line < 4; line++) {
            if (line == 2) {               /* :28-29 skip + */
                if (pos >= n || data[pos++] != '+') return -1;
            }
            int64_t from = pos;
            while (pos < n && data[pos] != '\n' && data[pos] != '\r') pos++; /* :40-43 */
            f[2 * line] = from;
            f[2 * line + 1] = pos - from;
            if (pos < n) pos++;            /* :46 consume \n */
        }
        if (recs && count < cap) memcpy(recs + 8 * count, f, sizeof f);
        count++;
    }
    return count;
}

/* ------------------------------------------------------------ digests ------- */

/*
 * Order-sensitive, parallel-friendly digests (sums mod 2^64 of position-keyed terms) that the GPU
 * library computes per chunk (pp_job_digests, include/ppb200.h) so that parity at BASELINE sizes
 * needs only a few integers per chunk to cross PCIe.  The oracle computes the same function over
 * ITS OWN bytes / records; equality of digests is the comparison.
 *   mix(x)            splitmix64 finaliser
 *   bytes:  mix(n) + sum_i mix(i) * (W_i + 1),  W_i = little-endian u64 of bytes [8i, 8i+8), zero padded
 *   fields: sum_r sum_{f<9} mix(9r+f) * (field_{r,f} + 1) over the nine integers of Parsing.cs:20-39
 */
static inline uint64_t ora_mix64(uint64_t x)
{
    uint64_t z = x + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
uint64_t ora_digest_bytes(const uint8_t *p, int64_t n)
{
    uint64_t d = ora_mix64((uint64_t)n);
    int64_t nw = n / 8;
    for (int64_t i = 0; i < nw; i++) {
        uint64_t w;
        memcpy(&w, p + 8 * i, 8);
        d += ora_mix64((uint64_t)i) * (w + 1);
    }
    if (n & 7) {
        uint64_t w = 0;
        memcpy(&w, p + 8 * nw, (size_t)(n & 7));
        d += ora_mix64((uint64_t)nw) * (w + 1);
    }
    return d;
}
uint64_t ora_digest_fields(const int64_t *recs, int64_t n)
{
    uint64_t d = 0;
    for (int64_t r = 0; r < n; r++)
        for (int f = 0; f < 9; f++) d += ora_mix64((uint64_t)(9 * r + f)) * ((uint64_t)recs[9 * r + f] + 1);
    return d;
}

/*
 * Every chunk of [first, first+n) through ora_chunk on `threads` workers, keeping only what a
 * size-independent comparison needs per chunk: out4[4k..4k+3] = inflated length, record count,
 * bytes digest, fields digest.  Test infrastructure for the BASELINE-size parity gates.
 * Returns 0 or the first negative chunk status.
 */
typedef struct {
    const uint8_t *gz; size_t gz_len; const ora_index *ix;
    int first, last; int next; pthread_mutex_t mu; uint64_t *out4; int err;
} dig_job;

static void *dig_worker(void *arg)
{
    dig_job *j = (dig_job *)arg;
    int64_t maxlen = 0, maxoff = 0;
    for (int k = j->first; k < j->last; k++) {
        int64_t l = j->ix->pts[k + 1]->output - j->ix->pts[k]->output;
        if (l > maxlen) maxlen = l;
        if (j->ix->pts[k]->offset_len > maxoff) maxoff = j->ix->pts[k]->offset_len;
    }
    uint8_t *buf = (uint8_t *)malloc((size_t)ora_rent_size(maxlen) + 16);
    int64_t cap = (maxlen + maxoff) / 4 + 16;
    int64_t *recs = (int64_t *)malloc(sizeof(int64_t) * 9 * (size_t)cap);
    for (;;) {
        pthread_mutex_lock(&j->mu);
        int k = j->next < j->last ? j->next++ : -1;
        pthread_mutex_unlock(&j->mu);
        if (k < 0) break;
        int64_t produced = 0;
        int64_t n = ora_chunk(j->gz, j->gz_len, j->ix, k, buf, recs, cap, NULL, &produced, 0);
        if (n < 0) { j->err = (int)n; break; }
        uint64_t *o = j->out4 + 4 * (size_t)(k - j->first);
        o[0] = (uint64_t)produced;
        o[1] = (uint64_t)n;
        o[2] = ora_digest_bytes(buf, produced);
        o[3] = ora_digest_fields(recs, n);
    }
    free(buf);
    free(recs);
    return NULL;
}

int ora_chunk_digests_mt(const uint8_t *gz, size_t gz_len, const ora_index *ix, int first_chunk, int n_chunks,
                         int threads, uint64_t *out4)
{
    dig_job j;
    memset(&j, 0, sizeof j);
    j.gz = gz; j.gz_len = gz_len; j.ix = ix; j.out4 = out4;
    j.first = first_chunk; j.last = first_chunk + n_chunks; j.next = first_chunk;
    pthread_mutex_init(&j.mu, NULL);
    if (threads < 1) threads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, dig_worker, &j);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th);
    pthread_mutex_destroy(&j.mu);
    return j.err;
}

/*
 * The stops of Core.BuildDeflateIndex's inflate(Z_BLOCK) loop (Core.cs:64) at which a checkpoint may be
 * taken (Core.cs:98: data_type & 128, not in the last block): for each, bits[i] = 8*totin - (data_type & 7)
 * — the first bit of the next block, as Index.AddPoint stores it (Input, Bits) — and outs[i] = totout.
 * kinds[i] (may be NULL) = the 2-bit BTYPE of the block that starts there.  Also reports the bit after
 * the final block and the stream length.  Ground truth for the GPU block scanner (pp_scan_blocks).
 * Returns the number of stops (counting continues past cap) or a negative ZResult.
 */
int64_t ora_block_stops(const uint8_t *gz, size_t gz_len, int64_t *bits, int64_t *outs, uint8_t *kinds, int64_t cap,
                        int64_t *end_bit, int64_t *total_out)
{
    z_stream strm;
    memset(&strm, 0, sizeof strm);
    if (inflateInit2(&strm, 47) != Z_OK) return Z_MEM_ERROR;
    uint8_t *window = (uint8_t *)malloc(WINSIZE);
    int64_t totin = 0, totout = 0, n = 0;
    size_t fed = 0;
    int ret = Z_OK;
    strm.avail_out = 0;
    while (ret != Z_STREAM_END) {
        if (strm.avail_in == 0) {
            size_t m = gz_len - fed < CHUNK ? gz_len - fed : CHUNK;
            if (m == 0) { ret = Z_DATA_ERROR; break; }
            strm.next_in = (Bytef *)(gz + fed);
            strm.avail_in = (uInt)m;
            fed += m;
        }
        if (strm.avail_out == 0) { strm.avail_out = WINSIZE; strm.next_out = window; }
        uInt in_before = strm.avail_in, out_before = strm.avail_out;
        ret = inflate(&strm, Z_BLOCK);
        totin += in_before - strm.avail_in;
        totout += out_before - strm.avail_out;
        if (ret != Z_OK && ret != Z_STREAM_END && ret != Z_BUF_ERROR) break;
        if (ret == Z_BUF_ERROR) ret = Z_OK;
        if ((strm.data_type & 128) && !(strm.data_type & 64)) {
            int64_t bit = totin * 8 - (strm.data_type & 7);
            if (n < cap) {
                if (bits) bits[n] = bit;
                if (outs) outs[n] = totout;
                if (kinds) {
                    /* the block header's BTYPE: bits 1..2 counted from `bit` (LSB first) */
                    int64_t b1 = bit + 1;
                    unsigned v = (unsigned)(gz[b1 >> 3] >> (b1 & 7));
                    if ((b1 & 7) == 7 && (size_t)((b1 >> 3) + 1) < gz_len) v |= (unsigned)gz[(b1 >> 3) + 1] << 1;
                    kinds[n] = (uint8_t)(v & 3u);
                }
            }
            n++;
        }
    }
    if (ret == Z_STREAM_END) {
        if (end_bit) *end_bit = totin * 8 - (strm.data_type & 7);
        if (total_out) *total_out = totout;
    }
    inflateEnd(&strm);
    free(window);
    return ret == Z_STREAM_END ? n : (ret < 0 ? ret : Z_DATA_ERROR);
}

/* Whole-stream inflate (zcat) used by tests for concat(chunks) == stream. */
int64_t ora_zcat(const uint8_t *gz, size_t gz_len, uint8_t *out, int64_t out_cap)
{
    z_stream strm;
    memset(&strm, 0, sizeof strm);
    if (inflateInit2(&strm, 47) != Z_OK) return -1;
    size_t fed = 0;
    int64_t total = 0;
    int ret = Z_OK;
    while (ret != Z_STREAM_END) {
        if (strm.avail_in == 0) {
            size_t n = gz_len - fed > 0x40000000u ? 0x40000000u : gz_len - fed;
            if (n == 0) break;
            strm.next_in = (Bytef *)(gz + fed);
            strm.avail_in = (uInt)n;
            fed += n;
        }
        int64_t room = out_cap - total;
        if (room <= 0) break;
        strm.next_out = out + total;
        strm.avail_out = (uInt)(room > 0x40000000 ? 0x40000000 : room);
        uInt before = strm.avail_out;
        ret = inflate(&strm, Z_NO_FLUSH);
        total += before - strm.avail_out;
        if (ret != Z_OK && ret != Z_STREAM_END) { inflateEnd(&strm); return ret; }
    }
    inflateEnd(&strm);
    return total;
}

/*
 * ppgzip — parallel single-member gzip writer for the big benchmark corpora.
 *
 * Not part of the reference (which is fed by system `gzip`); it exists because
 * `gzip -6` compresses Generator data at ~5 MB/s here and the
 * 10 M / 100 M-read configs must be produced inside a gpurun call.
 *
 * pigz-style: the input is cut into segments; each segment is raw-deflated
 * independently (level 6, memLevel 9 so that blocks close every 32 K symbols as
 * gzip's own deflate does) with the previous segment's last 32 KB as preset
 * dictionary, and ends with Z_SYNC_FLUSH (an empty stored block, byte aligned) so
 * the pieces concatenate into ONE deflate stream inside ONE gzip member — the
 * reference cannot extract chunks that span members.  The
 * sync-flush joins also exercise the decoder's stored-block path.
 *
 * usage: ppgzip [-l level] [-t threads] [-s segment_bytes] in out.gz
 */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <zlib.h>

typedef struct {
    const uint8_t *in;
    size_t in_len;
    size_t seg;
    size_t nseg;
    int level;
    uint8_t **out;
    size_t *out_len;
    uLong *crc;
    size_t next;
    pthread_mutex_t mu;
    int err;
} job_t;

static void *worker(void *arg)
{
    job_t *j = (job_t *)arg;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        size_t s = j->next < j->nseg ? j->next++ : (size_t)-1;
        pthread_mutex_unlock(&j->mu);
        if (s == (size_t)-1) break;
        size_t off = s * j->seg;
        size_t len = j->in_len - off < j->seg ? j->in_len - off : j->seg;
        int last = (s + 1 == j->nseg);
        z_stream z;
        memset(&z, 0, sizeof z);
        if (deflateInit2(&z, j->level, Z_DEFLATED, -15, 9, Z_DEFAULT_STRATEGY) != Z_OK) { j->err = 1; break; }
        if (off > 0) {
            size_t d = off < 32768 ? off : 32768;
            deflateSetDictionary(&z, j->in + off - d, (uInt)d);
        }
        size_t cap = deflateBound(&z, (uLong)len) + 64;
        uint8_t *o = (uint8_t *)malloc(cap);
        z.next_in = (Bytef *)(j->in + off);
        z.avail_in = (uInt)len;
        z.next_out = o;
        z.avail_out = (uInt)cap;
        int rc = deflate(&z, last ? Z_FINISH : Z_SYNC_FLUSH);
        if ((last && rc != Z_STREAM_END) || (!last && rc != Z_OK) || z.avail_in != 0) { j->err = 2; free(o); deflateEnd(&z); break; }
        j->out[s] = o;
        j->out_len[s] = cap - z.avail_out;
        j->crc[s] = crc32(crc32(0L, Z_NULL, 0), j->in + off, (uInt)len);
        deflateEnd(&z);
    }
    return NULL;
}

int ppgzip_buffer(const uint8_t *in, size_t in_len, int level, int threads, size_t seg, FILE *f)
{
    job_t j;
    memset(&j, 0, sizeof j);
    if (seg < 65536) seg = 65536;
    j.in = in; j.in_len = in_len; j.seg = seg; j.level = level;
    j.nseg = in_len ? (in_len + seg - 1) / seg : 1;
    j.out = (uint8_t **)calloc(j.nseg, sizeof(*j.out));
    j.out_len = (size_t *)calloc(j.nseg, sizeof(*j.out_len));
    j.crc = (uLong *)calloc(j.nseg, sizeof(*j.crc));
    pthread_mutex_init(&j.mu, NULL);
    if (threads < 1) threads = 1;
    if ((size_t)threads > j.nseg) threads = (int)j.nseg;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, worker, &j);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th);
    if (j.err) return -1;
    static const uint8_t hdr[10] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 3};
    fwrite(hdr, 1, 10, f);
    uLong crc = crc32(0L, Z_NULL, 0);
    for (size_t s = 0; s < j.nseg; s++) {
        size_t off = s * seg;
        size_t len = in_len - off < seg ? in_len - off : seg;
        fwrite(j.out[s], 1, j.out_len[s], f);
        crc = crc32_combine(crc, j.crc[s], (z_off_t)len);
        free(j.out[s]);
    }
    uint8_t tr[8];
    uint32_t c = (uint32_t)crc, n = (uint32_t)in_len;
    for (int i = 0; i < 4; i++) { tr[i] = (uint8_t)(c >> (8 * i)); tr[4 + i] = (uint8_t)(n >> (8 * i)); }
    fwrite(tr, 1, 8, f);
    free(j.out); free(j.out_len); free(j.crc);
    pthread_mutex_destroy(&j.mu);
    return 0;
}

#ifndef PPGZIP_NO_MAIN
int main(int argc, char **argv)
{
    int level = 6, threads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    size_t seg = 8u << 20;
    int i = 1;
    for (; i < argc && argv[i][0] == '-' && argv[i][1]; i++) {
        if (!strcmp(argv[i], "-l") && i + 1 < argc) level = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-t") && i + 1 < argc) threads = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-s") && i + 1 < argc) seg = strtoull(argv[++i], NULL, 10);
        else break;
    }
    if (argc - i != 2) { fprintf(stderr, "usage: ppgzip [-l level] [-t threads] [-s segment_bytes] in out.gz\n"); return 2; }
    FILE *fi = fopen(argv[i], "rb");
    if (!fi) { perror("ppgzip: in"); return 1; }
    fseek(fi, 0, SEEK_END);
    size_t n = (size_t)ftell(fi);
    fseek(fi, 0, SEEK_SET);
    uint8_t *in = (uint8_t *)malloc(n ? n : 1);
    if (fread(in, 1, n, fi) != n) { perror("ppgzip: read"); return 1; }
    fclose(fi);
    FILE *fo = fopen(argv[i + 1], "wb");
    if (!fo) { perror("ppgzip: out"); return 1; }
    int rc = ppgzip_buffer(in, n, level, threads, seg, fo);
    fclose(fo);
    free(in);
    return rc ? 1 : 0;
}
#endif

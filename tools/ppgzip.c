/*
 * ppgzip — parallel, streaming, single-member gzip writer for the big benchmark corpora.
 *
 * Not part of the reference (which is fed by system `gzip`); it exists because
 * `gzip -6` compresses Generator data at ~5 MB/s here and the
 * 10 M / 100 M-read configs must be produced inside a benchmark run.
 *
 * pigz-style: the input is cut into segments; each segment is raw-deflated
 * independently (level 6, memLevel 9 so that blocks close every 32 K symbols as
 * gzip's own deflate does) with the previous segment's last 32 KB as preset
 * dictionary, and ends with Z_SYNC_FLUSH (an empty stored block, byte aligned) so
 * the pieces concatenate into ONE deflate stream inside ONE gzip member — the
 * reference cannot extract chunks that span members.  The
 * sync-flush joins also exercise the decoder's stored-block path.
 *
 * Streaming: the main thread reads segments (file or stdin, so `ppgen | ppgzip -`
 * overlaps generation with compression), workers compress, results are written
 * in order; at most 3*threads segments are in flight.
 *
 * usage: ppgzip [-l level] [-t threads] [-s segment_bytes] in|- out.gz
 */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <zlib.h>

typedef struct seg {
    uint8_t *in;
    size_t in_len;
    uint8_t dict[32768];
    size_t dict_len;
    int last;
    uint8_t *out;
    size_t out_len;
    uLong crc;
    int done;
    struct seg *next_todo;
} seg_t;

typedef struct {
    int level;
    pthread_mutex_t mu;
    pthread_cond_t cv_todo, cv_done;
    seg_t *todo_head, *todo_tail;
    int closing, err;
} pool_t;

static void *worker(void *arg)
{
    pool_t *p = (pool_t *)arg;
    for (;;) {
        pthread_mutex_lock(&p->mu);
        while (!p->todo_head && !p->closing) pthread_cond_wait(&p->cv_todo, &p->mu);
        seg_t *s = p->todo_head;
        if (s) {
            p->todo_head = s->next_todo;
            if (!p->todo_head) p->todo_tail = NULL;
        }
        pthread_mutex_unlock(&p->mu);
        if (!s) break;
        z_stream z;
        memset(&z, 0, sizeof z);
        int bad = deflateInit2(&z, p->level, Z_DEFLATED, -15, 9, Z_DEFAULT_STRATEGY) != Z_OK;
        if (!bad) {
            if (s->dict_len) deflateSetDictionary(&z, s->dict, (uInt)s->dict_len);
            size_t cap = deflateBound(&z, (uLong)s->in_len) + 64;
            s->out = (uint8_t *)malloc(cap);
            z.next_in = s->in;
            z.avail_in = (uInt)s->in_len;
            z.next_out = s->out;
            z.avail_out = (uInt)cap;
            int rc = deflate(&z, s->last ? Z_FINISH : Z_SYNC_FLUSH);
            if ((s->last && rc != Z_STREAM_END) || (!s->last && rc != Z_OK) || z.avail_in != 0) bad = 1;
            s->out_len = cap - z.avail_out;
            s->crc = crc32(crc32(0L, Z_NULL, 0), s->in, (uInt)s->in_len);
            deflateEnd(&z);
        }
        free(s->in);
        s->in = NULL;
        pthread_mutex_lock(&p->mu);
        if (bad) p->err = 1;
        s->done = 1;
        pthread_cond_broadcast(&p->cv_done);
        pthread_mutex_unlock(&p->mu);
    }
    return NULL;
}

static size_t read_full(FILE *f, uint8_t *buf, size_t n)
{
    size_t got = 0;
    while (got < n) {
        size_t r = fread(buf + got, 1, n - got, f);
        if (r == 0) break;
        got += r;
    }
    return got;
}

int ppgzip_stream(FILE *fi, FILE *fo, int level, int threads, size_t seg_bytes)
{
    if (seg_bytes < 65536) seg_bytes = 65536;
    if (threads < 1) threads = 1;
    pool_t p;
    memset(&p, 0, sizeof p);
    p.level = level;
    pthread_mutex_init(&p.mu, NULL);
    pthread_cond_init(&p.cv_todo, NULL);
    pthread_cond_init(&p.cv_done, NULL);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, worker, &p);

    const int window = 3 * threads;  /* segments in flight */
    seg_t **ring = (seg_t **)calloc((size_t)window, sizeof(seg_t *));
    size_t n_read = 0, n_written = 0;
    static const uint8_t hdr[10] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 3};
    fwrite(hdr, 1, 10, fo);
    uLong crc = crc32(0L, Z_NULL, 0);
    uint64_t total = 0;
    uint8_t tail[32768];
    size_t tail_len = 0;
    /* one segment of read-ahead so that we know which one is last */
    uint8_t *cur = (uint8_t *)malloc(seg_bytes);
    size_t cur_len = read_full(fi, cur, seg_bytes);
    int eof = 0, rc = 0;
    while (!eof || n_written < n_read) {
        /* write finished segments in order; block when the window is full or input is over */
        while (n_written < n_read) {
            seg_t *s = ring[n_written % (size_t)window];
            pthread_mutex_lock(&p.mu);
            int must_wait = eof || (n_read - n_written) >= (size_t)window;
            while (!s->done && must_wait) pthread_cond_wait(&p.cv_done, &p.mu);
            int done = s->done;
            pthread_mutex_unlock(&p.mu);
            if (!done) break;
            if (s->out) fwrite(s->out, 1, s->out_len, fo);
            crc = crc32_combine(crc, s->crc, (z_off_t)s->in_len);
            total += s->in_len;
            free(s->out);
            free(s);
            n_written++;
        }
        if (eof) continue;
        uint8_t *nxt = (uint8_t *)malloc(seg_bytes);
        size_t nxt_len = cur_len == seg_bytes ? read_full(fi, nxt, seg_bytes) : 0;
        seg_t *s = (seg_t *)calloc(1, sizeof(seg_t));
        s->in = cur;
        s->in_len = cur_len;
        memcpy(s->dict, tail, tail_len);
        s->dict_len = tail_len;
        s->last = nxt_len == 0;
        /* remember this segment's last 32 KB as the next one's dictionary */
        if (cur_len >= 32768) {
            memcpy(tail, cur + cur_len - 32768, 32768);
            tail_len = 32768;
        } else {
            size_t keep = 32768 - cur_len < tail_len ? 32768 - cur_len : tail_len;
            memmove(tail, tail + tail_len - keep, keep);
            memcpy(tail + keep, cur, cur_len);
            tail_len = keep + cur_len;
        }
        ring[n_read % (size_t)window] = s;
        n_read++;
        pthread_mutex_lock(&p.mu);
        if (p.todo_tail) p.todo_tail->next_todo = s; else p.todo_head = s;
        p.todo_tail = s;
        pthread_cond_signal(&p.cv_todo);
        pthread_mutex_unlock(&p.mu);
        if (s->last) { eof = 1; free(nxt); }
        else { cur = nxt; cur_len = nxt_len; }
    }
    pthread_mutex_lock(&p.mu);
    p.closing = 1;
    pthread_cond_broadcast(&p.cv_todo);
    pthread_mutex_unlock(&p.mu);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th);
    free(ring);
    if (p.err) rc = -1;
    uint8_t tr[8];
    uint32_t c = (uint32_t)crc, n = (uint32_t)total;
    for (int i = 0; i < 4; i++) { tr[i] = (uint8_t)(c >> (8 * i)); tr[4 + i] = (uint8_t)(n >> (8 * i)); }
    fwrite(tr, 1, 8, fo);
    pthread_mutex_destroy(&p.mu);
    pthread_cond_destroy(&p.cv_todo);
    pthread_cond_destroy(&p.cv_done);
    return rc;
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
    if (argc - i != 2) { fprintf(stderr, "usage: ppgzip [-l level] [-t threads] [-s segment_bytes] in|- out.gz\n"); return 2; }
    FILE *fi = strcmp(argv[i], "-") ? fopen(argv[i], "rb") : stdin;
    if (!fi) { perror("ppgzip: in"); return 1; }
    FILE *fo = fopen(argv[i + 1], "wb");
    if (!fo) { perror("ppgzip: out"); return 1; }
    setvbuf(fo, NULL, _IOFBF, 1 << 22);
    int rc = ppgzip_stream(fi, fo, level, threads, seg);
    if (fclose(fo) != 0) rc = -1;
    if (fi != stdin) fclose(fi);
    return rc ? 1 : 0;
}
#endif

/* TEST SCAFFOLDING: drives libppb200.so from plain C, the way a P/Invoke binding sees it
 * (blittable arguments only, caller-owned buffers, integer return codes): CreateIndex,
 * Serialize / Deserialize (both file versions), the point accessors, error strings — and
 * the loud failure of every decode entry point on a machine without an sm_100 GPU.
 *   harness <file.gz> <chunksize> <tmpdir>   prints "key value" lines; exit 0 on success */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ppb200.h"

#define CHECK(x)                                                             \
    do {                                                                     \
        int rc_ = (x);                                                       \
        if (rc_ < 0) {                                                       \
            fprintf(stderr, "%s -> %d (%s)\n", #x, rc_, pp_strerror(rc_));   \
            return 1;                                                        \
        }                                                                    \
    } while (0)

int main(int argc, char **argv)
{
    if (argc != 4) return 2;
    char p0[512], p1[512];
    snprintf(p0, sizeof p0, "%s/c_v0.gzi", argv[3]);
    snprintf(p1, sizeof p1, "%s/c_v1.gzi", argv[3]);
    printf("abi %d\n", pp_abi_version());

    pp_index *ix = NULL, *a = NULL, *b = NULL;
    CHECK(pp_index_create_file(argv[1], (uint32_t)atoi(argv[2]), 0, &ix));
    const int n = pp_index_count(ix);
    printf("points %d\nchunk_max_bytes %d\n", n, pp_index_chunk_max_bytes(ix));
    CHECK(pp_index_serialize(ix, p0));
    CHECK(pp_index_serialize_v1(ix, p1));
    CHECK(pp_index_deserialize(p0, &a));
    CHECK(pp_index_deserialize(p1, &b));
    if (pp_index_count(a) != n || pp_index_count(b) != n) return 3;
    unsigned long long sum = 0;
    for (int i = 0; i < n; i++) {
        pp_point p, q, r;
        CHECK(pp_index_point(ix, i, &p));
        CHECK(pp_index_point(a, i, &q));
        CHECK(pp_index_point(b, i, &r));
        if (p.output != q.output || p.input != q.input || p.bits != q.bits || p.offset_len != q.offset_len) return 4;
        if (p.output != r.output || p.input != r.input || p.bits != r.bits || p.offset_len != r.offset_len) return 5;
        if (memcmp(p.window, q.window, 32768) || memcmp(p.window, r.window, 32768)) return 6;
        if (p.offset_len && (memcmp(p.offset, q.offset, (size_t)p.offset_len) || memcmp(p.offset, r.offset, (size_t)p.offset_len)))
            return 7;
        sum += (unsigned long long)p.output * 31u + (unsigned long long)p.input * 7u + (unsigned)p.bits;
    }
    printf("point_sum %llu\n", sum);

    /* an index built point by point (Index.Add) equals the original */
    pp_index *c = NULL;
    CHECK(pp_index_new(&c));
    for (int i = 0; i < n; i++) {
        pp_point p;
        CHECK(pp_index_point(ix, i, &p));
        CHECK(pp_index_add(c, p.bits, p.input, p.output, p.window, p.offset, p.offset_len));
    }
    printf("rebuilt_points %d\n", pp_index_count(c));

    /* the multi-GPU split (host code): contiguous, ordered, covering */
    {
        int32_t first[5], cnt[5], next = 0, total = 0;
        CHECK(pp_partition_chunks(ix, 5, first, cnt));
        for (int r = 0; r < 5; r++) {
            if (first[r] != next || cnt[r] < 0) return 8;
            next += cnt[r];
            total += cnt[r];
        }
        if (total != n - 1) return 9;
        printf("partition5 %d %d %d %d %d\n", cnt[0], cnt[1], cnt[2], cnt[3], cnt[4]);
    }

    /* the device: either it opens (B200 box) or every entry point says so — never a CPU fallback */
    pp_ctx *ctx = NULL;
    const int rc = pp_open(0, &ctx);
    printf("open %d %s\n", rc, pp_strerror(rc));
    if (rc == 0) pp_close(ctx);
    {
        /* the one-call entry points over several GPUs / paired files fail the same loud way without a device */
        const int32_t devs[1] = {0};
        pp_multi *m = NULL;
        pp_pair *pr = NULL;
        const uint8_t dummy[16] = {0};
        const int rm = pp_decompress_all_multi(devs, 1, ix, dummy, sizeof dummy, 0, &m);
        const int rp = pp_pair_decompress_all(devs, 1, ix, dummy, sizeof dummy, ix, dummy, sizeof dummy, 0, &pr);
        printf("multi_nodata %d\npair_nodata %d\n", rm, rp);
        /* CreateIndex on the GPU needs a context: without one it is an argument error, the host pass stays available */
        pp_index *gi = NULL;
        pp_create_stats cs;
        printf("create_gpu_noctx %d\n", pp_index_create_gpu(NULL, dummy, sizeof dummy, 1000, 0, &gi, &cs));
        if (m) pp_multi_free(m);
        if (pr) pp_pair_free(pr);
    }
    pp_index_free(c);
    pp_index_free(b);
    pp_index_free(a);
    pp_index_free(ix);
    return 0;
}

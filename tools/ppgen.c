/*
 * ppgen — synthetic FASTQ corpus writer.
 *
 * Restates the reference's Generator (Generator/Generator.cs:8-61) bit-exactly,
 * including the seeded .NET System.Random it draws from (`new Random(0)`,
 * Generator.cs:8).  The seeded System.Random is the runtime's Knuth subtractive
 * generator (Net5CompatSeedImpl); its source is not in the reference tree, so it
 * is restated from the published algorithm and pinned by the known answers in
 * tests/test_generator.py (Random(0).Next()==1559595546, Random(42).Next()==1434747710).
 *
 * Draw order per record (Generator.cs:14-18): Next(128,512) -> '@' id line
 * (Next(1e7,2e7), :41) -> L x NextDouble (bases, :28-32) -> '+' id line (another
 * Next(1e7,2e7)) -> L x NextDouble (qualities, :53-56).
 *
 * Extensions the reference lacks, all documented:
 *   --fixed L        read length fixed to L; the Next(128,512) draw is still consumed
 *   --lognormal M S  length = round(exp(N(mu,S))) with mean M (mu = ln M - S^2/2),
 *                    Box-Muller on two extra NextDouble draws in place of Next(128,512)
 *   --cap C          clamp lengths to [1, C]  (reference-legal long reads: C<=16000, H2)
 *   --seed S         Random(S) instead of Random(0) (R2 of a pair, per-shard seeds)
 *   --first N0       record numbering starts at N0 (per-shard corpora)
 *
 * usage: ppgen <reads> [options] > out.fastq      (or  -o path)
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MBIG 2147483647
#define MSEED 161803398

typedef struct {
    int32_t sa[56];
    int inext, inextp;
} dotnet_random;

static void rng_init(dotnet_random *r, int32_t seed)
{
    int ii = 0;
    int32_t mj, mk;
    int32_t subtraction = (seed == INT32_MIN) ? INT32_MAX : (seed < 0 ? -seed : seed);
    mj = MSEED - subtraction;
    memset(r->sa, 0, sizeof r->sa);
    r->sa[55] = mj;
    mk = 1;
    for (int i = 1; i < 55; i++) {
        if ((ii += 21) >= 55) ii -= 55;
        r->sa[ii] = mk;
        mk = mj - mk;
        if (mk < 0) mk += MBIG;
        mj = r->sa[ii];
    }
    for (int k = 1; k < 5; k++) {
        for (int i = 1; i < 56; i++) {
            int n = i + 30;
            if (n >= 55) n -= 55;
            r->sa[i] -= r->sa[1 + n];
            if (r->sa[i] < 0) r->sa[i] += MBIG;
        }
    }
    r->inext = 0;
    r->inextp = 21;
}

static inline int32_t rng_internal_sample(dotnet_random *r)
{
    int locINext = r->inext, locINextp = r->inextp;
    if (++locINext >= 56) locINext = 1;
    if (++locINextp >= 56) locINextp = 1;
    int32_t retVal = r->sa[locINext] - r->sa[locINextp];
    if (retVal == MBIG) retVal--;
    if (retVal < 0) retVal += MBIG;
    r->sa[locINext] = retVal;
    r->inext = locINext;
    r->inextp = locINextp;
    return retVal;
}

static inline double rng_next_double(dotnet_random *r) { return rng_internal_sample(r) * (1.0 / MBIG); }

static inline int32_t rng_next_range(dotnet_random *r, int32_t lo, int32_t hi)
{
    /* range <= int.MaxValue branch of Net5CompatSeedImpl.Next(min,max) */
    int64_t range = (int64_t)hi - lo;
    return (int32_t)(rng_next_double(r) * (double)range) + lo;
}

/* exposed for the known-answer tests (tests load ppgen as a shared object too) */
int32_t ppgen_kat_next(int32_t seed)
{
    dotnet_random r;
    rng_init(&r, seed);
    return rng_internal_sample(&r);
}
int32_t ppgen_kat_next_range(int32_t seed, int32_t lo, int32_t hi, double *following_double)
{
    dotnet_random r;
    rng_init(&r, seed);
    int32_t v = rng_next_range(&r, lo, hi);
    if (following_double) *following_double = rng_next_double(&r);
    return v;
}

typedef struct {
    uint64_t reads;
    int32_t seed;
    int fixed_len;     /* 0 = native */
    double ln_mean, ln_sigma; /* lognormal if ln_mean > 0 */
    int cap;
    uint64_t first;
} gen_opts;

static size_t put_u64(char *p, uint64_t v)
{
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    for (int i = 0; i < n; i++) p[i] = tmp[n - 1 - i];
    return (size_t)n;
}

/* Generator.cs:39-46  GenerateSrrId */
static size_t gen_id(char *p, dotnet_random *rng, int seqlen, uint64_t no, char prefix)
{
    int32_t id = rng_next_range(rng, 10000000, 20000000);
    uint64_t major = no / 2 + 1, minor = no % 2 + 1;
    char *s = p;
    *p++ = prefix; *p++ = 'S'; *p++ = 'R'; *p++ = 'R';
    p += put_u64(p, (uint64_t)id);
    *p++ = '.';
    p += put_u64(p, major);
    *p++ = '.';
    p += put_u64(p, minor);
    *p++ = ' ';
    p += put_u64(p, major);
    memcpy(p, " length=", 8); p += 8;
    p += put_u64(p, (uint64_t)seqlen);
    *p++ = '\n';
    return (size_t)(p - s);
}

/* Smallest sample s with s*(1.0/MBIG) >= x, so that `NextDouble() < x` (the
 * comparisons of Generator.cs:29-31,54-55) becomes the branch-free `s < T(x)`
 * with bit-identical results. */
static int32_t threshold(double x)
{
    int32_t lo = 0, hi = MBIG;
    while (lo < hi) {
        int32_t mid = lo + (hi - lo) / 2;
        if (mid * (1.0 / MBIG) >= x) hi = mid; else lo = mid + 1;
    }
    return lo;
}

int ppgen_run(const gen_opts *o, FILE *out)
{
    dotnet_random rng;
    rng_init(&rng, o->seed);
    const int32_t t25 = threshold(0.25), t50 = threshold(0.5), t75 = threshold(0.75);
    const int32_t t90 = threshold(0.9), t95 = threshold(0.95);
    size_t cap = 1 << 22, len = 0;
    char *buf = (char *)malloc(cap);
    double mu = 0;
    if (o->ln_mean > 0) mu = log(o->ln_mean) - 0.5 * o->ln_sigma * o->ln_sigma;
    for (uint64_t i = 0; i < o->reads; i++) {
        uint64_t no = o->first + i;
        int L;
        if (o->ln_mean > 0) {
            double u1 = rng_next_double(&rng), u2 = rng_next_double(&rng);
            if (u1 < 1e-300) u1 = 1e-300;
            double z = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
            L = (int)llround(exp(mu + o->ln_sigma * z));
        } else {
            L = rng_next_range(&rng, 128, 512); /* Generator.cs:14 */
            if (o->fixed_len > 0) L = o->fixed_len;
        }
        if (L < 1) L = 1;
        if (o->cap > 0 && L > o->cap) L = o->cap;
        size_t need = 2 * (size_t)L + 256;
        if (len + need > cap) {
            if (fwrite(buf, 1, len, out) != len) return -1;
            len = 0;
            if (need > cap) { cap = need * 2; buf = (char *)realloc(buf, cap); }
        }
        len += gen_id(buf + len, &rng, L, no, '@'); /* :15 */
        for (int j = 0; j < L; j++) {                /* :23-37 */
            int32_t s = rng_internal_sample(&rng);
            buf[len++] = "ATCG"[(s >= t25) + (s >= t50) + (s >= t75)];
        }
        buf[len++] = '\n';
        len += gen_id(buf + len, &rng, L, no, '+'); /* :17 */
        for (int j = 0; j < L; j++) {                /* :48-61 */
            int32_t s = rng_internal_sample(&rng);
            buf[len++] = "?*!"[(s >= t90) + (s >= t95)];
        }
        buf[len++] = '\n';
    }
    if (len && fwrite(buf, 1, len, out) != len) return -1;
    free(buf);
    return 0;
}

#ifndef PPGEN_NO_MAIN
int main(int argc, char **argv)
{
    gen_opts o;
    memset(&o, 0, sizeof o);
    const char *path = NULL;
    if (argc < 2) {
        fprintf(stderr, "usage: ppgen <reads> [--fixed L] [--lognormal MEAN SIGMA] [--cap C] [--seed S] [--first N0] [-o path]\n");
        return 2;
    }
    o.reads = strtoull(argv[1], NULL, 10);
    for (int i = 2; i < argc; i++) {
        if (!strcmp(argv[i], "--fixed") && i + 1 < argc) o.fixed_len = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--lognormal") && i + 2 < argc) { o.ln_mean = atof(argv[++i]); o.ln_sigma = atof(argv[++i]); }
        else if (!strcmp(argv[i], "--cap") && i + 1 < argc) o.cap = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--seed") && i + 1 < argc) o.seed = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--first") && i + 1 < argc) o.first = strtoull(argv[++i], NULL, 10);
        else if (!strcmp(argv[i], "-o") && i + 1 < argc) path = argv[++i];
        else { fprintf(stderr, "ppgen: bad argument %s\n", argv[i]); return 2; }
    }
    FILE *out = path ? fopen(path, "wb") : stdout;
    if (!out) { perror("ppgen"); return 1; }
    setvbuf(out, NULL, _IOFBF, 1 << 22);
    int rc = ppgen_run(&o, out);
    if (path) fclose(out); else fflush(out);
    return rc ? 1 : 0;
}
#endif

/*
 * evdb_oracle.c -- CPU restatement of ErlVectorDB's search hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (erlvectordb_b200/,
 * include/) may link, import or call this file.  It is the parity checker for
 * tests/, for __graft_entry__.smoke() and for bench.py's cpu_baseline /
 * --impl reference legs.
 *
 * The reference is pure Erlang and no Erlang/OTP toolchain exists in this
 * image, so this is a "port" oracle (not a compiled copy of the reference).
 * Every function cites the reference text it restates (paths relative to the
 * reference tree).  It is pinned against the known-answer vectors derived
 * from the reference's own test fixtures (tests/test_oracle_kat.py):
 *   test/vector_store_SUITE.erl:66-87, test/persistence_SUITE.erl:88-166,
 *   test/compression_SUITE.erl:43-82, examples/mcp_client.py:300-317.
 *
 * Arithmetic contract (OTP semantics relied on):
 *   - Erlang float() is IEEE-754 binary64; no FMA contraction, no
 *     re-association: build with -O2 -ffp-contract=off and WITHOUT -ffast-math.
 *   - lists:sum/1 is a left fold starting from integer 0: ((0+p0)+p1)+...
 *   - math:sqrt/1 is C sqrt (correctly rounded).
 *   - erlang:round/1 rounds half away from zero == C round().
 *   - lists:sort/1 on {Distance, Id, Entry}: Distance by numeric value, then
 *     Id by term order.  Callers pass an integer rank per row that realises
 *     the Id order (bench ids are <<Row:64/big>>, so rank == row).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EVO_COSINE 0
#define EVO_EUCLIDEAN 1
#define EVO_MANHATTAN 2

/* ---- src/vector_store.erl:248-249 (same body src/vector_utils.erl:46-47) --
 * dot_product(V1,V2) -> lists:sum([X*Y || {X,Y} <- lists:zip(V1,V2)]).      */
double evo_dot(const double *a, const double *b, int d) {
    double s = 0.0;
    for (int i = 0; i < d; ++i) {
        double p = a[i] * b[i];
        s = s + p;
    }
    return s;
}

/* ---- src/vector_store.erl:251-252 (same src/vector_utils.erl:56-57) -------
 * vector_norm(V) -> math:sqrt(lists:sum([X*X || X <- V])).                  */
double evo_norm(const double *v, int d) {
    double s = 0.0;
    for (int i = 0; i < d; ++i) {
        double p = v[i] * v[i];
        s = s + p;
    }
    return sqrt(s);
}

/* ---- src/vector_store.erl:238-246 ----------------------------------------
 * 1.0 - Dot/(Norm1*Norm2); 1.0 when either norm == 0.0.  The query norm is
 * recomputed on every call, exactly as the reference does.                  */
double evo_cosine_distance(const double *q, const double *v, int d) {
    double dot = evo_dot(q, v, d);
    double n1 = evo_norm(q, d);
    double n2 = evo_norm(v, d);
    if (n1 == 0.0) return 1.0;
    if (n2 == 0.0) return 1.0;
    return 1.0 - (dot / (n1 * n2));
}

/* ---- src/vector_utils.erl:28-36 (similarity, not distance) --------------- */
double evo_cosine_similarity(const double *a, const double *b, int d) {
    double dot = evo_dot(a, b, d);
    double n1 = evo_norm(a, d);
    double n2 = evo_norm(b, d);
    if (n1 == 0.0) return 0.0;
    if (n2 == 0.0) return 0.0;
    return dot / (n1 * n2);
}

/* ---- src/vector_utils.erl:38-40 + :62-63 ---------------------------------
 * euclidean_distance = vector_norm(vector_subtract(V1,V2)): direct form.    */
double evo_euclidean_distance(const double *a, const double *b, int d) {
    double s = 0.0;
    for (int i = 0; i < d; ++i) {
        double t = a[i] - b[i];
        double p = t * t;
        s = s + p;
    }
    return sqrt(s);
}

/* ---- src/vector_utils.erl:42-43 ------------------------------------------
 * manhattan_distance = lists:sum([abs(X-Y) || ...]).                        */
double evo_manhattan_distance(const double *a, const double *b, int d) {
    double s = 0.0;
    for (int i = 0; i < d; ++i) {
        double t = fabs(a[i] - b[i]);
        s = s + t;
    }
    return s;
}

double evo_distance(const double *q, const double *v, int d, int metric) {
    switch (metric) {
        case EVO_EUCLIDEAN: return evo_euclidean_distance(q, v, d);
        case EVO_MANHATTAN: return evo_manhattan_distance(q, v, d);
        default: return evo_cosine_distance(q, v, d);
    }
}

/* ---- src/vector_store.erl:227-236 perform_search/3 -----------------------
 * Distances for ALL n rows, full sort on (Distance, Id), sublist(K).        */
typedef struct {
    double dist;
    uint64_t rank; /* realises Erlang term order of the Id */
    int64_t row;
} evo_cand;

static int evo_cand_cmp(const void *pa, const void *pb) {
    const evo_cand *a = (const evo_cand *)pa, *b = (const evo_cand *)pb;
    if (a->dist < b->dist) return -1;
    if (a->dist > b->dist) return 1;
    if (a->rank < b->rank) return -1;
    if (a->rank > b->rank) return 1;
    return 0;
}

/* rows: n x d fp64 row-major; ranks: n id-ranks or NULL (rank == row).
 * Returns the number of results written (min(k, n)); k < 0 -> -1 (the
 * reference crashes with function_clause in lists:sublist/2).               */
int64_t evo_search(const double *rows, int64_t n, int d, const double *q,
                   int64_t k, int metric, const uint64_t *ranks,
                   int64_t *out_rows, double *out_dist) {
    if (k < 0) return -1;
    evo_cand *c = (evo_cand *)malloc(sizeof(evo_cand) * (size_t)(n > 0 ? n : 1));
    if (!c) return -2;
    for (int64_t r = 0; r < n; ++r) {
        c[r].dist = evo_distance(q, rows + (size_t)r * d, d, metric);
        c[r].rank = ranks ? ranks[r] : (uint64_t)r;
        c[r].row = r;
    }
    qsort(c, (size_t)n, sizeof(evo_cand), evo_cand_cmp); /* lists:sort/1 */
    int64_t m = k < n ? k : n;                           /* lists:sublist/2 */
    for (int64_t i = 0; i < m; ++i) {
        out_rows[i] = c[i].row;
        out_dist[i] = c[i].dist;
    }
    free(c);
    return m;
}

/* Same, over an fp32-resident corpus widened to fp64 row by row (exact).
 * Used by the CPU baseline so that 1M x 768 fits host memory as 3 GB.       */
int64_t evo_search_f32(const float *rows, int64_t n, int d, const double *q,
                       int64_t k, int metric, int64_t *out_rows,
                       double *out_dist) {
    if (k < 0) return -1;
    evo_cand *c = (evo_cand *)malloc(sizeof(evo_cand) * (size_t)(n > 0 ? n : 1));
    double *v = (double *)malloc(sizeof(double) * (size_t)(d > 0 ? d : 1));
    if (!c || !v) { free(c); free(v); return -2; }
    for (int64_t r = 0; r < n; ++r) {
        const float *src = rows + (size_t)r * d;
        for (int i = 0; i < d; ++i) v[i] = (double)src[i];
        c[r].dist = evo_distance(q, v, d, metric);
        c[r].rank = (uint64_t)r;
        c[r].row = r;
    }
    qsort(c, (size_t)n, sizeof(evo_cand), evo_cand_cmp);
    int64_t m = k < n ? k : n;
    for (int64_t i = 0; i < m; ++i) {
        out_rows[i] = c[i].row;
        out_dist[i] = c[i].dist;
    }
    free(c);
    free(v);
    return m;
}

/* All distances, no selection (for tolerance-aware parity checks). */
void evo_distances(const double *rows, int64_t n, int d, const double *q,
                   int metric, double *out) {
    for (int64_t r = 0; r < n; ++r)
        out[r] = evo_distance(q, rows + (size_t)r * d, d, metric);
}

/* ---- src/vector_compression.erl:306-309 find_min_max/1 ------------------- */
static void evo_min_max(const double *v, int d, double *mn, double *mx) {
    double lo = v[0], hi = v[0];
    for (int i = 1; i < d; ++i) {
        if (v[i] < lo) lo = v[i];
        if (v[i] > hi) hi = v[i];
    }
    *mn = lo;
    *mx = hi;
}

/* ---- src/vector_compression.erl:167-178 compress_8bit_quantization/1 -----
 * Scale = (Max-Min)/255.0; code = round((V-Min)/Scale).  Max == Min divides
 * by 0.0 => badarith in the reference => return -1 (caller stores raw).     */
int evo_quantize_8bit(const double *v, int d, uint8_t *codes, double *mn,
                      double *mx, double *scale) {
    if (d <= 0) return -1;
    evo_min_max(v, d, mn, mx);
    *scale = (*mx - *mn) / 255.0;
    if (*scale == 0.0) return -1;
    for (int i = 0; i < d; ++i) {
        double t = (v[i] - *mn) / *scale;
        double r = round(t);
        if (r < 0.0 || r > 255.0) return -1; /* list_to_binary badarg */
        codes[i] = (uint8_t)r;
    }
    return 0;
}

/* ---- src/vector_compression.erl:180-183 ---------------------------------- */
void evo_dequantize_8bit(const uint8_t *codes, int d, double mn, double scale,
                         double *out) {
    for (int i = 0; i < d; ++i) out[i] = mn + ((double)codes[i] * scale);
}

/* ---- src/vector_compression.erl:186-199 + pack_4bit_values :311-319 ------
 * /15.0; two codes per byte, FIRST element in the HIGH nibble; odd tail is
 * padded with a zero low nibble.  packed must hold (d+1)/2 bytes.           */
int evo_quantize_4bit(const double *v, int d, uint8_t *packed, double *mn,
                      double *mx, double *scale) {
    if (d <= 0) return -1;
    evo_min_max(v, d, mn, mx);
    *scale = (*mx - *mn) / 15.0;
    if (*scale == 0.0) return -1;
    memset(packed, 0, (size_t)(d + 1) / 2);
    for (int i = 0; i < d; ++i) {
        double t = (v[i] - *mn) / *scale;
        double r = round(t);
        if (r < 0.0 || r > 15.0) return -1;
        uint8_t c = (uint8_t)r;
        if ((i & 1) == 0) packed[i >> 1] |= (uint8_t)(c << 4);
        else packed[i >> 1] |= c;
    }
    return 0;
}

/* ---- src/vector_compression.erl:201-204 + unpack_4bit_values :321-329 ---- */
void evo_dequantize_4bit(const uint8_t *packed, int d, double mn, double scale,
                         double *out) {
    for (int i = 0; i < d; ++i) {
        uint8_t b = packed[i >> 1];
        uint8_t c = (i & 1) == 0 ? (uint8_t)(b >> 4) : (uint8_t)(b & 0x0F);
        out[i] = mn + ((double)c * scale);
    }
}

/* ---- counter-based synthetic generator --------------------
 * value(seed, idx) = ((splitmix64(seed ^ idx) >> 40) - 2^23) * 2^-23, a
 * 24-bit grid in [-1,1): exact in fp32 and fp64.  idx = row*d + col.        */
static inline uint64_t evo_mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

double evo_synth_value(uint64_t seed, uint64_t idx) {
    uint64_t u = evo_mix64(seed ^ idx) >> 40;
    return ((double)((int64_t)u - 8388608)) * (1.0 / 8388608.0);
}

void evo_synth_fill_f64(uint64_t seed, uint64_t row0, int64_t nrows, int d,
                        double *out) {
    for (int64_t r = 0; r < nrows; ++r)
        for (int c = 0; c < d; ++c)
            out[(size_t)r * d + c] =
                evo_synth_value(seed, (row0 + (uint64_t)r) * (uint64_t)d + (uint64_t)c);
}

void evo_synth_fill_f32(uint64_t seed, uint64_t row0, int64_t nrows, int d,
                        float *out) {
    for (int64_t r = 0; r < nrows; ++r)
        for (int c = 0; c < d; ++c)
            out[(size_t)r * d + c] = (float)evo_synth_value(
                seed, (row0 + (uint64_t)r) * (uint64_t)d + (uint64_t)c);
}

/* ---- CPU baseline driver --------------------------------------------------
 * The reference serialises one store behind one gen_server
 * (src/vector_store.erl:143-150); its only parallelism is across stores.
 * T threads == T independent store replicas, each answering its own query
 * with the full reference algorithm over the same read-only corpus.         */
typedef struct {
    const float *rows;
    int64_t n;
    int d;
    const double *q;
    int64_t k;
    int metric;
    int64_t *out_rows;
    double *out_dist;
    int64_t rc;
} evo_job;

static void *evo_job_main(void *p) {
    evo_job *j = (evo_job *)p;
    j->rc = evo_search_f32(j->rows, j->n, j->d, j->q, j->k, j->metric,
                           j->out_rows, j->out_dist);
    return NULL;
}

/* queries: nq x d fp64; one thread per query (nq == number of threads).     */
int evo_search_f32_replicas(const float *rows, int64_t n, int d,
                            const double *queries, int nq, int64_t k,
                            int metric, int64_t *out_rows, double *out_dist) {
    if (nq <= 0) return 0;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nq);
    evo_job *jobs = (evo_job *)malloc(sizeof(evo_job) * (size_t)nq);
    for (int i = 0; i < nq; ++i) {
        jobs[i].rows = rows;
        jobs[i].n = n;
        jobs[i].d = d;
        jobs[i].q = queries + (size_t)i * d;
        jobs[i].k = k;
        jobs[i].metric = metric;
        jobs[i].out_rows = out_rows + (size_t)i * k;
        jobs[i].out_dist = out_dist + (size_t)i * k;
        jobs[i].rc = 0;
        pthread_create(&th[i], NULL, evo_job_main, &jobs[i]);
    }
    int bad = 0;
    for (int i = 0; i < nq; ++i) {
        pthread_join(th[i], NULL);
        if (jobs[i].rc < 0) bad = 1;
    }
    free(th);
    free(jobs);
    return bad ? -1 : 0;
}

typedef struct {
    uint64_t seed;
    uint64_t row0;
    int64_t nrows;
    int d;
    float *out;
} evo_fill_job;

static void *evo_fill_main(void *p) {
    evo_fill_job *j = (evo_fill_job *)p;
    evo_synth_fill_f32(j->seed, j->row0, j->nrows, j->d, j->out);
    return NULL;
}

/* Threaded fill (setup only; never inside a timed region). */
void evo_synth_fill_f32_mt(uint64_t seed, int64_t nrows, int d, float *out,
                           int threads) {
    if (threads < 1) threads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    evo_fill_job *jobs = (evo_fill_job *)malloc(sizeof(evo_fill_job) * (size_t)threads);
    int64_t per = (nrows + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        int64_t r0 = per * t;
        int64_t cnt = nrows - r0 < per ? nrows - r0 : per;
        if (cnt < 0) cnt = 0;
        jobs[t].seed = seed;
        jobs[t].row0 = (uint64_t)r0;
        jobs[t].nrows = cnt;
        jobs[t].d = d;
        jobs[t].out = out + (size_t)r0 * d;
        pthread_create(&th[t], NULL, evo_fill_main, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
    free(th);
    free(jobs);
}

/*
 * evdb_nif.c -- Erlang NIF shim over the C ABI of libevdb_b200 (include/evdb.h).
 *
 * This is the reference-side binding a maintainer adds to ErlVectorDB: the
 * vector_store gen_server (erlang/src/vector_store.erl in this repo, replacing
 * reference src/vector_store.erl) keeps Id <-> slot maps and metadata in its
 * state and calls these NIFs for everything that touches vectors.
 *
 * It CANNOT be compiled or loaded in this image (no Erlang/OTP, no erl_nif.h);
 * `gcc -fsyntax-only -I. -DEVDB_NIF_SYNTAX_CHECK evdb_nif.c` checks it against
 * the minimal declarations in erl_nif_stub.h.  All logic lives below the C ABI,
 * which is what the test-suite exercises.
 *
 * Scheduling: every call that can block on the GPU (bulk_load, upsert, search*)
 * is flagged ERL_NIF_DIRTY_JOB_CPU_BOUND so it runs on a dirty scheduler and
 * never stalls a normal BEAM scheduler thread.  The store handle is a NIF
 * resource: when the owning gen_server dies, the destructor frees device memory
 * (the equivalent of terminate/2, reference src/vector_store.erl:201-207).
 */
#ifdef EVDB_NIF_SYNTAX_CHECK
#include "erl_nif_stub.h"
#else
#include <erl_nif.h>
#endif
#include <stdlib.h>
#include <string.h>

#include "../../include/evdb.h"

static ErlNifResourceType *STORE_RT;
static ERL_NIF_TERM a_ok, a_error, a_none, a_dimension_mismatch, a_invalid_vector_format, a_cosine,
    a_euclidean, a_manhattan, a_enomem;

typedef struct { evdb_store *s; } store_res;

static void store_dtor(ErlNifEnv *env, void *obj) {
    (void)env;
    store_res *r = (store_res *)obj;
    if (r->s) evdb_store_destroy(r->s);
    r->s = NULL;
}

static ERL_NIF_TERM mk_error(ErlNifEnv *env, int rc) {
    if (rc == EVDB_E_DIM_MISMATCH) return enif_make_tuple2(env, a_error, a_dimension_mismatch);
    if (rc == EVDB_E_BAD_VECTOR) return enif_make_tuple2(env, a_error, a_invalid_vector_format);
    return enif_make_tuple2(env, a_error, enif_make_atom(env, evdb_strerror(rc)));
}

/* [number()] -> malloc'ed doubles; ints are accepted like is_number/1 does
 * (reference validate_vector/2, src/vector_store.erl:213-225).               */
static int list_to_doubles(ErlNifEnv *env, ERL_NIF_TERM list, double **out, unsigned *n) {
    unsigned len;
    if (!enif_get_list_length(env, list, &len)) return 0;
    double *v = (double *)malloc(sizeof(double) * (len ? len : 1));
    if (!v) return 0;
    ERL_NIF_TERM head, tail = list;
    for (unsigned i = 0; i < len; ++i) {
        double d;
        ErlNifSInt64 i64;
        if (!enif_get_list_cell(env, tail, &head, &tail)) { free(v); return 0; }
        if (enif_get_double(env, head, &d)) v[i] = d;
        else if (enif_get_int64(env, head, &i64)) v[i] = (double)i64;
        else if (enif_is_number(env, head)) {
            /* a bignum: is_number/1 accepts it (reference :213-225); its magnitude is >= 2^63, which as a
             * vector element can only overflow the arithmetic (the reference store would crash with
             * badarith) -- passed on as an infinity, which the library rejects as invalid_vector_format */
            v[i] = 1.0 / 0.0;
        }
        else { free(v); return 0; }
    }
    *out = v;
    *n = len;
    return 1;
}

static int get_metric(ErlNifEnv *env, ERL_NIF_TERM t, int *m) {
    (void)env;
    if (enif_is_identical(t, a_cosine)) *m = EVDB_COSINE;
    else if (enif_is_identical(t, a_euclidean)) *m = EVDB_EUCLIDEAN;
    else if (enif_is_identical(t, a_manhattan)) *m = EVDB_MANHATTAN;
    else return 0;
    return 1;
}

/* new([Device, ...], Dtype 0..3, GemmShadow 0|1) -> {ok, Ref}
 * one ordinal = a single-device store; several = ONE store behind one handle on those GPUs
 * (evdb_opts.n_shards: the BEAM is one OS process, reference src/vector_store.erl:38-39)      */
static ERL_NIF_TERM nif_new(ErlNifEnv *env, int argc, const ERL_NIF_TERM argv[]) {
    int dtype = EVDB_F32, shadow = 1;
    unsigned ndev = 0;
    if (argc != 3 || !enif_get_list_length(env, argv[0], &ndev) || ndev < 1 || ndev > EVDB_MAX_SHARDS ||
        !enif_get_int(env, argv[1], &dtype) || !enif_get_int(env, argv[2], &shadow))
        return enif_make_badarg(env);
    evdb_opts o;
    memset(&o, 0, sizeof(o));
    ERL_NIF_TERM head, tail = argv[0];
    for (unsigned i = 0; i < ndev; ++i) {
        int dv;
        if (!enif_get_list_cell(env, tail, &head, &tail) || !enif_get_int(env, head, &dv)) return enif_make_badarg(env);
        o.devices[i] = dv;
    }
    o.device = o.devices[0];
    o.n_shards = ndev > 1 ? (int)ndev : 0;
    o.dtype = dtype;
    o.gemm_shadow = shadow;
    evdb_store *s = NULL;
    int rc = evdb_store_create(&o, &s);
    if (rc != EVDB_OK) return mk_error(env, rc);
    store_res *r = (store_res *)enif_alloc_resource(STORE_RT, sizeof(store_res));
    r->s = s;
    ERL_NIF_TERM ref = enif_make_resource(env, r);
    enif_release_resource(r);
    return enif_make_tuple2(env, a_ok, ref);
}

/* upsert(Ref, Slot, [number()]) -> ok | {error, dimension_mismatch | invalid_vector_format} */
static ERL_NIF_TERM nif_upsert(ErlNifEnv *env, int argc, const ERL_NIF_TERM argv[]) {
    store_res *r;
    unsigned slot, n;
    double *v;
    if (argc != 3 || !enif_get_resource(env, argv[0], STORE_RT, (void **)&r) ||
        !enif_get_uint(env, argv[1], &slot))
        return enif_make_badarg(env);
    if (!list_to_doubles(env, argv[2], &v, &n)) return enif_make_tuple2(env, a_error, a_invalid_vector_format);
    int rc = evdb_store_upsert_f64(r->s, slot, v, (int)n);
    free(v);
    return rc == EVDB_OK ? a_ok : mk_error(env, rc);
}

/* bulk_load(Ref, <<F:32/float-native,...>>, N, D) -> ok   (vector_store:init/1) */
static ERL_NIF_TERM nif_bulk_load(ErlNifEnv *env, int argc, const ERL_NIF_TERM argv[]) {
    store_res *r;
    ErlNifBinary bin;
    ErlNifUInt64 n;
    int d;
    if (argc != 4 || !enif_get_resource(env, argv[0], STORE_RT, (void **)&r) ||
        !enif_inspect_binary(env, argv[1], &bin) || !enif_get_uint64(env, argv[2], &n) ||
        !enif_get_int(env, argv[3], &d) || bin.size != n * (size_t)d * sizeof(float))
        return enif_make_badarg(env);
    int rc = evdb_store_bulk_load_f32(r->s, (const float *)bin.data, n, d);
    return rc == EVDB_OK ? a_ok : mk_error(env, rc);
}

/* append(Ref, <<F:64/float-native,...>>, N, D) -> {ok, FirstSlot}
 * N new ids in one call (a run of handle_call({insert,..}) with fresh keys)     */
static ERL_NIF_TERM nif_append(ErlNifEnv *env, int argc, const ERL_NIF_TERM argv[]) {
    store_res *r;
    ErlNifBinary bin;
    ErlNifUInt64 n;
    int d;
    if (argc != 4 || !enif_get_resource(env, argv[0], STORE_RT, (void **)&r) ||
        !enif_inspect_binary(env, argv[1], &bin) || !enif_get_uint64(env, argv[2], &n) ||
        !enif_get_int(env, argv[3], &d) || bin.size != n * (size_t)d * sizeof(double))
        return enif_make_badarg(env);
    uint64_t first = 0;
    int rc = evdb_store_append_f64(r->s, (const double *)bin.data, n, d, &first);
    return rc == EVDB_OK ? enif_make_tuple2(env, a_ok, enif_make_uint64(env, first)) : mk_error(env, rc);
}

/* bulk_load_codes(Ref, CodesBin, MinsBin(f64), ScalesBin(f64), N, D) -> ok
 * compressed records of vector_persistence straight to device code columns    */
static ERL_NIF_TERM nif_bulk_load_codes(ErlNifEnv *env, int argc, const ERL_NIF_TERM argv[]) {
    store_res *r;
    ErlNifBinary codes, mins, scales;
    ErlNifUInt64 n;
    int d;
    if (argc != 6 || !enif_get_resource(env, argv[0], STORE_RT, (void **)&r) ||
        !enif_inspect_binary(env, argv[1], &codes) || !enif_inspect_binary(env, argv[2], &mins) ||
        !enif_inspect_binary(env, argv[3], &scales) || !enif_get_uint64(env, argv[4], &n) ||
        !enif_get_int(env, argv[5], &d) || d <= 0 || mins.size != n * sizeof(double) || scales.size != n * sizeof(double))
        return enif_make_badarg(env);
    {   /* the code binary must hold n rows of d bytes (8-bit) or (d+1)/2 bytes (4-bit): a short one would be read past its end */
        evdb_stats st;
        if (evdb_store_stats(r->s, &st) != EVDB_OK) return enif_make_badarg(env);
        const size_t row = st.dtype == EVDB_U8 ? (size_t)d : ((size_t)d + 1) / 2;
        if ((st.dtype != EVDB_U8 && st.dtype != EVDB_U4) || codes.size != n * row) return enif_make_badarg(env);
    }
    int rc = evdb_store_bulk_load_codes(r->s, codes.data, (const double *)mins.data,
                                        (const double *)scales.data, n, d);
    return rc == EVDB_OK ? a_ok : mk_error(env, rc);
}

/* delete(Ref, Slot) -> {ok, MovedFromSlot | none} */
static ERL_NIF_TERM nif_delete(ErlNifEnv *env, int argc, const ERL_NIF_TERM argv[]) {
    store_res *r;
    unsigned slot;
    if (argc != 2 || !enif_get_resource(env, argv[0], STORE_RT, (void **)&r) ||
        !enif_get_uint(env, argv[1], &slot))
        return enif_make_badarg(env);
    int64_t moved = -1;
    int rc = evdb_store_delete(r->s, slot, &moved);
    if (rc != EVDB_OK) return mk_error(env, rc);
    return enif_make_tuple2(env, a_ok, moved < 0 ? a_none : enif_make_int64(env, moved));
}

/* K > N returns all N rows (lists:sublist/2): the result buffers never need more than count places */
static int clamp_k(evdb_store *s, int k) {
    evdb_stats st;
    if (evdb_store_stats(s, &st) != EVDB_OK) return 0;
    return (uint64_t)k > st.count ? (int)st.count : k;
}

/* search(Ref, [number()], K, Metric) -> {ok, [{Distance, Slot}]} ascending by (Distance, Slot) */
static ERL_NIF_TERM nif_search(ErlNifEnv *env, int argc, const ERL_NIF_TERM argv[]) {
    store_res *r;
    unsigned n;
    int k, metric;
    double *q;
    if (argc != 4 || !enif_get_resource(env, argv[0], STORE_RT, (void **)&r) ||
        !enif_get_int(env, argv[2], &k) || k < 0 || !get_metric(env, argv[3], &metric))
        return enif_make_badarg(env);
    if (!list_to_doubles(env, argv[1], &q, &n)) return enif_make_tuple2(env, a_error, a_invalid_vector_format);
    k = clamp_k(r->s, k);   /* lists:sublist/2 tolerates any K: never size buffers from the caller's number */
    uint32_t *slots = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(k ? k : 1));
    double *dists = (double *)malloc(sizeof(double) * (size_t)(k ? k : 1));
    if (!slots || !dists) { free(q); free(slots); free(dists); return enif_make_tuple2(env, a_error, a_enomem); }
    int32_t count = 0;
    int rc = evdb_store_search_f64(r->s, q, 1, (int)n, k, metric, slots, dists, &count);
    ERL_NIF_TERM out;
    if (rc != EVDB_OK) {
        out = mk_error(env, rc);
    } else {
        ERL_NIF_TERM list = enif_make_list(env, 0);
        for (int i = count - 1; i >= 0; --i)
            list = enif_make_list_cell(env, enif_make_tuple2(env, enif_make_double(env, dists[i]),
                                                             enif_make_uint(env, slots[i])), list);
        out = enif_make_tuple2(env, a_ok, list);
    }
    free(q); free(slots); free(dists);
    return out;
}

/* search_batch(Ref, <<Q:64/float-native,...>>, B, D, K, Metric) -> {ok, [[{Distance, Slot}]]} */
static ERL_NIF_TERM nif_search_batch(ErlNifEnv *env, int argc, const ERL_NIF_TERM argv[]) {
    store_res *r;
    ErlNifBinary qb;
    int B, d, k, metric;
    if (argc != 6 || !enif_get_resource(env, argv[0], STORE_RT, (void **)&r) ||
        !enif_inspect_binary(env, argv[1], &qb) || !enif_get_int(env, argv[2], &B) ||
        !enif_get_int(env, argv[3], &d) || !enif_get_int(env, argv[4], &k) || k < 0 || B < 0 ||
        !get_metric(env, argv[5], &metric) || qb.size != (size_t)B * (size_t)d * sizeof(double))
        return enif_make_badarg(env);
    k = clamp_k(r->s, k);
    size_t nk = (size_t)B * (size_t)(k ? k : 1);
    uint32_t *slots = (uint32_t *)malloc(sizeof(uint32_t) * (nk ? nk : 1));
    double *dists = (double *)malloc(sizeof(double) * (nk ? nk : 1));
    int32_t *counts = (int32_t *)calloc((size_t)(B ? B : 1), sizeof(int32_t));
    if (!slots || !dists || !counts) { free(slots); free(dists); free(counts); return enif_make_tuple2(env, a_error, a_enomem); }
    int rc = evdb_store_search_f64(r->s, (const double *)qb.data, B, d, k, metric, slots, dists, counts);
    ERL_NIF_TERM out;
    if (rc != EVDB_OK) {
        out = mk_error(env, rc);
    } else {
        ERL_NIF_TERM outer = enif_make_list(env, 0);
        for (int b = B - 1; b >= 0; --b) {
            ERL_NIF_TERM list = enif_make_list(env, 0);
            for (int i = counts[b] - 1; i >= 0; --i)
                list = enif_make_list_cell(env, enif_make_tuple2(env, enif_make_double(env, dists[(size_t)b * k + i]),
                                                                 enif_make_uint(env, slots[(size_t)b * k + i])), list);
            outer = enif_make_list_cell(env, list, outer);
        }
        out = enif_make_tuple2(env, a_ok, outer);
    }
    free(slots); free(dists); free(counts);
    return out;
}

/* get(Ref, Slot, D) -> {ok, [float()]}   (get_all_vectors support) */
static ERL_NIF_TERM nif_get(ErlNifEnv *env, int argc, const ERL_NIF_TERM argv[]) {
    store_res *r;
    unsigned slot;
    int d;
    if (argc != 3 || !enif_get_resource(env, argv[0], STORE_RT, (void **)&r) ||
        !enif_get_uint(env, argv[1], &slot) || !enif_get_int(env, argv[2], &d) || d <= 0)
        return enif_make_badarg(env);
    double *v = (double *)malloc(sizeof(double) * (size_t)d);
    if (!v) return enif_make_tuple2(env, a_error, a_enomem);
    int rc = evdb_store_get_f64(r->s, slot, v, d);
    ERL_NIF_TERM out;
    if (rc != EVDB_OK) {
        out = mk_error(env, rc);
    } else {
        ERL_NIF_TERM list = enif_make_list(env, 0);
        for (int i = d - 1; i >= 0; --i) list = enif_make_list_cell(env, enif_make_double(env, v[i]), list);
        out = enif_make_tuple2(env, a_ok, list);
    }
    free(v);
    return out;
}

/* stats(Ref) -> {ok, {Count, Dimension, DeviceBytes, Searches}} */
static ERL_NIF_TERM nif_stats(ErlNifEnv *env, int argc, const ERL_NIF_TERM argv[]) {
    store_res *r;
    if (argc != 1 || !enif_get_resource(env, argv[0], STORE_RT, (void **)&r)) return enif_make_badarg(env);
    evdb_stats st;
    int rc = evdb_store_stats(r->s, &st);
    if (rc != EVDB_OK) return mk_error(env, rc);
    return enif_make_tuple2(env, a_ok,
                            enif_make_tuple4(env, enif_make_uint64(env, st.count), enif_make_int(env, st.dimension),
                                             enif_make_uint64(env, st.device_bytes), enif_make_uint64(env, st.searches)));
}

static int load(ErlNifEnv *env, void **priv, ERL_NIF_TERM info) {
    (void)priv; (void)info;
    STORE_RT = enif_open_resource_type(env, NULL, "evdb_store", store_dtor, ERL_NIF_RT_CREATE, NULL);
    if (!STORE_RT) return -1;
    a_ok = enif_make_atom(env, "ok");
    a_error = enif_make_atom(env, "error");
    a_none = enif_make_atom(env, "none");
    a_dimension_mismatch = enif_make_atom(env, "dimension_mismatch");
    a_invalid_vector_format = enif_make_atom(env, "invalid_vector_format");
    a_cosine = enif_make_atom(env, "cosine");
    a_euclidean = enif_make_atom(env, "euclidean");
    a_manhattan = enif_make_atom(env, "manhattan");
    a_enomem = enif_make_atom(env, "enomem");
    return evdb_init(NULL, 0) == EVDB_OK ? 0 : -1;  /* no GPU -> the NIF refuses to load: no CPU fallback */
}

static ErlNifFunc nif_funcs[] = {
    {"new", 3, nif_new, 0},
    {"upsert", 3, nif_upsert, ERL_NIF_DIRTY_JOB_CPU_BOUND},
    {"bulk_load", 4, nif_bulk_load, ERL_NIF_DIRTY_JOB_CPU_BOUND},
    {"append", 4, nif_append, ERL_NIF_DIRTY_JOB_CPU_BOUND},
    {"bulk_load_codes", 6, nif_bulk_load_codes, ERL_NIF_DIRTY_JOB_CPU_BOUND},
    {"delete", 2, nif_delete, ERL_NIF_DIRTY_JOB_CPU_BOUND},
    {"search", 4, nif_search, ERL_NIF_DIRTY_JOB_CPU_BOUND},
    {"search_batch", 6, nif_search_batch, ERL_NIF_DIRTY_JOB_CPU_BOUND},
    {"get", 3, nif_get, ERL_NIF_DIRTY_JOB_CPU_BOUND},
    {"stats", 1, nif_stats, 0},
};

ERL_NIF_INIT(evdb_nif, nif_funcs, load, NULL, NULL, NULL)

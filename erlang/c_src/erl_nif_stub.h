/* erl_nif_stub.h -- the SUBSET of erl_nif.h declarations evdb_nif.c uses, for
 * `gcc -fsyntax-only -DEVDB_NIF_SYNTAX_CHECK` in an image without Erlang/OTP.
 * Not a substitute for the real header: build against OTP's erl_nif.h. */
#ifndef ERL_NIF_STUB_H
#define ERL_NIF_STUB_H
#include <stddef.h>
#include <stdint.h>
typedef uintptr_t ERL_NIF_TERM;
typedef struct enif_environment_t ErlNifEnv;
typedef struct enif_resource_type_t ErlNifResourceType;
typedef int64_t ErlNifSInt64;
typedef uint64_t ErlNifUInt64;
typedef struct { size_t size; unsigned char *data; void *ref_bin; void *spare[2]; } ErlNifBinary;
typedef struct { const char *name; unsigned arity; ERL_NIF_TERM (*fptr)(ErlNifEnv *, int, const ERL_NIF_TERM[]); unsigned flags; } ErlNifFunc;
typedef void ErlNifResourceDtor(ErlNifEnv *, void *);
typedef enum { ERL_NIF_RT_CREATE = 1, ERL_NIF_RT_TAKEOVER = 2 } ErlNifResourceFlags;
#define ERL_NIF_DIRTY_JOB_CPU_BOUND 1
ErlNifResourceType *enif_open_resource_type(ErlNifEnv *, const char *, const char *, ErlNifResourceDtor *, ErlNifResourceFlags, ErlNifResourceFlags *);
void *enif_alloc_resource(ErlNifResourceType *, size_t);
void enif_release_resource(void *);
ERL_NIF_TERM enif_make_resource(ErlNifEnv *, void *);
int enif_get_resource(ErlNifEnv *, ERL_NIF_TERM, ErlNifResourceType *, void **);
ERL_NIF_TERM enif_make_atom(ErlNifEnv *, const char *);
ERL_NIF_TERM enif_make_badarg(ErlNifEnv *);
ERL_NIF_TERM enif_make_tuple2(ErlNifEnv *, ERL_NIF_TERM, ERL_NIF_TERM);
ERL_NIF_TERM enif_make_tuple4(ErlNifEnv *, ERL_NIF_TERM, ERL_NIF_TERM, ERL_NIF_TERM, ERL_NIF_TERM);
ERL_NIF_TERM enif_make_list(ErlNifEnv *, unsigned, ...);
ERL_NIF_TERM enif_make_list_cell(ErlNifEnv *, ERL_NIF_TERM, ERL_NIF_TERM);
ERL_NIF_TERM enif_make_double(ErlNifEnv *, double);
ERL_NIF_TERM enif_make_int(ErlNifEnv *, int);
ERL_NIF_TERM enif_make_uint(ErlNifEnv *, unsigned);
ERL_NIF_TERM enif_make_int64(ErlNifEnv *, ErlNifSInt64);
ERL_NIF_TERM enif_make_uint64(ErlNifEnv *, ErlNifUInt64);
int enif_get_list_length(ErlNifEnv *, ERL_NIF_TERM, unsigned *);
int enif_get_list_cell(ErlNifEnv *, ERL_NIF_TERM, ERL_NIF_TERM *, ERL_NIF_TERM *);
int enif_get_double(ErlNifEnv *, ERL_NIF_TERM, double *);
int enif_get_int(ErlNifEnv *, ERL_NIF_TERM, int *);
int enif_get_uint(ErlNifEnv *, ERL_NIF_TERM, unsigned *);
int enif_get_int64(ErlNifEnv *, ERL_NIF_TERM, ErlNifSInt64 *);
int enif_get_uint64(ErlNifEnv *, ERL_NIF_TERM, ErlNifUInt64 *);
int enif_inspect_binary(ErlNifEnv *, ERL_NIF_TERM, ErlNifBinary *);
int enif_is_identical(ERL_NIF_TERM, ERL_NIF_TERM);
int enif_is_number(ErlNifEnv *, ERL_NIF_TERM);
#define ERL_NIF_INIT(NAME, FUNCS, LOAD, RELOAD, UPGRADE, UNLOAD) \
    const ErlNifFunc *evdb_nif_init_##NAME(void) { (void)(LOAD); return (FUNCS); }
#endif

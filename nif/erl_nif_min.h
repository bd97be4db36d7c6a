/* Minimal declarations of the erl_nif API subset orbx_nif.c uses, ONLY so that the wrapper can be compile-checked in an
 * image without Erlang/OTP.  Signatures follow erts/emulator/beam/erl_nif.h (OTP 26). */
#ifndef ORBX_ERL_NIF_MIN_H
#define ORBX_ERL_NIF_MIN_H
#include <stddef.h>
#include <stdint.h>
typedef unsigned long ERL_NIF_TERM;
typedef struct enif_environment_t ErlNifEnv;
typedef struct enif_resource_type_t ErlNifResourceType;
typedef struct { size_t size; unsigned char *data; void *ref_bin; void *spare[2]; } ErlNifBinary;
typedef struct { const char *name; unsigned arity; ERL_NIF_TERM (*fptr)(ErlNifEnv *, int, const ERL_NIF_TERM[]); unsigned flags; } ErlNifFunc;
typedef void ErlNifResourceDtor(ErlNifEnv *, void *);
typedef struct ErlNifMutex_ ErlNifMutex;
typedef enum { ERL_NIF_RT_CREATE = 1, ERL_NIF_RT_TAKEOVER = 2 } ErlNifResourceFlags;
#define ERL_NIF_DIRTY_JOB_IO_BOUND 2
int enif_get_int(ErlNifEnv *, ERL_NIF_TERM, int *);
int enif_get_long(ErlNifEnv *, ERL_NIF_TERM, long *);
int enif_get_double(ErlNifEnv *, ERL_NIF_TERM, double *);
void *enif_alloc(size_t);
void enif_free(void *);
ErlNifMutex *enif_mutex_create(char *name);
void enif_mutex_destroy(ErlNifMutex *);
int enif_mutex_trylock(ErlNifMutex *);   /* 0 = acquired, EBUSY otherwise */
void enif_mutex_lock(ErlNifMutex *);
void enif_mutex_unlock(ErlNifMutex *);
int enif_get_tuple(ErlNifEnv *, ERL_NIF_TERM, int *, const ERL_NIF_TERM **);
int enif_inspect_binary(ErlNifEnv *, ERL_NIF_TERM, ErlNifBinary *);
int enif_get_resource(ErlNifEnv *, ERL_NIF_TERM, ErlNifResourceType *, void **);
void *enif_alloc_resource(ErlNifResourceType *, size_t);
void enif_release_resource(void *);
ERL_NIF_TERM enif_make_resource(ErlNifEnv *, void *);
ERL_NIF_TERM enif_make_atom(ErlNifEnv *, const char *);
ERL_NIF_TERM enif_make_int(ErlNifEnv *, int);
ERL_NIF_TERM enif_make_badarg(ErlNifEnv *);
ERL_NIF_TERM enif_make_tuple(ErlNifEnv *, unsigned, ...);
#define enif_make_tuple2(e, a, b) enif_make_tuple(e, 2, a, b)
#define enif_make_tuple3(e, a, b, c) enif_make_tuple(e, 3, a, b, c)
#define enif_make_tuple5(e, a, b, c, d, f) enif_make_tuple(e, 5, a, b, c, d, f)
#define enif_make_tuple6(e, a, b, c, d, f, g) enif_make_tuple(e, 6, a, b, c, d, f, g)
#define enif_make_tuple7(e, a, b, c, d, f, g, i) enif_make_tuple(e, 7, a, b, c, d, f, g, i)
unsigned char *enif_make_new_binary(ErlNifEnv *, size_t, ERL_NIF_TERM *);
ERL_NIF_TERM enif_make_sub_binary(ErlNifEnv *, ERL_NIF_TERM, size_t, size_t);
ErlNifResourceType *enif_open_resource_type(ErlNifEnv *, const char *, const char *, ErlNifResourceDtor *, ErlNifResourceFlags, ErlNifResourceFlags *);
/* the function table ends with a {NULL} sentinel in mock mode so that nif/mock_host.c can walk it */
#define ERL_NIF_INIT(MOD, FUNCS, LOAD, RELOAD, UPGRADE, UNLOAD) \
    const ErlNifFunc *orbx_nif_funcs_for_check(void) { return FUNCS; } \
    int orbx_nif_mock_load(void) { return LOAD(NULL, NULL, 0); }
#endif

/* Mock BEAM host for nif/orbx_nif.c: implements the erl_nif subset of erl_nif_min.h with a tiny term table, so that the NIF
 * entry points can be loaded and CALLED in an image without Erlang/OTP (SURVEY.md §8b seam b3: "mock-host harness").  It is a
 * test harness, not part of the product: tests/test_gpu_parity.py builds  orbx_nif.c + mock_host.c -> one shared object and
 * drives it through the mock_* functions below (ctypes).  Terms live in a per-thread arena that mock_reset() clears; that is
 * enough for call-at-a-time use, which is how a dirty-scheduler NIF call looks from the C side. */
#include <pthread.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "erl_nif_min.h"

typedef enum { T_INT = 1, T_DOUBLE, T_ATOM, T_BIN, T_TUPLE, T_RES, T_BADARG } kind_t;
typedef struct {
    kind_t kind;
    long i; double d; char atom[32];
    unsigned char *data; size_t size; int owns;          /* binaries (sub-binaries do not own) */
    ERL_NIF_TERM elems[8]; int arity;                    /* tuples */
    void *res;                                           /* resources */
} term_t;

#define MAX_TERMS 256
static __thread term_t g_terms[MAX_TERMS];
static __thread int g_nterms;
struct enif_resource_type_t { ErlNifResourceDtor *dtor; };
typedef struct { ErlNifResourceType *type; long refs; } res_hdr;

static ERL_NIF_TERM new_term(kind_t k) {
    if (g_nterms >= MAX_TERMS) abort();
    memset(&g_terms[g_nterms], 0, sizeof(term_t));
    g_terms[g_nterms].kind = k;
    return (ERL_NIF_TERM)(++g_nterms);                   /* 1-based handle */
}
static term_t *T(ERL_NIF_TERM t) { return (t >= 1 && (int)t <= g_nterms) ? &g_terms[t - 1] : NULL; }

/* ---- erl_nif API used by the NIF --------------------------------------------------------------------------------------- */
int enif_get_int(ErlNifEnv *e, ERL_NIF_TERM t, int *out) { (void)e; term_t *x = T(t); if (!x || x->kind != T_INT) return 0; *out = (int)x->i; return 1; }
int enif_get_long(ErlNifEnv *e, ERL_NIF_TERM t, long *out) { (void)e; term_t *x = T(t); if (!x || x->kind != T_INT) return 0; *out = x->i; return 1; }
void *enif_alloc(size_t n) { return malloc(n); }
void enif_free(void *p) { free(p); }
struct ErlNifMutex_ { pthread_mutex_t m; };
ErlNifMutex *enif_mutex_create(char *name) { (void)name; ErlNifMutex *m = (ErlNifMutex *)calloc(1, sizeof(*m)); if (m) pthread_mutex_init(&m->m, NULL); return m; }
void enif_mutex_destroy(ErlNifMutex *m) { if (m) { pthread_mutex_destroy(&m->m); free(m); } }
int enif_mutex_trylock(ErlNifMutex *m) { return pthread_mutex_trylock(&m->m); }
void enif_mutex_lock(ErlNifMutex *m) { pthread_mutex_lock(&m->m); }
void enif_mutex_unlock(ErlNifMutex *m) { pthread_mutex_unlock(&m->m); }
int enif_get_double(ErlNifEnv *e, ERL_NIF_TERM t, double *out) { (void)e; term_t *x = T(t); if (!x || x->kind != T_DOUBLE) return 0; *out = x->d; return 1; }
int enif_get_tuple(ErlNifEnv *e, ERL_NIF_TERM t, int *arity, const ERL_NIF_TERM **arr) {
    (void)e; term_t *x = T(t); if (!x || x->kind != T_TUPLE) return 0; *arity = x->arity; *arr = x->elems; return 1;
}
int enif_inspect_binary(ErlNifEnv *e, ERL_NIF_TERM t, ErlNifBinary *b) {
    (void)e; term_t *x = T(t); if (!x || x->kind != T_BIN) return 0; b->size = x->size; b->data = x->data; return 1;
}
int enif_get_resource(ErlNifEnv *e, ERL_NIF_TERM t, ErlNifResourceType *type, void **obj) {
    (void)e; term_t *x = T(t); if (!x || x->kind != T_RES) return 0;
    res_hdr *h = (res_hdr *)x->res - 1; if (h->type != type) return 0; *obj = x->res; return 1;
}
void *enif_alloc_resource(ErlNifResourceType *type, size_t size) {
    res_hdr *h = (res_hdr *)calloc(1, sizeof(res_hdr) + size); h->type = type; h->refs = 1; return h + 1;
}
static void res_unref(void *obj) {
    res_hdr *h = (res_hdr *)obj - 1;
    if (--h->refs == 0) { if (h->type->dtor) h->type->dtor(NULL, obj); free(h); }
}
void enif_release_resource(void *obj) { res_unref(obj); }
ERL_NIF_TERM enif_make_resource(ErlNifEnv *e, void *obj) { (void)e; ERL_NIF_TERM t = new_term(T_RES); T(t)->res = obj; ((res_hdr *)obj - 1)->refs++; return t; }
ERL_NIF_TERM enif_make_atom(ErlNifEnv *e, const char *name) { (void)e; ERL_NIF_TERM t = new_term(T_ATOM); strncpy(T(t)->atom, name, 31); return t; }
ERL_NIF_TERM enif_make_int(ErlNifEnv *e, int v) { (void)e; ERL_NIF_TERM t = new_term(T_INT); T(t)->i = v; return t; }
ERL_NIF_TERM enif_make_badarg(ErlNifEnv *e) { (void)e; return new_term(T_BADARG); }
ERL_NIF_TERM enif_make_tuple(ErlNifEnv *e, unsigned n, ...) {
    (void)e; ERL_NIF_TERM t = new_term(T_TUPLE); va_list ap; va_start(ap, n);
    T(t)->arity = (int)n; for (unsigned i = 0; i < n && i < 8; i++) T(t)->elems[i] = va_arg(ap, ERL_NIF_TERM);
    va_end(ap); return t;
}
unsigned char *enif_make_new_binary(ErlNifEnv *e, size_t size, ERL_NIF_TERM *out) {
    (void)e; ERL_NIF_TERM t = new_term(T_BIN); T(t)->data = (unsigned char *)malloc(size ? size : 1); T(t)->size = size; T(t)->owns = 1; *out = t; return T(t)->data;
}
ERL_NIF_TERM enif_make_sub_binary(ErlNifEnv *e, ERL_NIF_TERM bin, size_t pos, size_t size) {
    (void)e; term_t *b = T(bin); ERL_NIF_TERM t = new_term(T_BIN); T(t)->data = b->data + pos; T(t)->size = size; return t;
}
ErlNifResourceType *enif_open_resource_type(ErlNifEnv *e, const char *m, const char *n, ErlNifResourceDtor *dtor, ErlNifResourceFlags f, ErlNifResourceFlags *tried) {
    (void)e; (void)m; (void)n; (void)f; (void)tried;
    ErlNifResourceType *t = (ErlNifResourceType *)calloc(1, sizeof(*t)); t->dtor = dtor; return t;
}

/* ---- driver side (what the test calls) ------------------------------------------------------------------------------------ */
const ErlNifFunc *orbx_nif_funcs_for_check(void);
int orbx_nif_mock_load(void);                            /* defined by ERL_NIF_INIT in mock mode */

void mock_reset(void) {                                  /* end of a "call": free binaries, drop resource references of terms */
    for (int i = 0; i < g_nterms; i++) {
        if (g_terms[i].kind == T_BIN && g_terms[i].owns) free(g_terms[i].data);
        if (g_terms[i].kind == T_RES) res_unref(g_terms[i].res);
    }
    g_nterms = 0;
}
unsigned long mock_int(int v) { return enif_make_int(NULL, v); }
unsigned long mock_double(double v) { ERL_NIF_TERM t = new_term(T_DOUBLE); T(t)->d = v; return t; }
unsigned long mock_binary(const void *p, size_t n) { ERL_NIF_TERM t; unsigned char *d = enif_make_new_binary(NULL, n, &t); memcpy(d, p, n); return t; }
unsigned long mock_tuple4(unsigned long a, unsigned long b, unsigned long c, unsigned long d) { return enif_make_tuple(NULL, 4, a, b, c, d); }
/* a resource term that survives mock_reset: the test keeps the raw pointer (as the BEAM would keep the term alive) */
void *mock_resource_keep(unsigned long t) { term_t *x = T(t); if (!x || x->kind != T_RES) return NULL; ((res_hdr *)x->res - 1)->refs++; return x->res; }
unsigned long mock_resource_term(void *obj) { return enif_make_resource(NULL, obj); }
void mock_resource_drop(void *obj) { res_unref(obj); }
unsigned long mock_call(const char *name, int argc, const unsigned long *argv) {
    const ErlNifFunc *f = orbx_nif_funcs_for_check();
    for (int i = 0; i < 32 && f[i].name; i++)
        if (!strcmp(f[i].name, name) && (int)f[i].arity == argc) return f[i].fptr(NULL, argc, argv);
    return enif_make_badarg(NULL);
}
unsigned long mock_long(long v) { ERL_NIF_TERM t = new_term(T_INT); T(t)->i = v; return t; }
int mock_kind(unsigned long t) { term_t *x = T(t); return x ? (int)x->kind : 0; }
long mock_get_int(unsigned long t) { return T(t)->i; }
const char *mock_get_atom(unsigned long t) { return T(t)->atom; }
int mock_tuple_arity(unsigned long t) { return T(t)->arity; }
unsigned long mock_tuple_elem(unsigned long t, int i) { return T(t)->elems[i]; }
size_t mock_bin_size(unsigned long t) { return T(t)->size; }
const void *mock_bin_data(unsigned long t) { return T(t)->data; }

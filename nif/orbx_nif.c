/* erl_nif wrapper of liborbx.so for the Elixir side of SEND-SLAM (seam b3).
 *
 * The reference has no NIF: a process registered in CameraRegistry receives {:camera_frame, {:ok, opts}}
 * (send_slam/lib/send_slam/camera_producer.ex:190-208; consumers register like send_slam/lib/send_slam/timer.ex:24-27)
 * and SlamHandler ships every frame as PPM over TCP (send_slam/lib/send_slam/slam_handler.ex:59-88).  With this NIF a
 * consumer can extract in-VM: `SendSlam.OrbNif.extract(handle, Evision.Mat.to_binary(gray), w, h)`.
 * All calls block on the GPU for 100s of microseconds or more => ERL_NIF_DIRTY_JOB_IO_BOUND.  Errors come back as
 * {:error, reason}; nothing here raises or crashes the VM.
 * An orbx handle is single-flight (include/orbx.h), while a resource term can be handed to any number of BEAM processes: every
 * resource therefore carries an ErlNifMutex that each call takes with enif_mutex_trylock; a second process that calls into a
 * busy handle gets {:error, :busy} instead of racing on the handle's workspace, error string and CUDA-graph cache.  One process
 * should own a handle (one camera = one consumer = one handle); the lock is the safety net, not a scheduler.
 * Build (where Erlang is installed): cc -shared -fPIC -I$ERL_INCLUDE -I../include orbx_nif.c -L../send_slam_b200 -lorbx
 * Compile-check here (no erl_nif.h in this image): cc -DORBX_NIF_MIN -I../include -c orbx_nif.c
 */
#ifdef ORBX_NIF_MIN
#include "erl_nif_min.h"
#else
#include <erl_nif.h>
#endif
#include <string.h>

#include "orbx.h"

static ErlNifResourceType *g_handle_type, *g_db_type;

typedef struct { orbx_handle *h; int cap, max_batch; ErlNifMutex *lock; } nif_handle;
typedef struct { orbx_db *db; ErlNifMutex *lock; } nif_db;

static void handle_dtor(ErlNifEnv *env, void *obj) {
    (void)env;
    nif_handle *nh = (nif_handle *)obj;
    if (nh->h) orbx_destroy(nh->h);
    if (nh->lock) enif_mutex_destroy(nh->lock);
    nh->h = NULL; nh->lock = NULL;
}

static void db_dtor(ErlNifEnv *env, void *obj) {
    (void)env;
    nif_db *nd = (nif_db *)obj;
    if (nd->db) orbx_knn2_destroy_db(nd->db);
    if (nd->lock) enif_mutex_destroy(nd->lock);
    nd->db = NULL; nd->lock = NULL;
}

/* Erlang binaries start at arbitrary byte offsets: typed arrays are copied into aligned memory before the library reads them */
static void *aligned_copy(const ErlNifBinary *b) {
    void *p = enif_alloc(b->size ? b->size : 1);
    if (p && b->size) memcpy(p, b->data, b->size);
    return p;
}

static ERL_NIF_TERM mk_error(ErlNifEnv *env, const char *reason) {
    return enif_make_tuple2(env, enif_make_atom(env, "error"), enif_make_atom(env, reason));
}

static const char *code_atom(int rc) {
    switch (rc) {
        case ORBX_E_INVALID: return "invalid_argument";
        case ORBX_E_CUDA: return "cuda_error";
        case ORBX_E_CAPACITY: return "capacity";
        case ORBX_E_EMPTY: return "empty_image";
        case ORBX_E_OVERFLOW: return "overflow";
        default: return "unknown";
    }
}

/* create(nfeatures, scale_factor, nlevels, ini_th, min_th, device, max_width, max_height [, max_batch = 1]) */
static ERL_NIF_TERM nif_create(ErlNifEnv *env, int argc, const ERL_NIF_TERM argv[]) {
    int nf, nl, ini, mn, dev, mw, mh, mb = 1;
    double sf;
    if ((argc != 8 && argc != 9) || !enif_get_int(env, argv[0], &nf) || !enif_get_double(env, argv[1], &sf) || !enif_get_int(env, argv[2], &nl) ||
        !enif_get_int(env, argv[3], &ini) || !enif_get_int(env, argv[4], &mn) || !enif_get_int(env, argv[5], &dev) ||
        !enif_get_int(env, argv[6], &mw) || !enif_get_int(env, argv[7], &mh) || (argc == 9 && !enif_get_int(env, argv[8], &mb)))
        return enif_make_badarg(env);
    orbx_config cfg;
    cfg.nfeatures = nf; cfg.scale_factor = (float)sf; cfg.nlevels = nl; cfg.ini_th_fast = ini; cfg.min_th_fast = mn;
    cfg.device = dev; cfg.max_width = mw; cfg.max_height = mh; cfg.max_batch = mb;
    orbx_handle *h = NULL;
    int rc = orbx_create(&cfg, &h);
    if (rc != ORBX_OK) return mk_error(env, code_atom(rc));
    nif_handle *nh = (nif_handle *)enif_alloc_resource(g_handle_type, sizeof(nif_handle));
    if (!nh) { orbx_destroy(h); return mk_error(env, "enomem"); }
    nh->h = h; nh->cap = orbx_keypoint_capacity(h); nh->max_batch = mb;
    nh->lock = enif_mutex_create((char *)"orbx_handle");
    if (!nh->lock) { enif_release_resource(nh); return mk_error(env, "enomem"); }
    ERL_NIF_TERM term = enif_make_resource(env, nh);
    enif_release_resource(nh);
    return enif_make_tuple2(env, enif_make_atom(env, "ok"), term);
}

#define LOCK_OR_BUSY(nh) do { if (enif_mutex_trylock((nh)->lock) != 0) return mk_error(env, "busy"); } while (0)

/* extract(handle, gray_binary, width, height) -> {:ok, n, mono_index, keypoints_binary (n*28 B), descriptors_binary (n*32 B)} */
static ERL_NIF_TERM nif_extract(ErlNifEnv *env, int argc, const ERL_NIF_TERM argv[]) {
    nif_handle *nh;
    ErlNifBinary img;
    int w, h;
    if (argc != 4 || !enif_get_resource(env, argv[0], g_handle_type, (void **)&nh) || !enif_inspect_binary(env, argv[1], &img) ||
        !enif_get_int(env, argv[2], &w) || !enif_get_int(env, argv[3], &h))
        return enif_make_badarg(env);
    if (!nh->h) return mk_error(env, "closed");
    if (w < 1 || h < 1 || (size_t)w * (size_t)h != img.size) return mk_error(env, "size_mismatch");
    ERL_NIF_TERM kp_term, desc_term;
    unsigned char *kp = enif_make_new_binary(env, (size_t)nh->cap * sizeof(orbx_keypoint), &kp_term);
    unsigned char *desc = enif_make_new_binary(env, (size_t)nh->cap * ORBX_DESC_BYTES, &desc_term);
    if (!kp || !desc) return mk_error(env, "enomem");
    int n = 0, mono = -1;
    LOCK_OR_BUSY(nh);
    int rc = orbx_extract(nh->h, img.data, w, h, w, 0, 1000, (orbx_keypoint *)kp, desc, nh->cap, &n, &mono);
    enif_mutex_unlock(nh->lock);
    if (rc != ORBX_OK) return mk_error(env, code_atom(rc));
    return enif_make_tuple5(env, enif_make_atom(env, "ok"), enif_make_int(env, n), enif_make_int(env, mono),
                            enif_make_sub_binary(env, kp_term, 0, (size_t)n * sizeof(orbx_keypoint)),
                            enif_make_sub_binary(env, desc_term, 0, (size_t)n * ORBX_DESC_BYTES));
}

/* extract_color(handle, pixels_binary, width, height, format) with format = 1 RGB | 2 BGR | 3 RGBA | 4 BGRA (ORBX_FMT_*): the
 * Evision.Mat of camera_producer.ex (BGR) goes in as it is; cvtColor runs on the device (include/orbx.h, orbx_set_input_format).
 * Same result tuple as extract/4. */
static ERL_NIF_TERM nif_extract_color(ErlNifEnv *env, int argc, const ERL_NIF_TERM argv[]) {
    nif_handle *nh;
    ErlNifBinary img;
    int w, h, fmt;
    if (argc != 5 || !enif_get_resource(env, argv[0], g_handle_type, (void **)&nh) || !enif_inspect_binary(env, argv[1], &img) ||
        !enif_get_int(env, argv[2], &w) || !enif_get_int(env, argv[3], &h) || !enif_get_int(env, argv[4], &fmt))
        return enif_make_badarg(env);
    if (!nh->h) return mk_error(env, "closed");
    if (fmt < ORBX_FMT_RGB8 || fmt > ORBX_FMT_BGRA8) return enif_make_badarg(env);
    const int bpp = fmt >= ORBX_FMT_RGBA8 ? 4 : 3;
    if (w < 1 || h < 1 || (size_t)w * (size_t)h * (size_t)bpp != img.size) return mk_error(env, "size_mismatch");
    ERL_NIF_TERM kp_term, desc_term;
    unsigned char *kp = enif_make_new_binary(env, (size_t)nh->cap * sizeof(orbx_keypoint), &kp_term);
    unsigned char *desc = enif_make_new_binary(env, (size_t)nh->cap * ORBX_DESC_BYTES, &desc_term);
    if (!kp || !desc) return mk_error(env, "enomem");
    int n = 0, mono = -1;
    LOCK_OR_BUSY(nh);
    int rc = orbx_set_input_format(nh->h, fmt, ORBX_GRAY_Q15);
    if (rc == ORBX_OK) rc = orbx_extract(nh->h, img.data, w, h, w * bpp, 0, 1000, (orbx_keypoint *)kp, desc, nh->cap, &n, &mono);
    orbx_set_input_format(nh->h, ORBX_FMT_GRAY8, ORBX_GRAY_Q15);
    enif_mutex_unlock(nh->lock);
    if (rc != ORBX_OK) return mk_error(env, code_atom(rc));
    return enif_make_tuple5(env, enif_make_atom(env, "ok"), enif_make_int(env, n), enif_make_int(env, mono),
                            enif_make_sub_binary(env, kp_term, 0, (size_t)n * sizeof(orbx_keypoint)),
                            enif_make_sub_binary(env, desc_term, 0, (size_t)n * ORBX_DESC_BYTES));
}

/* extract_ppm(handle, ppm_binary, camera_rgb) -> {:ok, n, mono_index, keypoints_binary, descriptors_binary, width, height}: the binary
 * SlamHandler already builds for the wire (`Evision.imencode(".ppm", mat)`, send_slam/lib/send_slam/slam_handler.ex:66,275-277) goes in
 * unchanged; header parsing, the gray conversion the backend would do (imdecode + cvtColor by Camera.RGB, camera_rgb = 1 for the
 * reference's `rgb: 1`, slam_handler.ex:222) and extraction happen in the library.  {:error, :empty_image} where the backend would log
 * "Failed to decode frame image data" and skip the frame (orbslam3_mono_networked.cc:547-551). */
static ERL_NIF_TERM nif_extract_ppm(ErlNifEnv *env, int argc, const ERL_NIF_TERM argv[]) {
    nif_handle *nh;
    ErlNifBinary ppm;
    int camera_rgb;
    if (argc != 3 || !enif_get_resource(env, argv[0], g_handle_type, (void **)&nh) || !enif_inspect_binary(env, argv[1], &ppm) ||
        !enif_get_int(env, argv[2], &camera_rgb))
        return enif_make_badarg(env);
    if (!nh->h) return mk_error(env, "closed");
    ERL_NIF_TERM kp_term, desc_term;
    unsigned char *kp = enif_make_new_binary(env, (size_t)nh->cap * sizeof(orbx_keypoint), &kp_term);
    unsigned char *desc = enif_make_new_binary(env, (size_t)nh->cap * ORBX_DESC_BYTES, &desc_term);
    if (!kp || !desc) return mk_error(env, "enomem");
    int n = 0, mono = -1, w = 0, h = 0;
    LOCK_OR_BUSY(nh);
    int rc = orbx_extract_pnm(nh->h, ppm.data, ppm.size, camera_rgb != 0, 0, 1000, (orbx_keypoint *)kp, desc, nh->cap, &n, &mono, &w, &h);
    enif_mutex_unlock(nh->lock);
    if (rc != ORBX_OK) return mk_error(env, code_atom(rc));
    return enif_make_tuple7(env, enif_make_atom(env, "ok"), enif_make_int(env, n), enif_make_int(env, mono),
                            enif_make_sub_binary(env, kp_term, 0, (size_t)n * sizeof(orbx_keypoint)),
                            enif_make_sub_binary(env, desc_term, 0, (size_t)n * ORBX_DESC_BYTES), enif_make_int(env, w), enif_make_int(env, h));
}

/* match_windowed(handle, q_desc, q_uvr, q_levels, t_kp, t_desc, {minx, miny, maxx, maxy}) -> {:ok, best_idx, best_dist, second_idx, second_dist} (int32 binaries) */
static ERL_NIF_TERM nif_match_windowed(ErlNifEnv *env, int argc, const ERL_NIF_TERM argv[]) {
    nif_handle *nh;
    ErlNifBinary qd, quvr, qlev, tkp, td;
    const ERL_NIF_TERM *b;
    int arity;
    double bd[4];
    if (argc != 7 || !enif_get_resource(env, argv[0], g_handle_type, (void **)&nh) || !enif_inspect_binary(env, argv[1], &qd) ||
        !enif_inspect_binary(env, argv[2], &quvr) || !enif_inspect_binary(env, argv[3], &qlev) || !enif_inspect_binary(env, argv[4], &tkp) ||
        !enif_inspect_binary(env, argv[5], &td) || !enif_get_tuple(env, argv[6], &arity, &b) || arity != 4)
        return enif_make_badarg(env);
    for (int i = 0; i < 4; i++) if (!enif_get_double(env, b[i], &bd[i])) return enif_make_badarg(env);
    if (!nh->h) return mk_error(env, "closed");
    const int nq = (int)(qd.size / 32), nt = (int)(td.size / 32);
    if (qd.size % 32 || td.size % 32 || quvr.size != (size_t)nq * 12 || qlev.size != (size_t)nq * 8 || tkp.size != (size_t)nt * sizeof(orbx_keypoint))
        return mk_error(env, "size_mismatch");
    float bounds[4] = {(float)bd[0], (float)bd[1], (float)bd[2], (float)bd[3]};
    ERL_NIF_TERM t[4];
    int32_t *o[4];
    unsigned char *ob[4];
    for (int i = 0; i < 4; i++) { ob[i] = enif_make_new_binary(env, (size_t)nq * 4, &t[i]); if (!ob[i]) return mk_error(env, "enomem"); }
    /* typed views need aligned memory on both sides: inputs are copied out of the binaries, results are copied into them */
    float *a_uvr = (float *)aligned_copy(&quvr);
    int32_t *a_lev = (int32_t *)aligned_copy(&qlev);
    orbx_keypoint *a_kp = (orbx_keypoint *)aligned_copy(&tkp);
    int32_t *res = (int32_t *)enif_alloc((size_t)(nq ? nq : 1) * 16);
    int rc = ORBX_E_INVALID;
    if (a_uvr && a_lev && a_kp && res) {
        for (int i = 0; i < 4; i++) o[i] = res + (size_t)i * nq;
        if (enif_mutex_trylock(nh->lock) != 0) rc = 1;
        else {
            rc = orbx_match_windowed(nh->h, qd.data, a_uvr, a_lev, nq, a_kp, td.data, nt, bounds, o[0], o[1], o[2], o[3]);
            enif_mutex_unlock(nh->lock);
        }
        if (rc == ORBX_OK) for (int i = 0; i < 4; i++) memcpy(ob[i], o[i], (size_t)nq * 4);
    }
    if (a_uvr) enif_free(a_uvr);
    if (a_lev) enif_free(a_lev);
    if (a_kp) enif_free(a_kp);
    if (res) enif_free(res);
    if (rc == 1) return mk_error(env, "busy");
    if (rc != ORBX_OK) return mk_error(env, code_atom(rc));
    return enif_make_tuple5(env, enif_make_atom(env, "ok"), t[0], t[1], t[2], t[3]);
}

/* extract_batch(handle, frames_binary, batch, width, height): `batch` gray frames back to back in one binary (a camera group's
 * frames of one tick, or one camera's backlog) through orbx_extract_batch (uploads pipelined with the kernels) ->
 * {:ok, counts_binary (batch x int32), mono_binary (batch x int32), keypoints_binary, descriptors_binary, cap}: frame i owns records
 * [i * cap, i * cap + counts[i]) of the two result binaries.  The handle must have been created with max_batch >= batch. */
static ERL_NIF_TERM nif_extract_batch(ErlNifEnv *env, int argc, const ERL_NIF_TERM argv[]) {
    nif_handle *nh;
    ErlNifBinary img;
    int batch, w, h;
    if (argc != 5 || !enif_get_resource(env, argv[0], g_handle_type, (void **)&nh) || !enif_inspect_binary(env, argv[1], &img) ||
        !enif_get_int(env, argv[2], &batch) || !enif_get_int(env, argv[3], &w) || !enif_get_int(env, argv[4], &h))
        return enif_make_badarg(env);
    if (!nh->h) return mk_error(env, "closed");
    if (batch < 1 || batch > nh->max_batch) return mk_error(env, "capacity");
    if (w < 1 || h < 1 || (size_t)w * (size_t)h * (size_t)batch != img.size) return mk_error(env, "size_mismatch");
    ERL_NIF_TERM kp_term, desc_term, n_term, mono_term;
    unsigned char *kp = enif_make_new_binary(env, (size_t)batch * nh->cap * sizeof(orbx_keypoint), &kp_term);
    unsigned char *desc = enif_make_new_binary(env, (size_t)batch * nh->cap * ORBX_DESC_BYTES, &desc_term);
    unsigned char *nb = enif_make_new_binary(env, (size_t)batch * 4, &n_term);
    unsigned char *mb = enif_make_new_binary(env, (size_t)batch * 4, &mono_term);
    const uint8_t **frames = (const uint8_t **)enif_alloc(sizeof(uint8_t *) * (size_t)batch);
    int *counts = (int *)enif_alloc(sizeof(int) * 2 * (size_t)batch);
    if (!kp || !desc || !nb || !mb || !frames || !counts) { if (frames) enif_free((void *)frames); if (counts) enif_free(counts); return mk_error(env, "enomem"); }
    for (int i = 0; i < batch; i++) frames[i] = img.data + (size_t)i * w * h;
    int rc = 1;
    if (enif_mutex_trylock(nh->lock) == 0) {
        /* the binaries' payloads are byte arrays to the library's DMA / memcpy paths: no typed access on unaligned memory */
        rc = orbx_extract_batch(nh->h, frames, batch, w, h, w, 0, 1000, (orbx_keypoint *)kp, desc, nh->cap, counts, counts + batch);
        enif_mutex_unlock(nh->lock);
    }
    if (rc == ORBX_OK) { memcpy(nb, counts, (size_t)batch * 4); memcpy(mb, counts + batch, (size_t)batch * 4); }
    enif_free((void *)frames); enif_free(counts);
    if (rc == 1) return mk_error(env, "busy");
    if (rc != ORBX_OK) return mk_error(env, code_atom(rc));
    return enif_make_tuple6(env, enif_make_atom(env, "ok"), n_term, mono_term, kp_term, desc_term, enif_make_int(env, nh->cap));
}

/* knn2_create(descriptors_binary (rows x 32 B), device, row_offset) -> {:ok, db}: one row shard of a descriptor database resident in
 * HBM (orbx_knn2_create_db); row_offset = global index of the shard's first row. */
static ERL_NIF_TERM nif_knn2_create(ErlNifEnv *env, int argc, const ERL_NIF_TERM argv[]) {
    ErlNifBinary rows;
    int dev;
    long off;
    if (argc != 3 || !enif_inspect_binary(env, argv[0], &rows) || !enif_get_int(env, argv[1], &dev) || !enif_get_long(env, argv[2], &off))
        return enif_make_badarg(env);
    if (rows.size % ORBX_DESC_BYTES || off < 0) return mk_error(env, "size_mismatch");
    orbx_db *db = NULL;
    int rc = orbx_knn2_create_db(dev, rows.data, (long long)(rows.size / ORBX_DESC_BYTES), (long long)off, &db);
    if (rc != ORBX_OK) return mk_error(env, code_atom(rc));
    nif_db *nd = (nif_db *)enif_alloc_resource(g_db_type, sizeof(nif_db));
    if (!nd) { orbx_knn2_destroy_db(db); return mk_error(env, "enomem"); }
    nd->db = db;
    nd->lock = enif_mutex_create((char *)"orbx_db");
    if (!nd->lock) { enif_release_resource(nd); return mk_error(env, "enomem"); }
    ERL_NIF_TERM term = enif_make_resource(env, nd);
    enif_release_resource(nd);
    return enif_make_tuple2(env, enif_make_atom(env, "ok"), term);
}

/* knn2(db, queries_binary (nq x 32 B), backend) -> {:ok, idx_binary (nq x 2 int32, global rows, -1 = missing), dist_binary (nq x 2
 * int32)}: cv::BFMatcher(NORM_HAMMING).knnMatch(k = 2) of the queries against the shard; backend 0 = POPC, 1 = tensor cores (int8),
 * 2 = tensor cores (block-scaled FP4, the fastest); identical results. */
static ERL_NIF_TERM nif_knn2(ErlNifEnv *env, int argc, const ERL_NIF_TERM argv[]) {
    nif_db *nd;
    ErlNifBinary q;
    int backend;
    if (argc != 3 || !enif_get_resource(env, argv[0], g_db_type, (void **)&nd) || !enif_inspect_binary(env, argv[1], &q) ||
        !enif_get_int(env, argv[2], &backend))
        return enif_make_badarg(env);
    if (!nd->db) return mk_error(env, "closed");
    if (q.size % ORBX_DESC_BYTES || q.size == 0) return mk_error(env, "size_mismatch");
    const int nq = (int)(q.size / ORBX_DESC_BYTES);
    ERL_NIF_TERM it, dt;
    unsigned char *ib = enif_make_new_binary(env, (size_t)nq * 8, &it), *db_ = enif_make_new_binary(env, (size_t)nq * 8, &dt);
    int32_t *res = (int32_t *)enif_alloc((size_t)nq * 16);
    if (!ib || !db_ || !res) { if (res) enif_free(res); return mk_error(env, "enomem"); }
    int rc = 1;
    if (enif_mutex_trylock(nd->lock) == 0) {
        rc = orbx_knn2_set_backend(nd->db, backend);
        if (rc == ORBX_OK) rc = orbx_knn2_query(nd->db, q.data, nq, res, res + (size_t)2 * nq);
        enif_mutex_unlock(nd->lock);
    }
    if (rc == ORBX_OK) { memcpy(ib, res, (size_t)nq * 8); memcpy(db_, res + (size_t)2 * nq, (size_t)nq * 8); }
    enif_free(res);
    if (rc == 1) return mk_error(env, "busy");
    if (rc != ORBX_OK) return mk_error(env, code_atom(rc));
    return enif_make_tuple3(env, enif_make_atom(env, "ok"), it, dt);
}

static int on_load(ErlNifEnv *env, void **priv, ERL_NIF_TERM info) {
    (void)priv; (void)info;
    g_handle_type = enif_open_resource_type(env, NULL, "orbx_handle", handle_dtor, ERL_NIF_RT_CREATE, NULL);
    g_db_type = enif_open_resource_type(env, NULL, "orbx_db", db_dtor, ERL_NIF_RT_CREATE, NULL);
    return g_handle_type && g_db_type ? 0 : 1;
}

static ErlNifFunc nif_funcs[] = {
    {"create", 8, nif_create, ERL_NIF_DIRTY_JOB_IO_BOUND},
    {"create", 9, nif_create, ERL_NIF_DIRTY_JOB_IO_BOUND},
    {"extract_batch", 5, nif_extract_batch, ERL_NIF_DIRTY_JOB_IO_BOUND},
    {"knn2_create", 3, nif_knn2_create, ERL_NIF_DIRTY_JOB_IO_BOUND},
    {"knn2", 3, nif_knn2, ERL_NIF_DIRTY_JOB_IO_BOUND},
    {"extract", 4, nif_extract, ERL_NIF_DIRTY_JOB_IO_BOUND},
    {"extract_color", 5, nif_extract_color, ERL_NIF_DIRTY_JOB_IO_BOUND},
    {"extract_ppm", 3, nif_extract_ppm, ERL_NIF_DIRTY_JOB_IO_BOUND},
    {"match_windowed", 7, nif_match_windowed, ERL_NIF_DIRTY_JOB_IO_BOUND},
#ifdef ORBX_NIF_MIN
    {NULL, 0, NULL, 0},   /* sentinel for the mock host (nif/mock_host.c); the real ERL_NIF_INIT takes the array size */
#endif
};

ERL_NIF_INIT(Elixir.SendSlam.OrbNif, nif_funcs, on_load, NULL, NULL, NULL)

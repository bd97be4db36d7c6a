/* C harness for the row-sharded kNN entry points of include/orbx.h (orbx_comm_*, orbx_knn2_query_sharded): what a C / C++ backend
 * that links liborbx.so would do, one process per GPU, no Python and no torch in the process.
 *
 * Launch (2 or more GPUs):  python -m torch.distributed.run --no-python --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
 *                              --master-port P  tests/harness/_build/knn2_sharded_driver [rows_total] [queries]
 * Every rank reads RANK / WORLD_SIZE / LOCAL_RANK / MASTER_PORT from the environment.  Rank 0 creates the NCCL unique id and
 * publishes it through a file in /tmp named after MASTER_PORT; the other ranks poll for it.  The database is a deterministic
 * pseudo-random table every rank can regenerate; rank r holds rows [start_r, stop_r) (ragged split).  Rank 0 additionally holds
 * the WHOLE table as a single shard and checks that the sharded answer equals the single-shard answer bit for bit, and that
 * queries copied from known rows find those rows at distance 0.  Prints "knn2_sharded_driver OK ..." on success; exit code 0. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "orbx.h"

static uint64_t mix(uint64_t x) { x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31); }
static void fill_rows(uint8_t *dst, long long row0, long long n) {
    for (long long r = 0; r < n; r++) for (int w = 0; w < 4; w++) { const uint64_t v = mix((uint64_t)(row0 + r) * 4 + (uint64_t)w); memcpy(dst + r * 32 + w * 8, &v, 8); }
}
static int env_int(const char *k, int d) { const char *e = getenv(k); return e ? atoi(e) : d; }

int main(int argc, char **argv) {
    const int rank = env_int("RANK", 0), world = env_int("WORLD_SIZE", 1), local = env_int("LOCAL_RANK", rank), port = env_int("MASTER_PORT", 29500);
    const long long total = argc > 1 ? atoll(argv[1]) : 1000003;
    const int nq = argc > 2 ? atoi(argv[2]) : 777;
    char path[128];
    snprintf(path, sizeof(path), "/tmp/orbx_nccl_id_%d", port);
    uint8_t id[ORBX_NCCL_ID_BYTES];
    if (rank == 0) {
        if (orbx_comm_unique_id(id) != ORBX_OK) { fprintf(stderr, "unique id: %s\n", orbx_comm_last_error(NULL)); return 2; }
        char tmp[160]; snprintf(tmp, sizeof(tmp), "%s.tmp", path);
        FILE *f = fopen(tmp, "wb"); if (!f || fwrite(id, 1, sizeof(id), f) != sizeof(id)) return 2; fclose(f);
        if (rename(tmp, path) != 0) return 2;
    } else {
        FILE *f = NULL;
        for (int i = 0; i < 600 && !(f = fopen(path, "rb")); i++) usleep(100000);
        if (!f || fread(id, 1, sizeof(id), f) != sizeof(id)) { fprintf(stderr, "rank %d: no unique id at %s\n", rank, path); return 2; }
        fclose(f);
    }
    orbx_comm *comm = NULL;
    if (orbx_comm_create(local, rank, world, id, &comm) != ORBX_OK) { fprintf(stderr, "rank %d comm: %s\n", rank, orbx_comm_last_error(NULL)); return 3; }
    if (rank == 0) unlink(path);
    /* ragged row split: the first total % world ranks take one extra row */
    const long long base = total / world, extra = total % world;
    const long long start = rank * base + (rank < extra ? rank : extra), rows = base + (rank < extra ? 1 : 0);
    uint8_t *shard = (uint8_t *)malloc((size_t)(rows > 0 ? rows : 1) * 32);
    fill_rows(shard, start, rows);
    /* queries: rows spread over the whole table (exact hits in other ranks' shards), a few bits flipped in every second one */
    uint8_t *q = (uint8_t *)malloc((size_t)nq * 32);
    long long *src = (long long *)malloc(sizeof(long long) * (size_t)nq);
    for (int i = 0; i < nq; i++) {
        src[i] = (long long)(mix(0xABCDull + (uint64_t)i) % (uint64_t)total);
        fill_rows(q + (size_t)i * 32, src[i], 1);
        if (i & 1) for (int b = 0; b < 9; b++) q[(size_t)i * 32 + (mix((uint64_t)i * 31 + (uint64_t)b) & 31)] ^= (uint8_t)(1u << (b & 7));
    }
    orbx_db *db = NULL;
    if (orbx_knn2_create_db(local, shard, rows, start, &db) != ORBX_OK) { fprintf(stderr, "rank %d db: %s\n", rank, orbx_knn2_last_error(NULL)); return 4; }
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * 4 * (size_t)nq), *dist = idx + 2 * (size_t)nq;
    int bad = 0;
    for (int backend = 0; backend < 2 && !bad; backend++) {
        orbx_knn2_set_backend(db, backend);
        const int rc = orbx_knn2_query_sharded(db, comm, q, nq, idx, dist);
        if (rc != ORBX_OK) { fprintf(stderr, "rank %d sharded query: %s\n", rank, orbx_knn2_last_error(db)); return 5; }
        for (int i = 0; i < nq; i += 2) if (idx[2 * i] != (int32_t)src[i] || dist[2 * i] != 0) bad++;      /* unflipped queries: exact hit, global row */
        if (rank == 0) {                                                                                  /* against ONE shard holding everything */
            uint8_t *all = (uint8_t *)malloc((size_t)total * 32);
            fill_rows(all, 0, total);
            orbx_db *whole = NULL;
            if (orbx_knn2_create_db(local, all, total, 0, &whole) != ORBX_OK) return 6;
            orbx_knn2_set_backend(whole, backend);
            int32_t *widx = (int32_t *)malloc(sizeof(int32_t) * 4 * (size_t)nq), *wdist = widx + 2 * (size_t)nq;
            if (orbx_knn2_query(whole, q, nq, widx, wdist) != ORBX_OK) return 6;
            if (memcmp(widx, idx, sizeof(int32_t) * 4 * (size_t)nq) != 0) bad += 1000;
            orbx_knn2_destroy_db(whole); free(all); free(widx);
        }
    }
    orbx_knn2_destroy_db(db);
    orbx_comm_destroy(comm);
    if (bad) { fprintf(stderr, "rank %d: %d mismatches\n", rank, bad); return 7; }
    printf("knn2_sharded_driver OK rank %d of %d: %lld rows (%lld here from row %lld), %d queries, both backends\n", rank, world, total, rows, start, nq);
    return 0;
}

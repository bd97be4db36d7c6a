// Hamming matching kernels + their C ABI (include/orbx.h "matching").
//
//   k_distance_batch   ORBmatcher::DescriptorDistance over n pairs                         
//   k_match_windowed   Frame::GetFeaturesInArea + best / second-best distance per query   (C.2, SearchByProjection /
//                      SearchForInitialization inner loops)
//   k_knn2             brute-force k=2 nearest neighbours of nq queries in a row shard     (cv::BFMatcher NORM_HAMMING)
//   k_knn2_merge       top-2 merge of per-chunk / per-rank partial results
//
// Distances are XOR + POPC on eight 32-bit words (exact integers).  The kNN kernel keeps queries in registers and
// streams database rows through shared memory (broadcast reads), so the POPC pipe is the bound.
#include <cuda_runtime.h>

#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/orbx.h"
#include "orbx_dev.h"

using namespace orbx;

namespace {

__device__ __forceinline__ int hamming256(const uint4 &a0, const uint4 &a1, const uint4 &b0, const uint4 &b1) {
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

__global__ void __launch_bounds__(256) k_distance_batch(const uint4 *__restrict__ a, const uint4 *__restrict__ b, int n,
                                                        int32_t *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = hamming256(a[2 * i], a[2 * i + 1], b[2 * i], b[2 * i + 1]);
}

// ---- windowed search -------------------------------------------------------------------------------------------
constexpr int GRID_COLS = 64, GRID_ROWS = 48;   // FRAME_GRID_COLS / FRAME_GRID_ROWS of UPSTREAM Frame.h
constexpr unsigned long long NONE64 = ~0ull;

__global__ void __launch_bounds__(256) k_match_windowed(const uint4 *__restrict__ qdesc, const float *__restrict__ quvr,
                                                        const int32_t *__restrict__ qlev, int nq,
                                                        const KeypointRec *__restrict__ tkp, const uint4 *__restrict__ tdesc, int nt,
                                                        float minX, float minY, float invW, float invH,
                                                        int32_t *__restrict__ best_idx, int32_t *__restrict__ best_dist,
                                                        int32_t *__restrict__ second_idx, int32_t *__restrict__ second_dist) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    const uint4 q0 = qdesc[2 * q], q1 = qdesc[2 * q + 1];
    const float x = quvr[3 * q], y = quvr[3 * q + 1], r = quvr[3 * q + 2];
    const int minLevel = qlev[2 * q], maxLevel = qlev[2 * q + 1];
    const bool check = (minLevel > 0) || (maxLevel >= 0);
    const int cx0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, minX), r), invW)));
    const int cx1 = min(GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, minX), r), invW)));
    const int cy0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, minY), r), invH)));
    const int cy1 = min(GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, minY), r), invH)));
    unsigned long long k1 = NONE64, k2 = NONE64;
    for (int t = lane; t < nt; t += 32) {
        const KeypointRec kp = tkp[t];
        const int px = (int)roundf(__fmul_rn(__fsub_rn(kp.x, minX), invW)), py = (int)roundf(__fmul_rn(__fsub_rn(kp.y, minY), invH));
        if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) continue;   // Frame::PosInGrid rejected it
        if (px < cx0 || px > cx1 || py < cy0 || py > cy1) continue;
        if (check) {
            if (kp.octave < minLevel) continue;
            if (maxLevel >= 0 && kp.octave > maxLevel) continue;
        }
        if (!(fabsf(__fsub_rn(kp.x, x)) < r && fabsf(__fsub_rn(kp.y, y)) < r)) continue;
        const int d = hamming256(q0, q1, tdesc[2 * t], tdesc[2 * t + 1]);
        // ties resolve in the reference's visiting order: grid column, grid row, train index
        const unsigned long long key = ((unsigned long long)d << 40) | ((unsigned long long)(px * GRID_ROWS + py) << 24) | (unsigned long long)t;
        if (key < k1) { k2 = k1; k1 = key; } else if (key < k2) k2 = key;
    }
    unsigned long long b = k1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long v = __shfl_xor_sync(0xFFFFFFFFu, b, o); b = v < b ? v : b; }
    unsigned long long s = (k1 == b) ? k2 : k1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long v = __shfl_xor_sync(0xFFFFFFFFu, s, o); s = v < s ? v : s; }
    if (lane == 0) {
        best_idx[q] = b == NONE64 ? -1 : (int32_t)(b & 0xFFFFFFull);  best_dist[q] = b == NONE64 ? 256 : (int32_t)(b >> 40);
        second_idx[q] = s == NONE64 ? -1 : (int32_t)(s & 0xFFFFFFull); second_dist[q] = s == NONE64 ? 256 : (int32_t)(s >> 40);
    }
}

// The same search on the feature grid built by k_frame_grid (orbx_frame.cu): Frame::GetFeaturesInArea walks the cells
// (ix outer, iy inner) of the query window; in the CSR layout (cell = ix * 48 + iy) the cells iy0..iy1 of one grid column are
// one contiguous item range, so a warp strides over a few dozen candidates instead of every train keypoint.  The CSR position
// is monotone in the reference's visiting order (grid column, grid row, train index), which makes it the tie-break key.
// One warp, one query.  The cells iy0..iy1 of every grid column of the window are one CSR range; the ranges of all columns are fetched
// at once (lane = column), laid end to end by a warp scan, and the lanes stride over the concatenated candidate list -- four dependent
// load latencies per query (cell_start -> cell_items -> keypoint -> descriptor) instead of four per grid column.
__device__ __forceinline__ void match_grid_query(int lane, uint4 q0, uint4 q1, float x, float y, float r, int minLevel, int maxLevel,
                                                 const KeypointRec *__restrict__ tkp, const uint4 *__restrict__ tdesc,
                                                 const int32_t *__restrict__ cell_start, const int32_t *__restrict__ cell_items,
                                                 float minX, float minY, float invW, float invH, int32_t *best_idx, int32_t *best_dist,
                                                 int32_t *second_idx, int32_t *second_dist) {
    const bool check = (minLevel > 0) || (maxLevel >= 0);
    const int cx0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, minX), r), invW)));
    const int cx1 = min(GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, minX), r), invW)));
    const int cy0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, minY), r), invH)));
    const int cy1 = min(GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, minY), r), invH)));
    unsigned long long k1 = NONE64, k2 = NONE64;
    if (cx0 < GRID_COLS && cx1 >= 0 && cy0 < GRID_ROWS && cy1 >= 0 && cy0 <= cy1 && cx0 <= cx1) {
        const int ncol = cx1 - cx0 + 1;                                    // <= 64: lane holds columns cx0 + lane and cx0 + 32 + lane
        int lo[2] = {0, 0}, cnt[2] = {0, 0}, off[2];
#pragma unroll
        for (int hlf = 0; hlf < 2; hlf++) {
            const int c = 32 * hlf + lane;
            if (c < ncol) {
                lo[hlf] = cell_start[(cx0 + c) * GRID_ROWS + cy0];
                cnt[hlf] = cell_start[(cx0 + c) * GRID_ROWS + cy1 + 1] - lo[hlf];
            }
        }
        int carry = 0;
#pragma unroll
        for (int hlf = 0; hlf < 2; hlf++) {                                 // inclusive scan of the counts, columns in visiting order
            int v = cnt[hlf];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xFFFFFFFFu, v, o); if (lane >= o) v += u; }
            off[hlf] = v + carry;
            carry = __shfl_sync(0xFFFFFFFFu, off[hlf], 31);
        }
        const int total = carry;
        for (int j0 = 0; j0 < total; j0 += 32) {
            const int j = j0 + lane;
            int p = -1;                                                     // CSR position of candidate j
            for (int c = 0; c < ncol; c++) {
                const int oc = __shfl_sync(0xFFFFFFFFu, c < 32 ? off[0] : off[1], c & 31);
                const int lc = __shfl_sync(0xFFFFFFFFu, c < 32 ? lo[0] : lo[1], c & 31);
                const int nc = __shfl_sync(0xFFFFFFFFu, c < 32 ? cnt[0] : cnt[1], c & 31);
                if (p < 0 && j < oc) p = lc + (j - (oc - nc));
            }
            if (j >= total || p < 0) continue;
            const int t = cell_items[p];
            const KeypointRec kp = tkp[t];
            if (check) {
                if (kp.octave < minLevel) continue;
                if (maxLevel >= 0 && kp.octave > maxLevel) continue;
            }
            if (!(fabsf(__fsub_rn(kp.x, x)) < r && fabsf(__fsub_rn(kp.y, y)) < r)) continue;
            const int d = hamming256(q0, q1, tdesc[2 * t], tdesc[2 * t + 1]);
            const unsigned long long key = ((unsigned long long)d << 40) | (unsigned long long)p;
            if (key < k1) { k2 = k1; k1 = key; } else if (key < k2) k2 = key;
        }
    }
    unsigned long long b = k1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long v = __shfl_xor_sync(0xFFFFFFFFu, b, o); b = v < b ? v : b; }
    unsigned long long s = (k1 == b) ? k2 : k1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long v = __shfl_xor_sync(0xFFFFFFFFu, s, o); s = v < s ? v : s; }
    if (lane == 0) {
        *best_idx = b == NONE64 ? -1 : cell_items[(int)(b & 0xFFFFFFFFull)];  *best_dist = b == NONE64 ? 256 : (int32_t)(b >> 40);
        *second_idx = s == NONE64 ? -1 : cell_items[(int)(s & 0xFFFFFFFFull)]; *second_dist = s == NONE64 ? 256 : (int32_t)(s >> 40);
    }
}

__global__ void __launch_bounds__(256) k_match_windowed_grid(const uint4 *__restrict__ qdesc, const float *__restrict__ quvr,
                                                             const int32_t *__restrict__ qlev, int nq,
                                                             const KeypointRec *__restrict__ tkp, const uint4 *__restrict__ tdesc,
                                                             const int32_t *__restrict__ cell_start, const int32_t *__restrict__ cell_items,
                                                             float minX, float minY, float invW, float invH,
                                                             int32_t *__restrict__ best_idx, int32_t *__restrict__ best_dist,
                                                             int32_t *__restrict__ second_idx, int32_t *__restrict__ second_dist) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    match_grid_query(lane, qdesc[2 * q], qdesc[2 * q + 1], quvr[3 * q], quvr[3 * q + 1], quvr[3 * q + 2], qlev[2 * q], qlev[2 * q + 1], tkp, tdesc,
                     cell_start, cell_items, minX, minY, invW, invH, best_idx + q, best_dist + q, second_idx + q, second_dist + q);
}

// Several (query frame, train frame) pairs of one batch in one launch: every per-frame array is [batch][cap] as orbx_extract_batch_device
// and orbx_frame_grid_batch_device leave it; blockIdx.y = pair, the query count of a frame is read from the device-resident n[].
struct MatchPairs { int qf[kMaxMatchPairs], tf[kMaxMatchPairs]; };

__global__ void __launch_bounds__(256) k_match_windowed_grid_batch(const __grid_constant__ MatchPairs prs, int cap, const uint4 *__restrict__ qdesc,
                                                                   const float *__restrict__ quvr, const int32_t *__restrict__ qlev,
                                                                   const int32_t *__restrict__ nq_of, const KeypointRec *__restrict__ tkp,
                                                                   const uint4 *__restrict__ tdesc, const int32_t *__restrict__ cell_start,
                                                                   const int32_t *__restrict__ cell_items, float minX, float minY, float invW,
                                                                   float invH, int32_t *__restrict__ best_idx, int32_t *__restrict__ best_dist,
                                                                   int32_t *__restrict__ second_idx, int32_t *__restrict__ second_dist) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int qf = prs.qf[blockIdx.y], tf = prs.tf[blockIdx.y];
    if (q >= min(nq_of[qf], cap)) return;
    const size_t qi = (size_t)qf * cap + q, tb = (size_t)tf * cap;
    match_grid_query(lane, qdesc[2 * qi], qdesc[2 * qi + 1], quvr[3 * qi], quvr[3 * qi + 1], quvr[3 * qi + 2], qlev[2 * qi], qlev[2 * qi + 1],
                     tkp + tb, tdesc + 2 * tb, cell_start + (size_t)tf * (GRID_COLS * GRID_ROWS + 1), cell_items + tb, minX, minY, invW, invH,
                     best_idx + qi, best_dist + qi, second_idx + qi, second_dist + qi);
}

// ---- brute-force kNN, k = 2 -----------------------------------------------------------------------------------
constexpr int KQ_THREADS = 128;    // threads per CTA
constexpr int KQ_QPT = 2;          // queries per thread (registers)
constexpr int KQ_TILE = 128;       // database rows staged per step
constexpr int KQ_ROW_BITS = 23;    // row-in-chunk bits of the packed (distance, row) key

__global__ void __launch_bounds__(KQ_THREADS) k_knn2(const uint4 *__restrict__ db, long long nrows, long long row_offset,
                                                     const uint4 *__restrict__ queries, int nq, int rows_per_chunk,
                                                     unsigned long long *__restrict__ partial) {
    __shared__ uint4 s_db[2][KQ_TILE * 2];
    const int tid = threadIdx.x;
    const int qa = blockIdx.x * (KQ_THREADS * KQ_QPT) + tid, qb = qa + KQ_THREADS;
    const uint4 z = make_uint4(0, 0, 0, 0);
    const uint4 a0 = qa < nq ? queries[2 * qa] : z, a1 = qa < nq ? queries[2 * qa + 1] : z;
    const uint4 b0 = qb < nq ? queries[2 * qb] : z, b1 = qb < nq ? queries[2 * qb + 1] : z;
    const long long r0 = (long long)blockIdx.y * rows_per_chunk;
    const long long r1 = min(r0 + (long long)rows_per_chunk, nrows);
    uint32_t ka1 = 0xFFFFFFFFu, ka2 = 0xFFFFFFFFu, kb1 = 0xFFFFFFFFu, kb2 = 0xFFFFFFFFu;
    const int ntiles = (int)((r1 - r0 + KQ_TILE - 1) / KQ_TILE);
    // prologue: stage tile 0
    if (ntiles > 0) {
        const int cnt = (int)min((long long)KQ_TILE, r1 - r0);
        for (int i = tid; i < cnt * 2; i += KQ_THREADS) s_db[0][i] = db[r0 * 2 + i];
    }
    __syncthreads();
    for (int t = 0; t < ntiles; t++) {
        const long long base = r0 + (long long)t * KQ_TILE;
        const int cnt = (int)min((long long)KQ_TILE, r1 - base);
        // prefetch the next tile into registers while this one is consumed
        uint4 pre[2]; int npre = 0;
        if (t + 1 < ntiles) {
            const long long nb = base + KQ_TILE;
            const int ncnt = (int)min((long long)KQ_TILE, r1 - nb);
            for (int i = tid, k = 0; i < ncnt * 2; i += KQ_THREADS, k++) { pre[k] = db[nb * 2 + i]; npre = k + 1; }
        }
        const uint4 *__restrict__ rows = s_db[t & 1];
        uint32_t rowkey = (uint32_t)(base - r0);
#pragma unroll 4
        for (int r = 0; r < cnt; r++, rowkey++) {
            const uint4 d0 = rows[2 * r], d1 = rows[2 * r + 1];
            const uint32_t da = (uint32_t)hamming256(a0, a1, d0, d1), dbb = (uint32_t)hamming256(b0, b1, d0, d1);
            const uint32_t keya = (da << KQ_ROW_BITS) + rowkey, keyb = (dbb << KQ_ROW_BITS) + rowkey;
            ka2 = min(ka2, max(ka1, keya)); ka1 = min(ka1, keya);
            kb2 = min(kb2, max(kb1, keyb)); kb1 = min(kb1, keyb);
        }
        if (t + 1 < ntiles) {
            for (int k = 0; k < npre; k++) s_db[(t + 1) & 1][tid + k * KQ_THREADS] = pre[k];
        }
        __syncthreads();
    }
    auto emit = [&](int q, uint32_t k1, uint32_t k2) {
        if (q >= nq) return;
        unsigned long long *o = partial + ((size_t)blockIdx.y * nq + q) * 2;
        const uint32_t m = (1u << KQ_ROW_BITS) - 1;
        o[0] = k1 == 0xFFFFFFFFu ? NONE64 : (((unsigned long long)(k1 >> KQ_ROW_BITS) << 32) | (unsigned long long)(row_offset + r0 + (k1 & m)));
        o[1] = k2 == 0xFFFFFFFFu ? NONE64 : (((unsigned long long)(k2 >> KQ_ROW_BITS) << 32) | (unsigned long long)(row_offset + r0 + (k2 & m)));
    };
    emit(qa, ka1, ka2);
    emit(qb, kb1, kb2);
}

// 32 queries per CTA (lanes), 16 warps striding over the partial blocks (coalesced 512-byte reads per warp), then one pass over the 16
// per-warp results in shared memory.  The two smallest distinct keys do not depend on the visiting order.
__device__ __forceinline__ void top2_insert(unsigned long long key, unsigned long long &k1, unsigned long long &k2) {
    if (key < k1) { k2 = k1; k1 = key; } else if (key < k2 && key != k1) k2 = key;
}

__global__ void __launch_bounds__(512) k_knn2_merge(const unsigned long long *__restrict__ partial, int nparts, int nq,
                                                    unsigned long long *__restrict__ out) {
    __shared__ unsigned long long sm[16][32][2];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, q = blockIdx.x * 32 + lane;
    unsigned long long k1 = NONE64, k2 = NONE64;
    if (q < nq) {
#pragma unroll 4
        for (int p = w; p < nparts; p += 16) {
            const unsigned long long *v = partial + ((size_t)p * nq + q) * 2;
            const unsigned long long a = v[0], b = v[1];
            top2_insert(a, k1, k2); top2_insert(b, k1, k2);
        }
    }
    sm[w][lane][0] = k1; sm[w][lane][1] = k2;
    __syncthreads();
    if (w == 0 && q < nq) {
#pragma unroll
        for (int i = 1; i < 16; i++) { top2_insert(sm[i][lane][0], k1, k2); top2_insert(sm[i][lane][1], k1, k2); }
        out[2 * q] = k1; out[2 * q + 1] = k2;
    }
}

static inline void launch_merge(const unsigned long long *parts, int nparts, int nq, unsigned long long *out, cudaStream_t stream) {
    k_knn2_merge<<<(nq + 31) / 32, 512, 0, stream>>>(parts, nparts, nq, out);
}

thread_local std::string g_db_error;

void merge_launch(const unsigned long long *parts, int nparts, int nq, unsigned long long *out, cudaStream_t stream) {
    launch_merge(parts, nparts, nq, out, stream);
}

}  // namespace

struct orbx_db {
    int device = 0;
    cudaStream_t stream = nullptr, own_stream = nullptr;
    const uint8_t *d_rows = nullptr;
    bool owns_rows = false;
    long long nrows = 0, row_offset = 0;
    int nchunks = 0, rows_per_chunk = 0;
    uint8_t *d_q = nullptr; int q_cap = 0;
    unsigned long long *d_partial = nullptr; size_t partial_cap = 0;
    unsigned long long *d_out = nullptr;
    unsigned long long *h_out = nullptr;
    mutable std::string err;
    long long launches = 0;
    // tensor-core backend (orbx_knn_tc.cu): {-1,+1} int8 expansion of the shard (built on first use) and of the queries
    int backend = 2;                 // ORBX_KNN_TENSOR_FP4 (the fastest of the three; all three give identical results)
    int sm_count = 0;
    int8_t *d_dbe = nullptr;
    int8_t *d_qe = nullptr; int qe_cap = 0;
    uint8_t *d_dbe4 = nullptr;       // FP4 backend: E2M1 expansion of the shard (128 B per row) and of the queries
    uint8_t *d_qe4 = nullptr; int qe4_cap = 0;
    unsigned long long *d_partial_tc = nullptr; size_t partial_tc_cap = 0;
    // row-sharded queries: this rank's top-2 and the all-gathered [ranks][nq][2] partials
    unsigned long long *d_sh_local = nullptr, *d_sh_gather = nullptr; size_t sh_cap = 0;
};

#define DB_TRY(db, expr)                                                                       \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess) { (db)->err = std::string(#expr) + ": " + cudaGetErrorString(e__); return ORBX_E_CUDA; } \
    } while (0)

static int db_reserve(orbx_db *db, int nq) {
    if (nq <= db->q_cap) return ORBX_OK;
    if (db->d_q) cudaFree(db->d_q);
    if (db->d_partial) cudaFree(db->d_partial);
    if (db->d_out) cudaFree(db->d_out);
    if (db->h_out) cudaFreeHost(db->h_out);
    db->d_q = nullptr; db->d_partial = nullptr; db->d_out = nullptr; db->h_out = nullptr; db->q_cap = 0;
    DB_TRY(db, cudaMalloc((void **)&db->d_q, (size_t)nq * 32));
    DB_TRY(db, cudaMalloc((void **)&db->d_partial, (size_t)db->nchunks * nq * 2 * sizeof(unsigned long long)));
    DB_TRY(db, cudaMalloc((void **)&db->d_out, (size_t)nq * 2 * sizeof(unsigned long long)));
    DB_TRY(db, cudaMallocHost((void **)&db->h_out, (size_t)nq * 2 * sizeof(unsigned long long)));
    db->q_cap = nq;
    return ORBX_OK;
}

static int db_create_common(int device, long long nrows, long long row_offset, orbx_db **out) {
    if (!out) { g_db_error = "null argument"; return ORBX_E_INVALID; }
    *out = nullptr;
    if (nrows < 0 || row_offset < 0 || row_offset + nrows >= (1ll << 32)) { g_db_error = "row count / offset out of range"; return ORBX_E_INVALID; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev < 1) { g_db_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (orbx has no CPU fallback)"; return ORBX_E_CUDA; }
    if (device < 0 || device >= ndev) { g_db_error = "device ordinal out of range"; return ORBX_E_INVALID; }
    orbx_db *db = new (std::nothrow) orbx_db();
    if (!db) { g_db_error = "out of host memory"; return ORBX_E_INVALID; }
    db->device = device; db->nrows = nrows; db->row_offset = row_offset;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&db->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        g_db_error = std::string("cuda init: ") + cudaGetErrorString(e); delete db; return ORBX_E_CUDA;
    }
    // chunking: enough CTAs to fill 148 SMs a few times for ~2k queries, chunk <= 2^23 rows (packed key)
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, device);
    db->sm_count = prop.multiProcessorCount;
    int chunks = std::max(1, prop.multiProcessorCount * 4 / 8);
    long long rpc = (nrows + chunks - 1) / std::max(chunks, 1);
    rpc = std::max<long long>(rpc, KQ_TILE * 8);
    rpc = (rpc + KQ_TILE - 1) / KQ_TILE * KQ_TILE;
    rpc = std::min<long long>(rpc, (1ll << KQ_ROW_BITS) - KQ_TILE);
    db->own_stream = db->stream;
    db->rows_per_chunk = (int)rpc;
    db->nchunks = (int)std::max<long long>(1, (nrows + rpc - 1) / rpc);
    *out = db;
    return ORBX_OK;
}

extern "C" {

int orbx_knn2_create_db(int device, const uint8_t *rows, long long nrows, long long row_offset, orbx_db **out) {
    if (!rows && nrows > 0) { g_db_error = "null rows"; return ORBX_E_INVALID; }
    int rc = db_create_common(device, nrows, row_offset, out);
    if (rc) return rc;
    orbx_db *db = *out;
    uint8_t *d = nullptr;
    cudaError_t e = cudaMalloc((void **)&d, std::max<size_t>((size_t)nrows * 32, 256));
    if (e == cudaSuccess && nrows > 0) e = cudaMemcpy(d, rows, (size_t)nrows * 32, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamLegacy);   // the tail of a pageable upload may still be in DMA; queries run on a non-blocking stream
    if (e != cudaSuccess) {
        g_db_error = std::string("db upload: ") + cudaGetErrorString(e);
        if (d) cudaFree(d);
        cudaStreamDestroy(db->stream); delete db; *out = nullptr; return ORBX_E_CUDA;
    }
    db->d_rows = d; db->owns_rows = true;
    return ORBX_OK;
}

int orbx_knn2_create_db_device(int device, const uint8_t *d_rows, long long nrows, long long row_offset, orbx_db **out) {
    if (!d_rows && nrows > 0) { g_db_error = "null rows"; return ORBX_E_INVALID; }
    if (((uintptr_t)d_rows & 15) != 0) { g_db_error = "device rows must be 16-byte aligned"; return ORBX_E_INVALID; }
    int rc = db_create_common(device, nrows, row_offset, out);
    if (rc) return rc;
    (*out)->d_rows = d_rows; (*out)->owns_rows = false;
    return ORBX_OK;
}

void orbx_knn2_destroy_db(orbx_db *db) {
    if (!db) return;
    cudaSetDevice(db->device);
    if (db->stream) cudaStreamSynchronize(db->stream);
    if (db->owns_rows && db->d_rows) cudaFree(const_cast<uint8_t *>(db->d_rows));
    if (db->d_q) cudaFree(db->d_q);
    if (db->d_partial) cudaFree(db->d_partial);
    if (db->d_out) cudaFree(db->d_out);
    if (db->h_out) cudaFreeHost(db->h_out);
    if (db->d_dbe) cudaFree(db->d_dbe);
    if (db->d_qe) cudaFree(db->d_qe);
    if (db->d_dbe4) cudaFree(db->d_dbe4);
    if (db->d_qe4) cudaFree(db->d_qe4);
    if (db->d_partial_tc) cudaFree(db->d_partial_tc);
    if (db->d_sh_local) cudaFree(db->d_sh_local);
    if (db->d_sh_gather) cudaFree(db->d_sh_gather);
    if (db->own_stream) cudaStreamDestroy(db->own_stream);
    delete db;
}

int orbx_knn2_set_stream(orbx_db *db, void *cuda_stream) {
    if (!db) return ORBX_E_INVALID;
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : db->own_stream;
    if (s == db->stream) return ORBX_OK;
    DB_TRY(db, cudaSetDevice(db->device));
    DB_TRY(db, cudaStreamSynchronize(db->stream));
    db->stream = s;
    return ORBX_OK;
}

const char *orbx_knn2_last_error(const orbx_db *db) { return db ? db->err.c_str() : g_db_error.c_str(); }
long long orbx_knn2_launch_count(const orbx_db *db) { return db ? db->launches : 0; }

int orbx_knn2_sync(orbx_db *db) {
    if (!db) return ORBX_E_INVALID;
    DB_TRY(db, cudaSetDevice(db->device));
    DB_TRY(db, cudaStreamSynchronize(db->stream));
    return ORBX_OK;
}

int orbx_knn2_set_backend(orbx_db *db, int backend) {
    if (!db || (backend != ORBX_KNN_POPC && backend != ORBX_KNN_TENSOR && backend != ORBX_KNN_TENSOR_FP4)) return ORBX_E_INVALID;
    db->backend = backend;
    return ORBX_OK;
}

// tensor-core path: expand (once) the shard and (per call) the queries to {-1,+1} int8, GEMM tiles + top-2 in TMEM
static int query_device_tensor(orbx_db *db, const uint8_t *d_queries, int nq, unsigned long long *d_packed_out) {
    if (!db->d_dbe) {
        const long long rows_pad = knn_tc_padded_rows(db->nrows);
        DB_TRY(db, cudaMalloc((void **)&db->d_dbe, (size_t)rows_pad * 256));
        db->launches += launch_expand_pm1(db->d_rows, db->nrows, rows_pad, db->d_dbe, db->stream);
        DB_TRY(db, cudaGetLastError());
    }
    const int maxq = knn_tc_max_queries();
    const int qcap = knn_tc_padded_queries(std::min(nq, maxq));
    if (qcap > db->qe_cap) {
        if (db->d_qe) cudaFree(db->d_qe);
        db->d_qe = nullptr; db->qe_cap = 0;
        DB_TRY(db, cudaMalloc((void **)&db->d_qe, (size_t)qcap * 256));
        db->qe_cap = qcap;
    }
    const size_t need = (size_t)(db->sm_count + 1) * std::min(nq, maxq) * 2;
    if (need > db->partial_tc_cap) {
        if (db->d_partial_tc) cudaFree(db->d_partial_tc);
        db->d_partial_tc = nullptr; db->partial_tc_cap = 0;
        DB_TRY(db, cudaMalloc((void **)&db->d_partial_tc, need * sizeof(unsigned long long)));
        db->partial_tc_cap = need;
    }
    for (int q0 = 0; q0 < nq; q0 += maxq) {
        const int n = std::min(maxq, nq - q0);
        db->launches += launch_expand_pm1(d_queries + (size_t)q0 * 32, n, knn_tc_padded_queries(n), db->d_qe, db->stream);
        int nparts = 0;
        const int l = launch_knn2_tc(db->d_qe, n, db->d_dbe, db->nrows, db->row_offset, db->sm_count, db->d_partial_tc, &nparts,
                                     merge_launch, db->stream, db->err);
        if (!l) return ORBX_E_CUDA;
        launch_merge(db->d_partial_tc, nparts, n, d_packed_out + (size_t)q0 * 2, db->stream);
        db->launches += l + 1;
        DB_TRY(db, cudaGetLastError());
    }
    return ORBX_OK;
}

// FP4 tensor-core path: same flow as query_device_tensor on the E2M1 expansion
static int query_device_fp4(orbx_db *db, const uint8_t *d_queries, int nq, unsigned long long *d_packed_out) {
    if (!db->d_dbe4) {
        const long long rows_pad = knn_fp4_padded_rows(db->nrows);
        DB_TRY(db, cudaMalloc((void **)&db->d_dbe4, (size_t)rows_pad * 128));
        db->launches += launch_expand_fp4(db->d_rows, db->nrows, rows_pad, db->d_dbe4, db->stream);
        DB_TRY(db, cudaGetLastError());
    }
    const int maxq = knn_fp4_max_queries();
    const int qcap = knn_fp4_padded_queries(std::min(nq, maxq));
    if (qcap > db->qe4_cap) {
        if (db->d_qe4) cudaFree(db->d_qe4);
        db->d_qe4 = nullptr; db->qe4_cap = 0;
        DB_TRY(db, cudaMalloc((void **)&db->d_qe4, (size_t)qcap * 128));
        db->qe4_cap = qcap;
    }
    const size_t need = (size_t)(db->sm_count + 1) * std::min(nq, maxq) * 2 + (size_t)maxq / 2;   // + one u32 per query: the thresholds the CTAs share
    if (need > db->partial_tc_cap) {
        if (db->d_partial_tc) cudaFree(db->d_partial_tc);
        db->d_partial_tc = nullptr; db->partial_tc_cap = 0;
        DB_TRY(db, cudaMalloc((void **)&db->d_partial_tc, need * sizeof(unsigned long long)));
        db->partial_tc_cap = need;
    }
    for (int q0 = 0; q0 < nq; q0 += maxq) {
        const int n = std::min(maxq, nq - q0);
        db->launches += launch_expand_fp4(d_queries + (size_t)q0 * 32, n, knn_fp4_padded_queries(n), db->d_qe4, db->stream);
        int nparts = 0;
        const int l = launch_knn2_fp4(db->d_qe4, n, db->d_dbe4, db->nrows, db->row_offset, db->sm_count, db->d_partial_tc, &nparts, merge_launch,
                                      db->stream, db->err);
        if (!l) return ORBX_E_CUDA;
        launch_merge(db->d_partial_tc, nparts, n, d_packed_out + (size_t)q0 * 2, db->stream);
        db->launches += l + 1;
        DB_TRY(db, cudaGetLastError());
    }
    return ORBX_OK;
}

int orbx_knn2_query_device(orbx_db *db, const uint8_t *d_queries, int nq, unsigned long long *d_packed_out) {
    if (!db || !d_queries || !d_packed_out || nq < 1) return ORBX_E_INVALID;
    if (((uintptr_t)d_queries & 15) != 0) { db->err = "device queries must be 16-byte aligned"; return ORBX_E_INVALID; }
    DB_TRY(db, cudaSetDevice(db->device));
    if (db->backend == ORBX_KNN_TENSOR && db->nrows > 0) return query_device_tensor(db, d_queries, nq, d_packed_out);
    if (db->backend == ORBX_KNN_TENSOR_FP4 && db->nrows > 0) return query_device_fp4(db, d_queries, nq, d_packed_out);
    int rc = db_reserve(db, nq);
    if (rc) return rc;
    dim3 grid((nq + KQ_THREADS * KQ_QPT - 1) / (KQ_THREADS * KQ_QPT), db->nchunks);
    k_knn2<<<grid, KQ_THREADS, 0, db->stream>>>(reinterpret_cast<const uint4 *>(db->d_rows), db->nrows, db->row_offset,
                                                reinterpret_cast<const uint4 *>(d_queries), nq, db->rows_per_chunk, db->d_partial);
    launch_merge(db->d_partial, db->nchunks, nq, d_packed_out, db->stream);
    db->launches += 2;
    DB_TRY(db, cudaGetLastError());
    return ORBX_OK;
}

int orbx_knn2_merge_device(orbx_db *db, const unsigned long long *d_partials, int nparts, int nq, unsigned long long *d_packed_out) {
    if (!db || !d_partials || !d_packed_out || nparts < 1 || nq < 1) return ORBX_E_INVALID;
    DB_TRY(db, cudaSetDevice(db->device));
    launch_merge(d_partials, nparts, nq, d_packed_out, db->stream);
    db->launches += 1;
    DB_TRY(db, cudaGetLastError());
    return ORBX_OK;
}

int orbx_knn2_query(orbx_db *db, const uint8_t *queries, int nq, int32_t *idx_out, int32_t *dist_out) {
    if (!db || !queries || !idx_out || !dist_out || nq < 1) return ORBX_E_INVALID;
    DB_TRY(db, cudaSetDevice(db->device));
    int rc = db_reserve(db, nq);
    if (rc) return rc;
    DB_TRY(db, cudaMemcpyAsync(db->d_q, queries, (size_t)nq * 32, cudaMemcpyHostToDevice, db->stream));
    if ((rc = orbx_knn2_query_device(db, db->d_q, nq, db->d_out))) return rc;
    DB_TRY(db, cudaMemcpyAsync(db->h_out, db->d_out, (size_t)nq * 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, db->stream));
    DB_TRY(db, cudaStreamSynchronize(db->stream));
    for (int i = 0; i < 2 * nq; i++) {
        const unsigned long long k = db->h_out[i];
        if (k == NONE64) { idx_out[i] = -1; dist_out[i] = -1; }
        else { idx_out[i] = (int32_t)(k & 0xFFFFFFFFull); dist_out[i] = (int32_t)(k >> 32); }
    }
    return ORBX_OK;
}

}  // extern "C"

// ---- row-sharded kNN over NCCL ---------------------------------------------------------------------------------------------
// NCCL is bound at run time: the entry points used here have had the same C signatures since NCCL 2.0 (nccl.h: ncclGetUniqueId,
// ncclCommInitRank, ncclCommDestroy, ncclAllGather, ncclGetErrorString; ncclUniqueId = 128 opaque bytes passed by value,
// ncclUint64 = 5).
namespace {
struct NcclId { char internal[ORBX_NCCL_ID_BYTES]; };
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(NcclId *) = nullptr;
    int (*CommInitRank)(void **, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string why;
};
NcclApi *nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.AllGather ? &api : nullptr;
    tried = true;
    const char *env = getenv("ORBX_NCCL_LIB");
    if (env && *env) api.lib = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    if (!api.lib) api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // the copy the process already uses (torch)
    if (!api.lib) api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!api.lib) { api.why = std::string("cannot load libnccl.so.2 (set ORBX_NCCL_LIB): ") + (dlerror() ? dlerror() : ""); return nullptr; }
    api.GetUniqueId = (int (*)(NcclId *))dlsym(api.lib, "ncclGetUniqueId");
    api.CommInitRank = (int (*)(void **, int, NcclId, int))dlsym(api.lib, "ncclCommInitRank");
    api.CommDestroy = (int (*)(void *))dlsym(api.lib, "ncclCommDestroy");
    api.GetErrorString = (const char *(*)(int))dlsym(api.lib, "ncclGetErrorString");
    api.AllGather = (int (*)(const void *, void *, size_t, int, void *, cudaStream_t))dlsym(api.lib, "ncclAllGather");
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather) { api.AllGather = nullptr; api.why = "libnccl.so.2 lacks the expected symbols"; return nullptr; }
    return &api;
}
thread_local std::string g_comm_error;
std::string nccl_err(NcclApi *a, const char *what, int rc) { return std::string(what) + ": " + (a->GetErrorString ? a->GetErrorString(rc) : "nccl error"); }
}  // namespace

struct orbx_comm {
    void *comm = nullptr;
    int rank = 0, nranks = 1, device = -1;
    bool owned = false;
    mutable std::string err;
};

extern "C" {

int orbx_comm_unique_id(uint8_t *id_out) {
    if (!id_out) { g_comm_error = "null argument"; return ORBX_E_INVALID; }
    NcclApi *a = nccl_api();
    if (!a) { g_comm_error = "NCCL unavailable: cannot load libnccl.so.2 (set ORBX_NCCL_LIB to its path)"; return ORBX_E_CUDA; }
    NcclId id;
    const int rc = a->GetUniqueId(&id);
    if (rc) { g_comm_error = nccl_err(a, "ncclGetUniqueId", rc); return ORBX_E_CUDA; }
    std::memcpy(id_out, id.internal, ORBX_NCCL_ID_BYTES);
    return ORBX_OK;
}

int orbx_comm_create(int device, int rank, int nranks, const uint8_t *id, orbx_comm **out) {
    if (!out || !id || nranks < 1 || rank < 0 || rank >= nranks) { g_comm_error = "bad argument"; return ORBX_E_INVALID; }
    *out = nullptr;
    NcclApi *a = nccl_api();
    if (!a) { g_comm_error = "NCCL unavailable: cannot load libnccl.so.2 (set ORBX_NCCL_LIB to its path)"; return ORBX_E_CUDA; }
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { g_comm_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e); return ORBX_E_CUDA; }
    NcclId nid;
    std::memcpy(nid.internal, id, ORBX_NCCL_ID_BYTES);
    void *comm = nullptr;
    const int rc = a->CommInitRank(&comm, nranks, nid, rank);
    if (rc) { g_comm_error = nccl_err(a, "ncclCommInitRank", rc); return ORBX_E_CUDA; }
    orbx_comm *c = new (std::nothrow) orbx_comm();
    if (!c) { a->CommDestroy(comm); g_comm_error = "out of host memory"; return ORBX_E_INVALID; }
    c->comm = comm; c->rank = rank; c->nranks = nranks; c->device = device; c->owned = true;
    *out = c;
    return ORBX_OK;
}

int orbx_comm_adopt(void *nccl_comm, int rank, int nranks, orbx_comm **out) {
    if (!out || !nccl_comm || nranks < 1 || rank < 0 || rank >= nranks) { g_comm_error = "bad argument"; return ORBX_E_INVALID; }
    *out = nullptr;
    if (!nccl_api()) { g_comm_error = "NCCL unavailable: cannot load libnccl.so.2 (set ORBX_NCCL_LIB to its path)"; return ORBX_E_CUDA; }
    orbx_comm *c = new (std::nothrow) orbx_comm();
    if (!c) { g_comm_error = "out of host memory"; return ORBX_E_INVALID; }
    c->comm = nccl_comm; c->rank = rank; c->nranks = nranks; c->owned = false;
    *out = c;
    return ORBX_OK;
}

void orbx_comm_destroy(orbx_comm *c) {
    if (!c) return;
    if (c->owned && c->comm) {
        if (c->device >= 0) cudaSetDevice(c->device);
        NcclApi *a = nccl_api();
        if (a) a->CommDestroy(c->comm);
    }
    delete c;
}

const char *orbx_comm_last_error(const orbx_comm *c) { return c ? c->err.c_str() : g_comm_error.c_str(); }

int orbx_knn2_query_sharded_device(orbx_db *db, orbx_comm *comm, const uint8_t *d_queries, int nq, unsigned long long *d_packed_out) {
    if (!db || !comm || !d_queries || !d_packed_out || nq < 1) return ORBX_E_INVALID;
    NcclApi *a = nccl_api();
    if (!a) { db->err = "NCCL unavailable"; return ORBX_E_CUDA; }
    DB_TRY(db, cudaSetDevice(db->device));
    const size_t need = (size_t)nq * 2;
    if (need > db->sh_cap) {
        if (db->d_sh_local) cudaFree(db->d_sh_local);
        if (db->d_sh_gather) cudaFree(db->d_sh_gather);
        db->d_sh_local = db->d_sh_gather = nullptr; db->sh_cap = 0;
        DB_TRY(db, cudaMalloc((void **)&db->d_sh_local, need * sizeof(unsigned long long)));
        DB_TRY(db, cudaMalloc((void **)&db->d_sh_gather, need * sizeof(unsigned long long) * 64));   // room for 64 ranks
        db->sh_cap = need;
    }
    if (comm->nranks > 64) { db->err = "more than 64 ranks"; return ORBX_E_CAPACITY; }
    int rc = orbx_knn2_query_device(db, d_queries, nq, db->d_sh_local);
    if (rc) return rc;
    if (comm->nranks == 1) return orbx_knn2_merge_device(db, db->d_sh_local, 1, nq, d_packed_out);
    const int nrc = a->AllGather(db->d_sh_local, db->d_sh_gather, need, 5 /* ncclUint64 */, comm->comm, db->stream);   // rank-major [ranks][nq][2]
    if (nrc) { db->err = nccl_err(a, "ncclAllGather", nrc); comm->err = db->err; return ORBX_E_CUDA; }
    return orbx_knn2_merge_device(db, db->d_sh_gather, comm->nranks, nq, d_packed_out);
}

int orbx_knn2_query_sharded(orbx_db *db, orbx_comm *comm, const uint8_t *queries, int nq, int32_t *idx_out, int32_t *dist_out) {
    if (!db || !comm || !queries || !idx_out || !dist_out || nq < 1) return ORBX_E_INVALID;
    DB_TRY(db, cudaSetDevice(db->device));
    int rc = db_reserve(db, nq);
    if (rc) return rc;
    DB_TRY(db, cudaMemcpyAsync(db->d_q, queries, (size_t)nq * 32, cudaMemcpyHostToDevice, db->stream));
    if ((rc = orbx_knn2_query_sharded_device(db, comm, db->d_q, nq, db->d_out))) return rc;
    DB_TRY(db, cudaMemcpyAsync(db->h_out, db->d_out, (size_t)nq * 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, db->stream));
    DB_TRY(db, cudaStreamSynchronize(db->stream));
    for (int i = 0; i < 2 * nq; i++) {
        const unsigned long long k = db->h_out[i];
        if (k == NONE64) { idx_out[i] = -1; dist_out[i] = -1; }
        else { idx_out[i] = (int32_t)(k & 0xFFFFFFFFull); dist_out[i] = (int32_t)(k >> 32); }
    }
    return ORBX_OK;
}

}  // extern "C"

// ---- entry points that live on the extractor handle --------------------------------------------------------------
// (the handle type is opaque here; only its stream/device/launch counter are needed, passed through small accessors)
namespace orbx {
int match_distance_batch(int device, cudaStream_t stream, const uint8_t *a, const uint8_t *b, int n, int32_t *dist_out, std::string &err,
                         long long &launches) {
    if (n <= 0) return ORBX_OK;
    uint8_t *d_a = nullptr, *d_b = nullptr; int32_t *d_o = nullptr;
    cudaError_t e = cudaSetDevice(device);
    auto cleanup = [&]() { if (d_a) cudaFree(d_a); if (d_b) cudaFree(d_b); if (d_o) cudaFree(d_o); };
#define M_TRY(expr) do { e = (expr); if (e != cudaSuccess) { err = std::string(#expr) + ": " + cudaGetErrorString(e); cleanup(); return ORBX_E_CUDA; } } while (0)
    M_TRY(cudaMalloc((void **)&d_a, (size_t)n * 32)); M_TRY(cudaMalloc((void **)&d_b, (size_t)n * 32)); M_TRY(cudaMalloc((void **)&d_o, (size_t)n * 4));
    M_TRY(cudaMemcpyAsync(d_a, a, (size_t)n * 32, cudaMemcpyHostToDevice, stream));
    M_TRY(cudaMemcpyAsync(d_b, b, (size_t)n * 32, cudaMemcpyHostToDevice, stream));
    k_distance_batch<<<(n + 255) / 256, 256, 0, stream>>>(reinterpret_cast<const uint4 *>(d_a), reinterpret_cast<const uint4 *>(d_b), n, d_o);
    launches++;
    M_TRY(cudaGetLastError());
    M_TRY(cudaMemcpyAsync(dist_out, d_o, (size_t)n * 4, cudaMemcpyDeviceToHost, stream));
    M_TRY(cudaStreamSynchronize(stream));
    cleanup();
    return ORBX_OK;
}

int match_windowed(int device, cudaStream_t stream, const uint8_t *q_desc, const float *q_uvr, const int32_t *q_levels, int nq,
                   const orbx_keypoint *t_kp, const uint8_t *t_desc, int nt, const float *bounds4, int32_t *best_idx,
                   int32_t *best_dist, int32_t *second_idx, int32_t *second_dist, std::string &err, long long &launches) {
    if (nq <= 0) return ORBX_OK;
    if (nt >= (1 << 24)) { err = "too many train keypoints"; return ORBX_E_INVALID; }
    cudaError_t e = cudaSetDevice(device);
    std::vector<void *> allocs;
    auto cleanup = [&]() { for (void *p : allocs) cudaFree(p); };
    auto dalloc = [&](size_t bytes) -> void * { void *p = nullptr; if (cudaMalloc(&p, std::max<size_t>(bytes, 256)) != cudaSuccess) return nullptr; allocs.push_back(p); return p; };
    uint8_t *d_qd = (uint8_t *)dalloc((size_t)nq * 32); float *d_uvr = (float *)dalloc((size_t)nq * 12); int32_t *d_lev = (int32_t *)dalloc((size_t)nq * 8);
    KeypointRec *d_tkp = (KeypointRec *)dalloc((size_t)std::max(nt, 1) * sizeof(KeypointRec)); uint8_t *d_td = (uint8_t *)dalloc((size_t)std::max(nt, 1) * 32);
    int32_t *d_out = (int32_t *)dalloc((size_t)nq * 16);
    if (!d_qd || !d_uvr || !d_lev || !d_tkp || !d_td || !d_out) { err = "cudaMalloc failed"; cleanup(); return ORBX_E_CUDA; }
    M_TRY(cudaMemcpyAsync(d_qd, q_desc, (size_t)nq * 32, cudaMemcpyHostToDevice, stream));
    M_TRY(cudaMemcpyAsync(d_uvr, q_uvr, (size_t)nq * 12, cudaMemcpyHostToDevice, stream));
    M_TRY(cudaMemcpyAsync(d_lev, q_levels, (size_t)nq * 8, cudaMemcpyHostToDevice, stream));
    if (nt > 0) {
        M_TRY(cudaMemcpyAsync(d_tkp, t_kp, (size_t)nt * sizeof(KeypointRec), cudaMemcpyHostToDevice, stream));
        M_TRY(cudaMemcpyAsync(d_td, t_desc, (size_t)nt * 32, cudaMemcpyHostToDevice, stream));
    }
    const float minX = bounds4[0], minY = bounds4[1], maxX = bounds4[2], maxY = bounds4[3];
    const float invW = (float)GRID_COLS / (maxX - minX), invH = (float)GRID_ROWS / (maxY - minY);
    k_match_windowed<<<(nq + 7) / 8, 256, 0, stream>>>(reinterpret_cast<const uint4 *>(d_qd), d_uvr, d_lev, nq, d_tkp,
                                                       reinterpret_cast<const uint4 *>(d_td), nt, minX, minY, invW, invH,
                                                       d_out, d_out + nq, d_out + 2 * nq, d_out + 3 * nq);
    launches++;
    M_TRY(cudaGetLastError());
    M_TRY(cudaMemcpyAsync(best_idx, d_out, (size_t)nq * 4, cudaMemcpyDeviceToHost, stream));
    M_TRY(cudaMemcpyAsync(best_dist, d_out + nq, (size_t)nq * 4, cudaMemcpyDeviceToHost, stream));
    M_TRY(cudaMemcpyAsync(second_idx, d_out + 2 * nq, (size_t)nq * 4, cudaMemcpyDeviceToHost, stream));
    M_TRY(cudaMemcpyAsync(second_dist, d_out + 3 * nq, (size_t)nq * 4, cudaMemcpyDeviceToHost, stream));
    M_TRY(cudaStreamSynchronize(stream));
    cleanup();
#undef M_TRY
    return ORBX_OK;
}
}  // namespace orbx

namespace orbx {
int match_windowed_grid_batch_device(cudaStream_t stream, int npairs, const int32_t *pair_q, const int32_t *pair_t, int cap, const uint8_t *d_q_desc,
                                     const float *d_q_uvr, const int32_t *d_q_levels, const int32_t *d_n, const KeypointRec *d_t_kp,
                                     const uint8_t *d_t_desc, const int32_t *d_cell_start, const int32_t *d_cell_items, const float *bounds4,
                                     int32_t *d_best_idx, int32_t *d_best_dist, int32_t *d_second_idx, int32_t *d_second_dist) {
    int launches = 0;
    const float minX = bounds4[0], minY = bounds4[1], maxX = bounds4[2], maxY = bounds4[3];
    const float invW = (float)GRID_COLS / (maxX - minX), invH = (float)GRID_ROWS / (maxY - minY);
    for (int p0 = 0; p0 < npairs; p0 += kMaxMatchPairs) {
        const int np = std::min(kMaxMatchPairs, npairs - p0);
        MatchPairs prs;
        for (int i = 0; i < kMaxMatchPairs; i++) { prs.qf[i] = i < np ? pair_q[p0 + i] : 0; prs.tf[i] = i < np ? pair_t[p0 + i] : 0; }
        k_match_windowed_grid_batch<<<dim3((cap + 7) / 8, np), 256, 0, stream>>>(prs, cap, reinterpret_cast<const uint4 *>(d_q_desc), d_q_uvr, d_q_levels, d_n,
                                                                                d_t_kp, reinterpret_cast<const uint4 *>(d_t_desc), d_cell_start,
                                                                                d_cell_items, minX, minY, invW, invH, d_best_idx, d_best_dist,
                                                                                d_second_idx, d_second_dist);
        launches++;
    }
    return launches;
}
}  // namespace orbx

namespace orbx {
int match_windowed_grid_device(cudaStream_t stream, const uint8_t *d_q_desc, const float *d_q_uvr, const int32_t *d_q_levels, int nq,
                               const KeypointRec *d_t_kp, const uint8_t *d_t_desc, const int32_t *d_cell_start, const int32_t *d_cell_items,
                               const float *bounds4, int32_t *d_best_idx, int32_t *d_best_dist, int32_t *d_second_idx, int32_t *d_second_dist) {
    if (nq <= 0) return 0;
    const float minX = bounds4[0], minY = bounds4[1], maxX = bounds4[2], maxY = bounds4[3];
    const float invW = (float)GRID_COLS / (maxX - minX), invH = (float)GRID_ROWS / (maxY - minY);
    k_match_windowed_grid<<<(nq + 7) / 8, 256, 0, stream>>>(reinterpret_cast<const uint4 *>(d_q_desc), d_q_uvr, d_q_levels, nq, d_t_kp,
                                                            reinterpret_cast<const uint4 *>(d_t_desc), d_cell_start, d_cell_items, minX, minY,
                                                            invW, invH, d_best_idx, d_best_dist, d_second_idx, d_second_dist);
    return 1;
}
}  // namespace orbx

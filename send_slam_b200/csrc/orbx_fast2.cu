// K2 (second formulation)  FAST-9/16 score + per-cell NMS + per-cell threshold fallback on an UNPACKED pair plane.
//
// Replaces the cell loop of UPSTREAM ORB-SLAM3 ORBextractor::ComputeKeyPointsOctTree (cv::FAST(iniTh) / cv::FAST(minTh) per cell +
// 3x3 non-max suppression; built per slam_backends/orb_slam_3/CMakeLists.txt:52).  Same scores, same NMS rule, same
// candidate sets as k_fast_tma / k_fast_cells (orbx_kernels.cu); what changes is where the instructions go.
//
// ncu on k_fast_tma (profiles/k_fast_tma_r01_summary.txt): the integer ALU pipe is 68 % busy and only half of its work is the
// u16x2 min/max network -- the rest are the funnel shifts that cut ring slices out of packed bytes, the <<8 copies for the even
// pixels, and the byte-lane juggling of the NMS.  Here
//   * the ROI arrives by TMA as bytes (one box per chunk of cell rows) and is unpacked ONCE into a plane U[row][m] = (px[m], px[m+1])
//     of clean u16x2 lanes, one word per pixel column (5 ALU ops per 4 pixels);
//   * a work item is one pixel PAIR: each of its 16 ring words and the centre is a single LDS at an immediate offset from one
//     address (the plane pitch P is a template constant) -- no shifts, no byte permutes in the scoring loop;
//   * lanes walk the items column-major (consecutive lanes = consecutive rows) and P is odd, so every LDS of the loop is
//     bank-conflict free whatever the cell width;
//   * clean lanes are valid fp16 subnormals (0..255 x 2^-24), for which a - b, relu and + are exact: (min, max) of a pair is
//     t = relu(a - b); min = a - t; max = b + t as three HFMA2 on the FMA pipe instead of VIMNMX (ALU) + 2 IMAD -- NRELU of the 16
//     pair extrema per item take that route, balancing the two pipes;
//   * NMS works on the same pair words: column maxima of three rows, one PRMT each for the left / right neighbour lanes,
//     strict compare by max/xor; flagged lanes are rare (about 1.5 % of the pixels), so the hand-over sits behind a branch.
// Cells taller than the chunk height are processed in row chunks by the same warp (rolling score tile: the last two scored rows
// move to the top), so the per-warp shared memory is set by the chunk height, not by the tallest cell of the pyramid.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "orbx_dev.h"
#include "orbx_tma.cuh"

namespace orbx {

extern std::mutex g_attr_mutex;

namespace {

__device__ __forceinline__ uint32_t umax2(uint32_t a, uint32_t b) { return __vmaxu2(a, b); }
__device__ __forceinline__ uint32_t umin2(uint32_t a, uint32_t b) { return __vminu2(a, b); }
__device__ __forceinline__ uint32_t umax3(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_u16x2(a, b, c); }
__device__ __forceinline__ uint32_t umin3(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_u16x2(a, b, c); }

// fp16x2 arithmetic on the FMA pipe; operands are u16 lanes 0..255 read as fp16 subnormals, every result is exact
__device__ __forceinline__ uint32_t h_relu_sub(uint32_t a, uint32_t b) {   // max(a - b, 0)
    uint32_t d;
    asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(b), "r"(0xBC00BC00u), "r"(a));
    return d;
}
__device__ __forceinline__ uint32_t h_sub(uint32_t a, uint32_t t) {        // a - t
    uint32_t d;
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(t), "r"(0xBC00BC00u), "r"(a));
    return d;
}
__device__ __forceinline__ uint32_t h_add(uint32_t b, uint32_t t) {        // b + t
    uint32_t d;
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(t), "r"(0x3C003C00u), "r"(b));
    return d;
}

__constant__ uint32_t c_one2 = 1u;   // a multiplier ptxas cannot fold: keeps a + b - min as two IMADs (FMA pipe)

// (min, max) of the u16x2 lanes of a, b.  ROUTE 0: three HFMA2 (FMA pipe only); 1: VIMNMX + two IMAD (one ALU-pipe op, two on the
// FMA pipe); 2: two VIMNMX (fewest instructions, ALU pipe only).
template <int ROUTE>
__device__ __forceinline__ void minmax2(uint32_t a, uint32_t b, uint32_t &mn, uint32_t &mx) {
    if (ROUTE == 0) {
        const uint32_t t = h_relu_sub(a, b);
        mn = h_sub(a, t); mx = h_add(b, t);
    } else if (ROUTE == 1) {
        const uint32_t one = c_one2;
        mn = umin2(a, b);
        mx = (a + b * one) - mn * one;
    } else {
        mn = umin2(a, b); mx = umax2(a, b);
    }
}
template <int NRELU, int NIMAD, int IDX>
__device__ __forceinline__ void minmax_at(uint32_t a, uint32_t b, uint32_t &mn, uint32_t &mx) {
    minmax2<(IDX < NRELU) ? 0 : (IDX < NRELU + NIMAD) ? 1 : 2>(a, b, mn, mx);
}

// FAST-9/16 score of two pixels.  r[k]: ring pixel k of both pixels (clean u16x2 lanes), v: the centres.
//   min over arcs of (max over arc) = min_i max3(Q2x[i], Q2x[i+2], min(r[2i], r[2i+9])), Q2x = maxima of 4 ring pixels; same for
//   max over arcs of (min over arc).  score = max(v - minArcMax, maxArcMin - v, 0)  (= cv's cornerScore + 1).
// The 16 (min, max) pairs of an item are spread over the routes: the first NRELU by HFMA2, the next NIMAD by VIMNMX + IMAD, the
// rest by two VIMNMX; pairs 2k are the Q pairs, pairs 2k+1 the (r[2i], r[2i+9]) pairs, so that every setting mixes both kinds.
template <int NRELU, int NIMAD>
__device__ __forceinline__ uint32_t fast_score_pair(const uint32_t (&r)[16], uint32_t v) {
    uint32_t qx[8], qn[8];
#define QP(j) minmax_at<NRELU, NIMAD, 2 * (j)>(r[2 * (j) + 1], r[(2 * (j) + 2) & 15], qn[j], qx[j])
    QP(0); QP(1); QP(2); QP(3); QP(4); QP(5); QP(6); QP(7);
#undef QP
    uint32_t q2x[8], q2n[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        q2x[i] = umax2(qx[i], qx[(i + 1) & 7]);
        q2n[i] = umin2(qn[i], qn[(i + 1) & 7]);
    }
    uint32_t fx[8], fn[8];
#define FP(i)                                                                                       \
    {                                                                                               \
        uint32_t mn, mx;                                                                            \
        minmax_at<NRELU, NIMAD, 2 * (i) + 1>(r[2 * (i)], r[(2 * (i) + 9) & 15], mn, mx);            \
        fx[i] = umax3(q2x[i], q2x[((i) + 2) & 7], mn);                                              \
        fn[i] = umin3(q2n[i], q2n[((i) + 2) & 7], mx);                                              \
    }
    FP(0) FP(1) FP(2) FP(3) FP(4) FP(5) FP(6) FP(7)
#undef FP
    const uint32_t min_arc_max = umin3(umin3(fx[0], fx[1], fx[2]), umin3(fx[3], fx[4], fx[5]), umin2(fx[6], fx[7]));
    const uint32_t max_arc_min = umax3(umax3(fn[0], fn[1], fn[2]), umax3(fn[3], fn[4], fn[5]), umax2(fn[6], fn[7]));
    return umax2(h_relu_sub(v, min_arc_max), h_relu_sub(max_arc_min, v));
}

// score of the pair whose plane address is `ad` (address of U[row][2j + sh]; the centre sits 3 rows and 3 words further on)
template <int P, int NRELU, int NIMAD>
__device__ __forceinline__ uint32_t score_at(uint32_t ad) {
    uint32_t r[16], v;
#define ULD(dst, dx, dy) asm volatile("ld.shared.u32 %0, [%1 + %2];" : "=r"(dst) : "r"(ad), "n"((((dy) + 3) * P + (dx) + 3) * 4))
    ULD(r[0], 0, 3);   ULD(r[1], 1, 3);   ULD(r[2], 2, 2);    ULD(r[3], 3, 1);
    ULD(r[4], 3, 0);   ULD(r[5], 3, -1);  ULD(r[6], 2, -2);   ULD(r[7], 1, -3);
    ULD(r[8], 0, -3);  ULD(r[9], -1, -3); ULD(r[10], -2, -2); ULD(r[11], -3, -1);
    ULD(r[12], -3, 0); ULD(r[13], -3, 1); ULD(r[14], -2, 2);  ULD(r[15], -1, 3);
    ULD(v, 0, 0);
#undef ULD
    return fast_score_pair<NRELU, NIMAD>(r, v);
}

__device__ __forceinline__ uint2 lds64(uint32_t saddr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr));
    return v;
}

}  // namespace

struct Fast2Params {
    CUtensorMap map[kMaxLevels];        // level plane [frames][h][w], box = box_w x box_h x 1
    int box_w[kMaxLevels];              // bytes per staged row (multiple of 16)
    uint32_t *cand[kMaxLevels];         // [frames][cand_cap]
    int *cand_count[kMaxLevels];        // [frames]
    int cand_cap[kMaxLevels];
    int box_h;                          // staged rows per chunk = ch + 6 (same for every level)
    int ch;                             // tested rows per chunk
    int stage_bytes;                    // per-warp shared memory: TMA stage (multiple of 128) ...
    int u_bytes;                        // ... pair plane, (ch + 6) rows x P words ...
    int sc_pitch, sc_bytes;             // ... score tile (pitch in words, even; ch + 4 rows) ...
    int bits_pitch, bits_words;         // ... bitmap of the NMS survivors of a chunk (words per tile row, total) ...
    int list_cap;                       // ... candidate list (u32 entries) ...
    int warp_bytes;                     // ... total per warp (multiple of 128)
};

// P: pair-plane pitch in words (odd).  One WARP owns one cell at a time and walks (cell, frame) items with a grid-wide stride.
template <int P, int NRELU, int NIMAD>
__global__ void __launch_bounds__(128, 4) k_fast_pairs(const __grid_constant__ Fast2Params Q, const CellRect *__restrict__ cells, int ncells,
                                                       int total, int ini_th, int min_th, int f0, int *__restrict__ overflow) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    uint8_t *wsm = smem + (size_t)warp * Q.warp_bytes;
    const uint32_t *s_roi = reinterpret_cast<const uint32_t *>(wsm);
    uint32_t *s_u = reinterpret_cast<uint32_t *>(wsm + Q.stage_bytes);
    uint32_t *s_sc = reinterpret_cast<uint32_t *>(wsm + Q.stage_bytes + Q.u_bytes);
    uint32_t *s_bits = reinterpret_cast<uint32_t *>(wsm + Q.stage_bytes + Q.u_bytes + Q.sc_bytes);
    uint32_t *s_list = s_bits + Q.bits_words;
    uint64_t *s_full = reinterpret_cast<uint64_t *>(s_list + Q.list_cap);
    int *s_cnt = reinterpret_cast<int *>(s_full + 1);                      // [above iniTh, rest]
    const int SP = Q.sc_pitch, lcap = Q.list_cap, CH = Q.ch, BW = Q.bits_pitch;

    if (lane == 0) {
        tma_mbar_init(s_full, 1);
        tma_mbar_fence_init();
        s_cnt[0] = 0; s_cnt[1] = 0;
    }
    for (int i = lane; i < Q.bits_words; i += 32) s_bits[i] = 0;
    __syncwarp();
    const int stride = gridDim.x * nwarps;
    int item = blockIdx.x * nwarps + warp;
    if (item >= total) return;
    auto load_cell = [&](int it, int &fi) -> CellRect {
        CellRect c; c.level = -1; c.x0 = c.y0 = c.x1 = c.y1 = 0; fi = 0;
        if (it < total) { fi = it / ncells; c = cells[it - fi * ncells]; }
        return c;
    };
    auto issue = [&](const CellRect &c, int fi, int row0) {   // lane 0 only: rows y0 + row0 .. of the cell's ROI
        tma_mbar_expect_tx(s_full, (uint32_t)(Q.box_w[c.level] * Q.box_h));
        tma_load_3d(wsm, &Q.map[c.level], c.x0 & ~15, c.y0 + row0, f0 + fi, s_full);
    };
    int fi_cur, fi_nxt;
    CellRect cur = load_cell(item, fi_cur), nxt = load_cell(item + stride, fi_nxt);
    if (lane == 0) issue(cur, fi_cur, 0);
    const uint32_t thr2 = (uint32_t)min_th | ((uint32_t)min_th << 16);
    const uint32_t sc_base = tma_smem_u32(s_sc);
    uint32_t waits = 0;
    for (; item < total; item += stride) {
        int fi_nn;
        const CellRect nn = load_cell(item + 2 * stride, fi_nn);            // descriptor prefetch, consumed next iteration
        const int f = f0 + fi_cur, level = cur.level;
        const int rpw = Q.box_w[level] >> 2;                                // staged row pitch in words
        const int iw = cur.x1 - cur.x0 - 6, ih = cur.y1 - cur.y0 - 6;       // tested pixels
        const int np = (iw + 1) >> 1;                                       // pixel pairs per tested row
        const int off = cur.x0 & 15, abw = off >> 2, sh = off & 3;          // ROI column c sits at plane index c + sh
        const int K8 = (2 * np + 8 + sh + 7) >> 3;                          // 8-pixel groups per row to unpack (plane indices 0 .. 8 K8 - 1)
        const int nch = (ih + CH - 1) / CH, cr = (ih + nch - 1) / nch;      // balanced row chunks
        // zero ring of the score tile: word 0 and words np + 1, np + 2 of every tile row, tile row 1 (= tested row -1)
        for (int t = lane; t < cr + 4; t += 32) { s_sc[t * SP] = 0; s_sc[t * SP + np + 1] = 0; s_sc[t * SP + np + 2] = 0; }
        for (int i = lane; i < np + 3; i += 32) s_sc[SP + i] = 0;
        int prev_rows = 0;
        for (int c = 0; c < nch; c++) {
            const int a = c * cr, b = min(ih, a + cr), nrows = b - a;
            const bool last = c == nch - 1;
            tma_mbar_wait(s_full, waits & 1u);
            waits++;
            // ---- unpack ROI rows a .. b+5 (staged rows 0 .. nrows+5): U[row][m] = (byte m, byte m+1) as u16x2, 8 bytes per item
            {
                const int nit = (nrows + 6) * K8;
                const uint32_t inv = 65536u / (uint32_t)K8 + 1u;
                for (int it = lane; it < nit; it += 32) {
                    const int row = (int)(((uint32_t)it * inv) >> 16), k = it - row * K8;
                    const uint32_t *src = s_roi + row * rpw + abw + 2 * k;
                    const uint32_t w0 = src[0], w1 = src[1], w2 = src[2];
                    uint32_t *dst = s_u + row * P + 8 * k;
                    dst[0] = __byte_perm(w0, 0u, 0x4140);
                    dst[1] = __byte_perm(w0, 0u, 0x4241);
                    dst[2] = __byte_perm(w0, 0u, 0x4342);
                    dst[3] = __byte_perm(__funnelshift_r(w0, w1, 24), 0u, 0x4140);   // (byte 3 of w0, byte 0 of w1)
                    dst[4] = __byte_perm(w1, 0u, 0x4140);
                    dst[5] = __byte_perm(w1, 0u, 0x4241);
                    dst[6] = __byte_perm(w1, 0u, 0x4342);
                    dst[7] = __byte_perm(__funnelshift_r(w1, w2, 24), 0u, 0x4140);
                }
            }
            __syncwarp();
            // the stage is free again: request the next chunk (of this cell, or the first one of the warp's next cell)
            if (lane == 0) {
                if (!last) issue(cur, fi_cur, a + cr);
                else if (nxt.level >= 0) issue(nxt, fi_nxt, 0);
            }
            // ---- rolling score tile: tile row t <-> tested row a - 2 + t; the last two scored rows of the previous chunk move up
            if (c > 0) {
                for (int i = lane; i < 2 * SP; i += 32) {
                    const int rr = i >= SP ? 1 : 0, col = i - rr * SP;
                    if (col < np + 3) s_sc[rr * SP + col] = s_sc[(prev_rows + rr) * SP + col];
                }
                __syncwarp();
            }
            // ---- scores: work item = pair column j x rows (rr, rr + H); lanes walk rows first (odd plane pitch: conflict-free LDS)
            {
                const int H = (nrows + 1) >> 1;
                const int nit = np * H;
                const uint32_t inv = 65536u / (uint32_t)H + 1u;
                const uint32_t ubase = tma_smem_u32(s_u) + (uint32_t)sh * 4u;   // the centre of item (j, rr) sits at +(3 P + 3) words
                const uint32_t hp4 = (uint32_t)(min(H, nrows - 1) * P) * 4u;     // second row of an item (clamped into the staged rows)
                for (int it = lane; it < nit; it += 32) {
                    const int j = (int)(((uint32_t)it * inv) >> 16), rr = it - j * H;
                    const uint32_t ad = ubase + (uint32_t)(rr * P + 2 * j) * 4u;
                    uint32_t *so = s_sc + (rr + 2) * SP + j + 1;
                    // both rows unconditionally (two independent dependency graphs per thread); the second row of the last item of
                    // a column may lie behind the chunk: it is computed on the staged halo rows and not stored
                    const uint32_t s0 = score_at<P, NRELU, NIMAD>(ad), s1 = score_at<P, NRELU, NIMAD>(ad + hp4);
                    so[0] = s0;
                    if (rr + H < nrows) so[H * SP] = s1;
                }
            }
            __syncwarp();
            if (iw & 1) for (int t = 2 + lane; t < nrows + 2; t += 32) s_sc[t * SP + np] &= 0xFFFFu;   // second pixel of the last pair: outside
            if (last) for (int i = lane; i < np + 3; i += 32) s_sc[(nrows + 2) * SP + i] = 0;          // tested row ih
            __syncwarp();
            // ---- NMS (strict 8-neighbour maximum inside the cell, neighbours outside count 0) for tested rows first .. lastrow:
            //      work item = 2 pair columns x 2 rows; the 4 x 4 words around them arrive as eight 64-bit loads
            {
                const int first = c == 0 ? 0 : a - 1, lastrow = last ? ih - 1 : b - 2;
                const int QH = (lastrow - first + 2) >> 1, JH = (np + 1) >> 1;
                const int nit = JH * QH;
                const uint32_t inv = 65536u / (uint32_t)JH + 1u;
                for (int it = lane; it < nit; it += 32) {
                    const int q = (int)(((uint32_t)it * inv) >> 16), jj = it - q * JH;
                    const int rt = first + 2 * q;                     // tested rows rt, rt + 1; pairs 2 jj, 2 jj + 1
                    const uint32_t pa = sc_base + (uint32_t)((rt - a + 1) * SP + 2 * jj) * 4u;   // tile row above rt, tile column of pair 2 jj - 1
                    uint32_t w[4][4];
#pragma unroll
                    for (int r4 = 0; r4 < 4; r4++) {
                        const uint2 lo = lds64(pa + (uint32_t)(r4 * SP) * 4u), hi = lds64(pa + (uint32_t)(r4 * SP + 2) * 4u);
                        w[r4][0] = lo.x; w[r4][1] = lo.y; w[r4][2] = hi.x; w[r4][3] = hi.y;
                    }
                    if ((w[1][1] | w[1][2] | w[2][1] | w[2][2]) == 0) continue;
#pragma unroll
                    for (int rr = 0; rr < 2; rr++) {
                        uint32_t T[4];
#pragma unroll
                        for (int cc = 0; cc < 4; cc++) T[cc] = umax3(w[rr][cc], w[rr + 1][cc], w[rr + 2][cc]);
                        const uint32_t x01 = __byte_perm(T[0], T[1], 0x5432), x12 = __byte_perm(T[1], T[2], 0x5432), x23 = __byte_perm(T[2], T[3], 0x5432);
                        const uint32_t m0 = umax3(x01, x12, umax3(w[rr][1], w[rr + 2][1], thr2));
                        const uint32_t m1 = umax3(x12, x23, umax3(w[rr][2], w[rr + 2][2], thr2));
                        // strict maximum above minTh <=> NOT (ring maximum >= score) per lane: VIMNMX with predicate outputs
                        bool h0, l0, h1, l1;
                        __vibmax_u16x2(m0, w[rr + 1][1], &h0, &l0);
                        __vibmax_u16x2(m1, w[rr + 1][2], &h1, &l1);
                        uint32_t mask = (l0 ? 0u : 1u) | (h0 ? 0u : 2u) | (l1 ? 0u : 4u) | (h1 ? 0u : 8u);
                        if (rr == 1 && rt + 1 > lastrow) mask = 0;
                        // survivors are rare: one shared-memory OR per (item, row); the lists are filled once per chunk below
                        if (mask) atomicOr(&s_bits[(rt + rr - a + 2) * BW + (jj >> 3)], mask << (4 * (jj & 7)));
                    }
                }
            }
            __syncwarp();
            // ---- survivors of this chunk -> the cell's two lists (above iniTh from the front, the rest from the back)
            {
                const int nw = (nrows + 3) * BW;
                for (int wi = lane; wi < nw; wi += 32) {
                    uint32_t bits = s_bits[wi];
                    if (!bits) continue;
                    s_bits[wi] = 0;
                    const int t = wi / BW, cw = wi - t * BW;
                    const uint32_t *row = s_sc + t * SP + 1 + 16 * cw;
                    while (bits) {
                        const int bpos = __ffs(bits) - 1;
                        bits &= bits - 1;
                        const int col = 32 * cw + bpos;
                        if (col >= iw) continue;                      // the masked second pixel of an odd-width row's last pair
                        const uint32_t sc = (row[bpos >> 1] >> (16 * (bpos & 1))) & 0xFFFFu;
                        const uint32_t e = ((uint32_t)(t + a - 2) << 20) | ((uint32_t)col << 8) | sc;
                        if ((int)sc > ini_th) s_list[atomicAdd(&s_cnt[0], 1)] = e;
                        else s_list[lcap - 1 - atomicAdd(&s_cnt[1], 1)] = e;
                    }
                }
            }
            __syncwarp();
            prev_rows = nrows;
        }
        // ---- per-cell fallback: if any corner passes iniThFAST keep only those, else keep everything above minThFAST
        {
            const int n_ini = s_cnt[0], n_low = s_cnt[1];
            const int n = n_ini > 0 ? n_ini : n_low;
            const uint32_t *src = s_list + (n_ini > 0 ? 0 : lcap - n_low);
            if (n > 0) {
                int base = 0;
                if (lane == 0) base = atomicAdd(Q.cand_count[level] + f, n);
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                const int cap = Q.cand_cap[level];
                uint32_t *__restrict__ out = Q.cand[level] + (size_t)f * cap;
                // relative coordinates (x - 16, y - 16) of the reference's vToDistributeKeys entries
                const uint32_t xrel = (uint32_t)(cur.x0 + 3 - kMinBorder), yrel = (uint32_t)(cur.y0 + 3 - kMinBorder);
                for (int i = lane; i < n; i += 32) {
                    const uint32_t e = src[i];
                    const uint32_t vout = (((e >> 20) + yrel) << 20) | ((((e >> 8) & 0xFFFu) + xrel) << 8) | ((e & 0xFFu) - 1u);
                    if (base + i < cap) out[base + i] = vout; else *overflow = 1;
                }
            }
            __syncwarp();
            if (lane == 0) { s_cnt[0] = 0; s_cnt[1] = 0; }
            __syncwarp();
        }
        cur = nxt; fi_cur = fi_nxt; nxt = nn; fi_nxt = fi_nn;
    }
}

// ---- host side -------------------------------------------------------------------------------------------------------------
static int env_int(const char *name, int dflt) { const char *e = getenv(name); return e ? atoi(e) : dflt; }

typedef void (*Fast2Kernel)(const Fast2Params, const CellRect *, int, int, int, int, int, int *);
struct Fast2Variant { int pitch; Fast2Kernel fn; };

// instantiated route mixes (NRELU, NIMAD; the remaining pairs take two VIMNMX), selected by ORBX_FAST_MIX = index
template <int NRELU, int NIMAD>
static const Fast2Variant *fast2_variants() {
    static const Fast2Variant v[] = {{49, k_fast_pairs<49, NRELU, NIMAD>}, {57, k_fast_pairs<57, NRELU, NIMAD>}, {73, k_fast_pairs<73, NRELU, NIMAD>},
                                     {89, k_fast_pairs<89, NRELU, NIMAD>}, {0, nullptr}};
    return v;
}
static const Fast2Variant *fast2_mix(int mix) {
    switch (mix) {
        case 0: return fast2_variants<0, 0>();     // all pairs by two VIMNMX
        case 1: return fast2_variants<4, 0>();
        case 2: return fast2_variants<8, 0>();
        case 3: return fast2_variants<12, 0>();
        case 4: return fast2_variants<16, 0>();    // all pairs by HFMA2
        case 5: return fast2_variants<0, 16>();    // all pairs by VIMNMX + 2 IMAD (the first formulation's route)
        case 6: return fast2_variants<8, 8>();
        default: return fast2_variants<6, 0>();
    }
}

int fast2_pick_pitch(int min_words) {
    for (const Fast2Variant *v = fast2_variants<8, 0>(); v->fn; v++) if (v->pitch >= min_words) return v->pitch;
    return 0;
}

int launch_fast2(const LevelDev *h_levels, const CellRect *d_cells, int ncells, int f0, int batch, int ini_th, int min_th, int *d_overflow,
                 cudaStream_t stream, const Fast2Tma *tma, int sm_count) {
    static_assert(sizeof(Fast2Tma::map) == sizeof(Fast2Params::map), "tensor map storage mismatch");
    Fast2Params Q;
    memcpy(Q.map, tma->map, sizeof(Q.map));
    for (int l = 0; l < kMaxLevels; l++) {
        Q.box_w[l] = tma->box_w[l];
        Q.cand[l] = h_levels[l].cand; Q.cand_count[l] = h_levels[l].cand_count; Q.cand_cap[l] = h_levels[l].cand_cap;
    }
    Q.box_h = tma->box_h; Q.ch = tma->ch;
    Q.stage_bytes = tma->stage_bytes;
    Q.u_bytes = (tma->box_h * tma->pitch * 4 + 15) / 16 * 16;
    Q.sc_pitch = (tma->max_np + 3 + 1) & ~1;
    if ((Q.sc_pitch & 3) == 0) Q.sc_pitch += 2;      // pitch = 2 mod 4: the score stores of a column-major warp spread over 16 banks
    Q.sc_bytes = (Q.sc_pitch * (tma->rows + 4) * 4 + 15) / 16 * 16;
    Q.bits_pitch = (tma->max_np * 2 + 31) / 32;
    Q.bits_words = Q.bits_pitch * (tma->rows + 4);
    Q.list_cap = (((tma->max_iw + 1) / 2) * ((tma->max_ih + 1) / 2) + 8 + 3) / 4 * 4;
    Q.warp_bytes = (Q.stage_bytes + Q.u_bytes + Q.sc_bytes + Q.bits_words * 4 + Q.list_cap * 4 + 8 + 2 * 4 + 127) / 128 * 128;
    static const int mix = env_int("ORBX_FAST_MIX", 2);
    const Fast2Variant *vars = fast2_mix(mix);
    Fast2Kernel fn = nullptr;
    for (const Fast2Variant *v = vars; v->fn; v++) if (v->pitch == tma->pitch) fn = v->fn;
    if (!fn) return 0;
    static const int warps_env = env_int("ORBX_FAST_WARPS", 2);
    const int nwarps = warps_env >= 1 && warps_env <= 4 ? warps_env : 2;
    const size_t smem = (size_t)nwarps * Q.warp_bytes;
    // function attributes are per (device, kernel instantiation); handles of different frame sizes use different instantiations
    struct Conf { int dev; const void *fn; size_t smem; int threads; int per_sm; };
    static std::vector<Conf> confs;
    const int dv = current_device_slot();
    int per_sm = 0;
    {
        std::lock_guard<std::mutex> lock(g_attr_mutex);
        for (const Conf &cf : confs) if (cf.dev == dv && cf.fn == (const void *)fn && cf.smem == smem && cf.threads == nwarps * 32) per_sm = cf.per_sm;
        if (!per_sm) {
            size_t raised = 0;   // the attribute only ever goes up: another handle may still launch this instantiation with more
            for (const Conf &cf : confs) if (cf.dev == dv && cf.fn == (const void *)fn) raised = cf.smem > raised ? cf.smem : raised;
            if (smem > raised) cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            int n = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, nwarps * 32, smem) != cudaSuccess || n < 1) n = 1;
            confs.push_back(Conf{dv, (const void *)fn, smem, nwarps * 32, n});
            per_sm = n;
        }
    }
    const int total = ncells * batch;
    const int want = (total + nwarps - 1) / nwarps;
    static const int cap_env = env_int("ORBX_FAST_CTAS", 0);
    const int resident = cap_env > 0 && cap_env < per_sm ? cap_env : per_sm;
    const int grid = want < sm_count * resident ? want : sm_count * resident;
    fn<<<grid, nwarps * 32, smem, stream>>>(Q, d_cells, ncells, total, ini_th, min_th, f0, d_overflow);
    return 1;
}

}  // namespace orbx

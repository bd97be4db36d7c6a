// Extraction kernels of the B200-native ORB front end (sm_100a).
//
// Stage            replaces (UPSTREAM ORB-SLAM3 src/ORBextractor.cc, built per slam_backends/orb_slam_3/CMakeLists.txt:52)
//   k_resize       ComputePyramid: cv::resize(INTER_LINEAR), level l-1 -> l               
//   k_blur         GaussianBlur(7x7, sigma 2, REFLECT_101) of every level                  (A.2)
//   k_fast_cells   ComputeKeyPointsOctTree cell loop: cv::FAST(iniTh) / per-cell minTh fallback + 3x3 NMS (A.3)
//   k_octree       DistributeOctTree + DivideNode                                          (C.1)
//   k_finalize     tail of operator(): level -> image scaling, mono/stereo slot order      (C.1)
//   k_describe     computeOrientation (IC_Angle + fastAtan2) and computeOrbDescriptor      (A.4, A.5)
//
// All arithmetic is integer or individually rounded fp32 (no FMA contraction) so results are bit-identical to the
// CPU checker used by the tests.  Everything here is HBM/L2-resident byte work: no tensor cores on purpose.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <mutex>

#include "orbx_dev.h"
#include "orbx_tma.cuh"

namespace orbx {

// The level table travels in the kernel parameter bank (constant memory), so that no kernel starts with dependent global loads.
struct LevelTable { LevelDev lv[kMaxLevels]; };
static LevelTable make_table(const LevelDev *h_levels) {
    LevelTable t;
    memcpy(t.lv, h_levels, sizeof(t.lv));
    return t;
}

__device__ int8_t g_pattern[1024];
__constant__ int c_umax[16];

static const int8_t h_pattern[1024] = {
#include "orb_pattern.inc"
};

// kernel attributes are set once per device; handles on several host threads may reach their first launch at the same time
std::mutex g_attr_mutex;

int current_device_slot() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess) d = 0;
    return d & (kMaxDevices - 1);
}

void upload_constants() {
    // umax for HALF_PATCH_SIZE 15 (ctor of ORBextractor; identical for every parameter set)
    static const int umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
    cudaMemcpyToSymbol(g_pattern, h_pattern, sizeof(h_pattern));
    cudaMemcpyToSymbol(c_umax, umax, sizeof(umax));
}

// ---------------------------------------------------------------------------------------------------------------
// K0  colour -> gray (cv::cvtColor RGB2GRAY / BGR2GRAY / RGBA2GRAY / BGRA2GRAY, 8U fixed point): what UPSTREAM
//     Tracking::GrabImageMonocular does on the CPU before the Frame is built.  gray = (c0*ch0 + c1*ch1 + c2*ch2 + half) >> shift
//     with (R, G, B) weights (9798, 19235, 3735) >> 15 (cv2 4.13, bit-exact) or (4899, 9617, 1868) >> 14 (older builds).
//     4 pixels per thread: 3 or 4 aligned words in, one word out (level-0 plane of the workspace).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_gray(const uint8_t *__restrict__ src, size_t src_fstride, int src_pitch, int channels,
                                              uint32_t c0, uint32_t c1, uint32_t c2, uint32_t half, int shift,
                                              uint8_t *__restrict__ dst, size_t dst_fstride, int dst_pitch, int w, int f0) {
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (x0 >= w) return;
    const int y = blockIdx.y, f = f0 + blockIdx.z;
    const uint8_t *row = src + (size_t)f * src_fstride + (size_t)y * src_pitch + (size_t)x0 * channels;
    uint32_t px[4][3];
    const bool words = ((reinterpret_cast<uintptr_t>(row) & 3) == 0) && x0 + 3 < w;
    if (words && channels == 3) {
        const uint32_t *q = reinterpret_cast<const uint32_t *>(row);
        const uint32_t a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);      // bytes 0..11 = 4 x (ch0, ch1, ch2)
        px[0][0] = a & 0xFF;         px[0][1] = (a >> 8) & 0xFF;  px[0][2] = (a >> 16) & 0xFF;
        px[1][0] = a >> 24;          px[1][1] = b & 0xFF;         px[1][2] = (b >> 8) & 0xFF;
        px[2][0] = (b >> 16) & 0xFF; px[2][1] = b >> 24;          px[2][2] = c & 0xFF;
        px[3][0] = (c >> 8) & 0xFF;  px[3][1] = (c >> 16) & 0xFF; px[3][2] = c >> 24;
    } else if (words) {
        const uint32_t *q = reinterpret_cast<const uint32_t *>(row);
#pragma unroll
        for (int i = 0; i < 4; i++) { const uint32_t a = __ldg(q + i); px[i][0] = a & 0xFF; px[i][1] = (a >> 8) & 0xFF; px[i][2] = (a >> 16) & 0xFF; }
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int xi = min(x0 + i, w - 1) - x0;
            px[i][0] = row[xi * channels]; px[i][1] = row[xi * channels + 1]; px[i][2] = row[xi * channels + 2];
        }
    }
    uint32_t out = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) out |= ((px[i][0] * c0 + px[i][1] * c1 + px[i][2] * c2 + half) >> shift) << (8 * i);
    *reinterpret_cast<uint32_t *>(dst + (size_t)f * dst_fstride + (size_t)y * dst_pitch + x0) = out;
}

// format: 1 RGB, 2 BGR, 3 RGBA, 4 BGRA (ORBX_FMT_*); shift 15 or 14.  dst rows must be 4-byte aligned with a padded pitch.
int launch_gray(const uint8_t *d_src, size_t src_fstride, int src_pitch, int format, int shift, uint8_t *d_dst, size_t dst_fstride,
                int dst_pitch, int w, int h, int f0, int batch, cudaStream_t stream) {
    const int channels = format >= 3 ? 4 : 3;
    const bool rgb = format == 1 || format == 3;
    const uint32_t r = shift == 15 ? 9798u : 4899u, g = shift == 15 ? 19235u : 9617u, b = shift == 15 ? 3735u : 1868u;
    dim3 grid((w + 4 * 128 - 1) / (4 * 128), h, batch);
    k_gray<<<grid, 128, 0, stream>>>(d_src, src_fstride, src_pitch, channels, rgb ? r : b, g, rgb ? b : r, 1u << (shift - 1), shift, d_dst,
                                     dst_fstride, dst_pitch, w, f0);
    return 1;
}

// ---------------------------------------------------------------------------------------------------------------
// K1  pyramid resize (cv::resize INTER_LINEAR, 8UC1).  Taps are tabulated on the host (orbx_plan.cpp).
//     k_resize: 4 destination pixels x 2 destination rows per thread.  The <= 8 source bytes a 4-pixel group needs are
//     fetched as three aligned words per source row and realigned once; each pixel then costs one PRMT (its two
//     neighbouring source bytes) and one IDP.2A (s0*c0 + s1*c1) per source row.
//     k_resize_generic: byte-wise fallback for source planes that are not 4-byte aligned (caller-owned level 0).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_resize_generic(const __grid_constant__ LevelTable T, int level, int f0) {
    const LevelDev *lv = T.lv;   // level table in the kernel parameter (constant) bank: no dependent global loads
    const LevelDev &D = lv[level];
    const LevelDev &S = lv[level - 1];
    const int dx0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (dx0 >= D.w) return;
    const int dy = blockIdx.y, f = f0 + blockIdx.z;
    const ResizeTap ty = D.ytap[dy];
    const uint8_t *__restrict__ s0 = S.img + (size_t)f * S.img_fstride + (size_t)ty.ofs * S.pitch;
    const uint8_t *__restrict__ s1 = S.img + (size_t)f * S.img_fstride + (size_t)ty.ofs1 * S.pitch;
    uint32_t packed = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int dx = dx0 + i;
        if (dx < D.w) {
            const ResizeTap tx = D.xtap[dx];
            const int r0 = (int)s0[tx.ofs] * tx.c0 + (int)s0[tx.ofs1] * tx.c1;
            const int r1 = (int)s1[tx.ofs] * tx.c0 + (int)s1[tx.ofs1] * tx.c1;
            int v = ((((int)ty.c0 * (r0 >> 4)) >> 16) + (((int)ty.c1 * (r1 >> 4)) >> 16) + 2) >> 2;
            v = min(max(v, 0), 255);
            packed |= (uint32_t)v << (8 * i);
        }
    }
    *reinterpret_cast<uint32_t *>(D.img + (size_t)f * D.img_fstride + (size_t)dy * D.pitch + dx0) = packed;
}

// xpack[dx] = ofs << 16 | c1 (c0 = 2048 - c1), padded to a multiple of 4 entries.
// Work item = (4-pixel group, block of RR destination rows), flattened so that every lane of every warp has work on
// the narrow upper levels too; all 6*RR source words of an item are requested before the first one is used (the kernel
// is latency-bound otherwise: a level is a few MB).
constexpr int RR = 4;
// PADDED: the source plane is one of the workspace's own (>= 16 bytes of slack behind every row it can be asked for), so the three
// words of a row are read at constant offsets from one address; otherwise (level 0 aliasing caller memory) word offsets are clamped.
template <bool PADDED>
__global__ void __launch_bounds__(128) k_resize(const __grid_constant__ LevelTable T, int level, int f0, int ngx, int nitems) {
    const LevelDev *lv = T.lv;   // level table in the kernel parameter (constant) bank: no dependent global loads
    const LevelDev &D = lv[level];
    const LevelDev &S = lv[level - 1];
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= nitems) return;
    const int ry = item / ngx, dx0 = (item - ry * ngx) * 4, dy0 = ry * RR, f = f0 + blockIdx.y;
    const uint4 tp = __ldg(reinterpret_cast<const uint4 *>(D.xpack + dx0));
    const uint32_t t[4] = {tp.x, tp.y, tp.z, tp.w};
    const int ofs0 = (int)(t[0] >> 16);
    const int base = ofs0 & ~3;
    const uint32_t mis8 = 8u * (uint32_t)(ofs0 & 3);
    const int lastw = (S.w - 1) & ~3;                       // last word that still starts inside the row
    const int o0 = base, o1 = min(base + 4, lastw), o2 = min(base + 8, lastw);
    const uint8_t *__restrict__ sbase = S.img + (size_t)f * S.img_fstride;
    const int spitch = S.pitch, dh = D.h;
    ResizeTap ty[RR];
    uint32_t w[RR][2][3];
#pragma unroll
    for (int rr = 0; rr < RR; rr++) {    // one 8-byte load per tap record
        const uint2 q = __ldg(reinterpret_cast<const uint2 *>(D.ytap + min(dy0 + rr, dh - 1)));
        ty[rr].ofs = (int16_t)(q.x & 0xFFFF); ty[rr].c0 = (int16_t)(q.x >> 16); ty[rr].c1 = (int16_t)(q.y & 0xFFFF); ty[rr].ofs1 = (int16_t)(q.y >> 16);
    }
#pragma unroll
    for (int rr = 0; rr < RR; rr++)
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const uint8_t *row = sbase + (size_t)(k ? ty[rr].ofs1 : ty[rr].ofs) * spitch;
            if (PADDED) {
                const uint32_t *q = reinterpret_cast<const uint32_t *>(row + o0);
                w[rr][k][0] = __ldg(q); w[rr][k][1] = __ldg(q + 1); w[rr][k][2] = __ldg(q + 2);
            } else {
                w[rr][k][0] = __ldg(reinterpret_cast<const uint32_t *>(row + o0));
                w[rr][k][1] = __ldg(reinterpret_cast<const uint32_t *>(row + o1));
                w[rr][k][2] = __ldg(reinterpret_cast<const uint32_t *>(row + o2));
            }
        }
    uint32_t sel[4], coef[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t d = (t[i] >> 16) - (uint32_t)ofs0;   // 0..6: byte offset of the first tap inside the window
        sel[i] = d | ((d + 1) << 4);                          // PRMT: bytes d, d+1 -> result bytes 0, 1
        const uint32_t c1 = t[i] & 0xFFFFu;
        coef[i] = (c1 << 16) | (2048u - c1);                  // IDP.2A: lo16 * byte0 + hi16 * byte1
    }
    uint8_t *__restrict__ drow = D.img + (size_t)f * D.img_fstride + (size_t)dy0 * D.pitch + dx0;
#pragma unroll
    for (int rr = 0; rr < RR; rr++) {
        uint32_t r[2][4];
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const uint32_t a = __funnelshift_r(w[rr][k][0], w[rr][k][1], mis8), b = __funnelshift_r(w[rr][k][1], w[rr][k][2], mis8);   // bytes ofs0 .. ofs0+7
#pragma unroll
            for (int i = 0; i < 4; i++) r[k][i] = __dp2a_lo(coef[i], __byte_perm(a, b, sel[i]), 0u);
        }
        // (c * x) >> 16 as the high word of (c << 16) * x: one IMAD.HI on the FMA pipe instead of IMAD + shift (c >= 0 here)
        const uint32_t c0s = (uint32_t)ty[rr].c0 << 16, c1s = (uint32_t)ty[rr].c1 << 16;
        uint32_t packed = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            // <= ((2048 * 32640) >> 16) + 2 >> 2 = 255 because c0 + c1 = 2048 and both are non-negative on this path: no clamp
            const uint32_t v = (__umulhi(c0s, r[0][i] >> 4) + __umulhi(c1s, r[1][i] >> 4) + 2u) >> 2;
            packed |= v << (8 * i);
        }
        if (dy0 + rr < dh) *reinterpret_cast<uint32_t *>(drow + (size_t)rr * D.pitch) = packed;
    }
}

int launch_resize(const LevelDev *h_levels, int level, int f0, int batch, cudaStream_t stream) {
    const LevelDev &D = h_levels[level];
    const LevelDev &S = h_levels[level - 1];
    // the word loads of k_resize may touch up to 3 bytes behind the last pixel of a row: fine inside the workspace's padded planes,
    // not behind the last row of caller-owned memory, so an unpadded source qualifies only when its rows end on a word boundary
    const bool aligned = ((reinterpret_cast<uintptr_t>(S.img) | (uintptr_t)S.pitch | (uintptr_t)S.img_fstride) & 3) == 0 && D.xpack != nullptr &&
                         (S.padded || (S.w & 3) == 0);
    if (aligned) {
        const int ngx = (D.w + 3) / 4, nitems = ngx * ((D.h + RR - 1) / RR);
        dim3 grid((nitems + 127) / 128, batch);
        if (S.padded) k_resize<true><<<grid, 128, 0, stream>>>(make_table(h_levels), level, f0, ngx, nitems);
        else k_resize<false><<<grid, 128, 0, stream>>>(make_table(h_levels), level, f0, ngx, nitems);
    } else {
        dim3 grid((D.w + 4 * 128 - 1) / (4 * 128), D.h, batch);
        k_resize_generic<<<grid, 128, 0, stream>>>(make_table(h_levels), level, f0);
    }
    return 1;
}

// ---------------------------------------------------------------------------------------------------------------
// K1c  fused pyramid ("cone" tiling).  The per-level kernels above are seven dependent, latency-bound launches per frame range; with four
//      ranges per batch that is 28 small launches per step, and they cost the overlapped step more than their own duration (measured with
//      stages switched off, tools/whatif.py: 350 -> 238 us per 64-frame step without the pyramid).  Here ONE CTA carries a tile of the
//      source level down up to four levels in shared memory: the source region arrives as one TMA box, level l's region is computed from
//      level l-1's region (same item arithmetic as k_resize: 4 pixels x 4 rows, taps from the host tables) into the other half of a
//      ping-pong buffer, and the part of it this tile owns is written to the level's plane.  Regions overlap between neighbouring tiles
//      by what the bilinear taps of the levels above need (orbx_plan.cpp: build_cone_plan; ~15 % redundant pixels), so CTAs never wait
//      for each other.  Every pixel, owned or redundant, is the same arithmetic on the same inputs: bit-identical to the per-level path.
// ---------------------------------------------------------------------------------------------------------------
struct ConeParams {
    CUtensorMap map;                    // source level plane [frames][h][w], box = box_w x box_h x 1
    int src, nl;                        // source level; entries per tile (levels src .. src + nl - 1)
    int box_w, box_h, pitch, buf0_bytes;
};

__global__ void __launch_bounds__(256) k_pyramid_cone(const __grid_constant__ ConeParams C, const __grid_constant__ LevelTable T,
                                                      const ConeLevel *__restrict__ tiles, int f0) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ __align__(16) uint32_t s_xt[256];        // x taps of the current region (rw <= 256 - 16)
    __shared__ __align__(8) uint2 s_yt[272];            // y taps of its rows, padded to item blocks of four
    uint8_t *buf0 = smem, *buf1 = smem + C.buf0_bytes;
    const ConeLevel *R = tiles + (size_t)blockIdx.x * C.nl;
    const int f = f0 + blockIdx.y, tid = threadIdx.x;
    const ConeLevel box = R[0];
    if (box.rw <= 0 || box.rh <= 0) return;             // a tile without pixels (grids finer than a tiny level); uniform for the CTA
    if (tid == 0) { tma_mbar_init(&bar, 1); tma_mbar_fence_init(); }
    __syncthreads();
    if (tid == 0) {
        tma_mbar_expect_tx(&bar, (uint32_t)(C.box_w * C.box_h));
        tma_load_3d(buf0, &C.map, box.rx0, box.ry0, f, &bar);
    }
    tma_mbar_wait(&bar, 0);
    const uint8_t *src = buf0;
    int spitch = C.box_w, sx0 = box.rx0, sy0 = box.ry0;
    for (int k = 1; k < C.nl; k++) {
        const LevelDev &D = T.lv[C.src + k];
        const ConeLevel r = R[k];
        uint8_t *dst = (k & 1) ? buf1 : buf0;
        const int ngx = r.rw >> 2, nry = (r.rh + 3) >> 2, nitems = ngx * nry, dh = D.h, dpitch = C.pitch;
        uint8_t *__restrict__ gplane = D.img + (size_t)f * D.img_fstride;
        // the region's taps go to shared memory once (the items of a level would otherwise start with dependent global loads)
        for (int i = tid; i < r.rw; i += 256) s_xt[i] = __ldg(D.xpack + r.rx0 + i);
        for (int i = tid; i < 4 * nry; i += 256) s_yt[i] = __ldg(reinterpret_cast<const uint2 *>(D.ytap + min(r.ry0 + i, dh - 1)));
        __syncthreads();
        const uint32_t inv_ngx = 65536u / (uint32_t)ngx + 1u;
        for (int it = tid; it < nitems; it += 256) {
            const int ry = (int)(((uint32_t)it * inv_ngx) >> 16), gxi = it - ry * ngx;
            const int dx0 = r.rx0 + 4 * gxi, dy0 = r.ry0 + 4 * ry;
            const uint4 tp = *reinterpret_cast<const uint4 *>(s_xt + 4 * gxi);
            const uint32_t t[4] = {tp.x, tp.y, tp.z, tp.w};
            const int ofs0 = (int)(t[0] >> 16);
            const uint32_t mis8 = 8u * (uint32_t)(ofs0 & 3);
            const uint8_t *col = src + ((ofs0 & ~3) - sx0);
            ResizeTap ty[RR];
            uint32_t w[RR][2][3];
#pragma unroll
            for (int rr = 0; rr < RR; rr++) {
                const uint2 q = s_yt[4 * ry + rr];
                ty[rr].ofs = (int16_t)(q.x & 0xFFFF); ty[rr].c0 = (int16_t)(q.x >> 16); ty[rr].c1 = (int16_t)(q.y & 0xFFFF); ty[rr].ofs1 = (int16_t)(q.y >> 16);
            }
#pragma unroll
            for (int rr = 0; rr < RR; rr++)
#pragma unroll
                for (int k2 = 0; k2 < 2; k2++) {
                    const uint32_t *q = reinterpret_cast<const uint32_t *>(col + ((k2 ? ty[rr].ofs1 : ty[rr].ofs) - sy0) * spitch);
                    w[rr][k2][0] = q[0]; w[rr][k2][1] = q[1]; w[rr][k2][2] = q[2];
                }
            uint32_t sel[4], coef[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t d = (t[i] >> 16) - (uint32_t)ofs0;   // 0..6: byte offset of the first tap inside the window
                sel[i] = d | ((d + 1) << 4);                          // PRMT: bytes d, d+1 -> result bytes 0, 1
                const uint32_t c1 = t[i] & 0xFFFFu;
                coef[i] = (c1 << 16) | (2048u - c1);                  // IDP.2A: lo16 * byte0 + hi16 * byte1
            }
            const bool own_x = dx0 >= r.ox0 && dx0 < r.ox1;
            uint32_t *drow = reinterpret_cast<uint32_t *>(dst + (dy0 - r.ry0) * dpitch + (dx0 - r.rx0));
#pragma unroll
            for (int rr = 0; rr < RR; rr++) {
                uint32_t rv[2][4];
#pragma unroll
                for (int k2 = 0; k2 < 2; k2++) {
                    const uint32_t a = __funnelshift_r(w[rr][k2][0], w[rr][k2][1], mis8), b = __funnelshift_r(w[rr][k2][1], w[rr][k2][2], mis8);
#pragma unroll
                    for (int i = 0; i < 4; i++) rv[k2][i] = __dp2a_lo(coef[i], __byte_perm(a, b, sel[i]), 0u);
                }
                const uint32_t c0s = (uint32_t)ty[rr].c0 << 16, c1s = (uint32_t)ty[rr].c1 << 16;
                uint32_t packed = 0;
#pragma unroll
                for (int i = 0; i < 4; i++) packed |= ((__umulhi(c0s, rv[0][i] >> 4) + __umulhi(c1s, rv[1][i] >> 4) + 2u) >> 2) << (8 * i);
                drow[rr * (dpitch >> 2)] = packed;
                const int dy = dy0 + rr;
                if (own_x && dy >= r.oy0 && dy < r.oy1) *reinterpret_cast<uint32_t *>(gplane + (size_t)dy * D.pitch + dx0) = packed;
            }
        }
        __syncthreads();
        src = dst; spitch = dpitch; sx0 = r.rx0; sy0 = r.ry0;
    }
}

int launch_pyramid_cone(const LevelDev *h_levels, const ConeLaunch &cl, int f0, int batch, cudaStream_t stream) {
    static_assert(sizeof(cl.map) == sizeof(CUtensorMap), "tensor map storage mismatch");
    ConeParams C;
    memcpy(&C.map, cl.map, sizeof(C.map));
    C.src = cl.src; C.nl = cl.nl; C.box_w = cl.box_w; C.box_h = cl.box_h; C.pitch = cl.pitch; C.buf0_bytes = cl.buf0_bytes;
    const size_t smem = (size_t)cl.buf0_bytes + cl.buf1_bytes;
    static size_t configured_[kMaxDevices];
    {
        std::lock_guard<std::mutex> lock(g_attr_mutex);
        size_t &configured = configured_[current_device_slot()];
        if (smem > configured) { cudaFuncSetAttribute(k_pyramid_cone, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); configured = smem; }
    }
    dim3 grid(cl.ntiles, batch);
    k_pyramid_cone<<<grid, 256, smem, stream>>>(C, make_table(h_levels), cl.d_tiles, f0);
    return 1;
}

// ---------------------------------------------------------------------------------------------------------------
// K5  7x7 Gaussian, fixed point [18,34,48,56,48,34,18]/256 per axis, exact 16.16 accumulation, REFLECT_101.
//     Output tile 64 x 112 per CTA (two rounds of items per thread amortise the per-thread set-up; 64 x 56 cost 20 % more
//     instructions).  The halo tile (118 rows x 80 bytes) is staged with word loads; the horizontal pass is
//     two IDP.4A per pixel on byte-aligned word slices (sums <= 65280 fit 16 bits) and leaves its sums packed as
//     (row 2p, row 2p+1) pairs, so that the vertical pass is four IDP.2A per pixel (16-bit sums x 8-bit taps, 32-bit acc).
// ---------------------------------------------------------------------------------------------------------------
constexpr int BTW = kBlurTileW, BTH = kBlurTileH;   // 64 x 112
constexpr int BIN_PITCH = 80;          // bytes per staged input row: image columns x0-4 .. x0+75
constexpr int BROWS = BTH + 6;         // up to 118 staged rows = 59 row pairs
static_assert(BTW == 64 && BTH % 4 == 0, "k_blur's item mapping needs 64-column tiles and 4-row segments");

// BORDER_REFLECT_101 for any index (period 2(n-1)); n == 1 maps everything to 0
__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    if (i < 0) i = -i;
    if (i < n) return i;
    const int p = 2 * (n - 1);
    if (i >= p) i %= p;          // only for images narrower than the staging halo
    return i >= n ? p - i : i;
}

__global__ void __launch_bounds__(256) k_blur(const __grid_constant__ LevelTable T, const BlurTile *__restrict__ tiles, int f0) {
    const LevelDev *lv = T.lv;   // level table in the kernel parameter (constant) bank: no dependent global loads
    __shared__ __align__(16) uint8_t s_in[BROWS * BIN_PITCH];
    __shared__ __align__(16) uint32_t s_h[(BROWS / 2) * BTW];   // [row pair][column]: H(2p, c) | H(2p+1, c) << 16
    const BlurTile t = tiles[blockIdx.x];
    const LevelDev &L = lv[t.level];
    const int f = f0 + blockIdx.y;
    const int x0 = t.tx * BTW, y0 = t.ty * BTH;
    const int w = L.w, h = L.h, pitch = L.pitch;
    const uint8_t *__restrict__ src = L.img + (size_t)f * L.img_fstride;
    // rows of this tile (the last tile of a level is shorter): 4-row output segments, staged rows = 4 * nseg + 6
    const int nseg = (min(BTH, h - y0) + 3) >> 2, srows = 4 * nseg + 6, npairs = srows >> 1;
    // stage rows y0-3 .. y0-3+srows-1, columns x0-4 .. x0+75 (20 words per row): aligned word loads where the word lies
    // inside the image, per-byte BORDER_REFLECT_101 elsewhere; five words per thread are requested before the first store.
    const bool word_ok = ((reinterpret_cast<uintptr_t>(src) | (uintptr_t)pitch) & 3) == 0;
    const bool big = w >= 16 && h >= 16;   // one reflection per row index is enough (staged rows reach < 8 past an edge)
    const int nw = srows * (BIN_PITCH / 4);
    if (big && word_ok && L.padded) {
        // Fast staging (workspace planes of ordinary size): every word is an aligned load at a clamped column of the reflected
        // row - no per-word edge branch, which made whole warps of every border tile (half of all tiles) run the per-byte
        // path.  The three REFLECT_101 columns left of x = 0 and right of x = w - 1 are patched in shared memory afterwards.
        const int lastw = (w - 1) & ~3;
        // thread <-> (row tr of 12, word wc of 20): 240 threads stage 12 rows per step, five steps in flight
        const int tr = threadIdx.x / (BIN_PITCH / 4), wc = threadIdx.x - tr * (BIN_PITCH / 4);
        if (tr < 12) {
            const int cx = min(max(x0 - 4 + 4 * wc, 0), lastw);
            for (int r0 = tr; r0 < srows; r0 += 12 * 5) {
                uint32_t v[5];
#pragma unroll
                for (int k = 0; k < 5; k++) {
                    const int r = r0 + 12 * k;
                    if (r < srows) {
                        const int gy = y0 - 3 + r;
                        const int ry = gy < 0 ? -gy : (gy >= h ? 2 * h - 2 - gy : gy);
                        v[k] = __ldg(reinterpret_cast<const uint32_t *>(src + (size_t)ry * pitch + cx));
                    }
                }
#pragma unroll
                for (int k = 0; k < 5; k++) {
                    const int r = r0 + 12 * k;
                    if (r < srows) reinterpret_cast<uint32_t *>(s_in)[r * (BIN_PITCH / 4) + wc] = v[k];
                }
            }
        }
        const bool left = x0 == 0, right = x0 + 76 > w;      // the staged columns x0-4 .. x0+75 leave the image
        if (left || right) {
            __syncthreads();
            for (int r = threadIdx.x; r < srows; r += 256) {
                uint8_t *row = s_in + r * BIN_PITCH;
                if (left) { row[1] = row[7]; row[2] = row[6]; row[3] = row[5]; }      // x = -3, -2, -1 <- x = 3, 2, 1
                if (right) {
                    const int c = w - x0 + 4;                                           // staged column of image column w
#pragma unroll
                    for (int k = 0; k < 3; k++) if (c + k < BIN_PITCH) row[c + k] = row[c - 2 - k];   // x = w + k <- x = w - 2 - k
                }
            }
        }
    } else
    for (int it0 = threadIdx.x; it0 < nw; it0 += 256 * 5) {
        uint32_t v[5];
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const int it = it0 + 256 * k;
            if (it < nw) {
                const int r = it / (BIN_PITCH / 4), wc = it - r * (BIN_PITCH / 4);
                const int gy = y0 - 3 + r, gx = x0 - 4 + 4 * wc;
                if (word_ok && gx >= 0 && gx + 3 < w && gy >= 0 && gy < h) {
                    v[k] = __ldg(reinterpret_cast<const uint32_t *>(src + (size_t)gy * pitch + gx));
                } else {
                    const uint8_t *row = src + (size_t)reflect101(gy, h) * pitch;
                    v[k] = (uint32_t)row[reflect101(gx, w)] | ((uint32_t)row[reflect101(gx + 1, w)] << 8) |
                           ((uint32_t)row[reflect101(gx + 2, w)] << 16) | ((uint32_t)row[reflect101(gx + 3, w)] << 24);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const int it = it0 + 256 * k;
            if (it < nw) reinterpret_cast<uint32_t *>(s_in)[it] = v[k];
        }
    }
    __syncthreads();
    // horizontal: item = (row pair p, group g of 8 output columns)
    for (int it = threadIdx.x; it < npairs * 8; it += 256) {
        const int p = it >> 3, g = it & 7;
        constexpr uint32_t KA = 18u | (34u << 8) | (48u << 16) | (56u << 24), KB = 48u | (34u << 8) | (18u << 16);
        uint32_t hs[2][8];
#pragma unroll
        for (int rr = 0; rr < 2; rr++) {
            const uint2 a = *reinterpret_cast<const uint2 *>(s_in + (2 * p + rr) * BIN_PITCH + 8 * g);
            const uint2 b = *reinterpret_cast<const uint2 *>(s_in + (2 * p + rr) * BIN_PITCH + 8 * g + 8);
            const uint32_t W[4] = {a.x, a.y, b.x, b.y};
            // U[o] = staged bytes 8g+o .. 8g+o+3; output column 8g+i reads bytes i+1 .. i+7 = U[i+1] (4 taps) + U[i+5] (3 taps)
            uint32_t U[13];
#pragma unroll
            for (int o = 1; o <= 12; o++) U[o] = (o & 3) ? __funnelshift_r(W[o >> 2], W[(o >> 2) + 1], 8 * (o & 3)) : W[o >> 2];
#pragma unroll
            for (int i = 0; i < 8; i++) hs[rr][i] = __dp4a(U[i + 5], KB, __dp4a(U[i + 1], KA, 0u));
        }
        uint32_t o[8];
#pragma unroll
        for (int i = 0; i < 8; i++) o[i] = hs[0][i] | (hs[1][i] << 16);
        uint4 *dst = reinterpret_cast<uint4 *>(s_h + p * BTW + 8 * g);
        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
    }
    __syncthreads();
    // vertical: item = (4-column group cg, segment of 4 output rows).  Output row y reads staged rows y .. y+6: for even y the
    // pairs y/2 .. y/2+3 with taps (18,34)(48,56)(48,34)(18,0), for odd y the pairs (y-1)/2 .. (y-1)/2+3 with taps
    // (0,18)(34,48)(56,48)(34,18).
    for (int it = threadIdx.x; it < 16 * nseg; it += 256) {
        const int cg = it & 15, seg = it >> 4;
        const int gx = x0 + 4 * cg;
        if (gx < w) {
            constexpr uint32_t EA = 18u | (34u << 8) | (48u << 16) | (56u << 24), EB = 48u | (34u << 8) | (18u << 16);
            constexpr uint32_t OA = (18u << 8) | (34u << 16) | (48u << 24), OB = 56u | (48u << 8) | (34u << 16) | (18u << 24);
            uint32_t P[5][4];
#pragma unroll
            for (int k = 0; k < 5; k++) {
                const uint4 q = *reinterpret_cast<const uint4 *>(s_h + (2 * seg + k) * BTW + 4 * cg);
                P[k][0] = q.x; P[k][1] = q.y; P[k][2] = q.z; P[k][3] = q.w;
            }
            uint8_t *__restrict__ dst = L.blur + (size_t)f * L.blur_fstride + gx;
            const int bp = L.blur_pitch;
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int gy = y0 + 4 * seg + r;
                const int pb = r >> 1;
                const uint32_t ka = (r & 1) ? OA : EA, kb = (r & 1) ? OB : EB;
                uint32_t px = 0;
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    uint32_t acc = __dp2a_lo(P[pb][c], ka, 32768u);
                    acc = __dp2a_hi(P[pb + 1][c], ka, acc);
                    acc = __dp2a_lo(P[pb + 2][c], kb, acc);
                    acc = __dp2a_hi(P[pb + 3][c], kb, acc);
                    px |= (acc >> 16) << (8 * c);
                }
                if (gy < h) *reinterpret_cast<uint32_t *>(dst + (size_t)gy * bp) = px;
            }
        }
    }
}

// TMA-staged CUDA-core variant (what the pipeline runs with ORBX_BLUR_TC=0, or when the tensor-core kernel of orbx_blur_tc.cu cannot be
// used; needs every plane to meet the TMA alignment rules and to be at least 16 x 16): the halo
// tile arrives as ONE box of 96 x 118 bytes at (x0 - 16, y0 - 3) -- a TMA box starts on a 16-byte boundary, bytes outside the plane
// arrive as 0 -- and the BORDER_REFLECT_101 rows / columns are patched in shared memory afterwards.  That takes the ~6.6 staging
// instructions per pixel of k_blur off the SM (measured by tools/blur_tma_probe.cu on 64 x 640x480: 23.1 -> 18.8 us); the two passes
// are k_blur's, with the row pairs and the pixel quads packed by byte permutes instead of shift / mask pairs.
constexpr int BTP = 96;                 // staged row pitch: image columns x0-16 .. x0+79
struct BlurTmaParams { CUtensorMap map[kMaxLevels]; };

__global__ void __launch_bounds__(256) k_blur_tma(const __grid_constant__ BlurTmaParams M, const __grid_constant__ LevelTable T,
                                                  const BlurTile *__restrict__ tiles, int f0) {
    const LevelDev *lv = T.lv;
    __shared__ __align__(128) uint8_t s_in[BROWS * BTP];
    __shared__ __align__(16) uint32_t s_h[(BROWS / 2) * BTW];
    __shared__ __align__(8) uint64_t bar;
    const BlurTile t = tiles[blockIdx.x];
    const LevelDev &L = lv[t.level];
    const int f = f0 + blockIdx.y;
    const int x0 = t.tx * BTW, y0 = t.ty * BTH;
    const int w = L.w, h = L.h;
    const int nseg = (min(BTH, h - y0) + 3) >> 2, srows = 4 * nseg + 6, npairs = srows >> 1;
    if (threadIdx.x == 0) { tma_mbar_init(&bar, 1); tma_mbar_fence_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        tma_mbar_expect_tx(&bar, (uint32_t)(BTP * BROWS));
        tma_load_3d(s_in, &M.map[t.level], x0 - 16, y0 - 3, f, &bar);
    }
    tma_mbar_wait(&bar, 0);
    // rows outside the plane <- their reflections (staged row r holds image row y0 - 3 + r)
    const bool top = y0 == 0, bottom = y0 - 3 + srows > h;
    if (top || bottom) {
        for (int i = threadIdx.x; i < 3 * (BTP / 4); i += 256) {
            const int k = i / (BTP / 4), wq = i - k * (BTP / 4);
            uint32_t *S = reinterpret_cast<uint32_t *>(s_in);
            if (top) S[(2 - k) * (BTP / 4) + wq] = S[(4 + k) * (BTP / 4) + wq];                  // rows -1-k <- rows 1+k
            if (bottom) {
                const int r = h + k - (y0 - 3), sr = h - 2 - k - (y0 - 3);                       // row h+k <- row h-2-k
                if (r < srows && sr >= 0) S[r * (BTP / 4) + wq] = S[sr * (BTP / 4) + wq];
            }
        }
        __syncthreads();
    }
    const bool left = x0 == 0, right = x0 + 76 > w;
    if (left || right) {
        for (int r = threadIdx.x; r < srows; r += 256) {
            uint8_t *row = s_in + r * BTP + 12;                                                  // row[c]: image column x0 - 4 + c
            if (left) { row[1] = row[7]; row[2] = row[6]; row[3] = row[5]; }                     // x = -3, -2, -1 <- x = 3, 2, 1
            if (right) {
                const int c = w - x0 + 4;                                                        // staged column of image column w
#pragma unroll
                for (int k = 0; k < 3; k++) if (c + k < BTP - 12) row[c + k] = row[c - 2 - k];   // x = w + k <- x = w - 2 - k
            }
        }
    }
    __syncthreads();
    // horizontal: item = (row pair p, group g of 8 output columns)
    for (int it = threadIdx.x; it < npairs * 8; it += 256) {
        const int p = it >> 3, g = it & 7;
        constexpr uint32_t KA = 18u | (34u << 8) | (48u << 16) | (56u << 24), KB = 48u | (34u << 8) | (18u << 16);
        uint32_t hs[2][8];
#pragma unroll
        for (int rr = 0; rr < 2; rr++) {
            const uint32_t *q = reinterpret_cast<const uint32_t *>(s_in + (2 * p + rr) * BTP + 12 + 8 * g);   // image columns x0-4+8g ..
            const uint32_t W[4] = {q[0], q[1], q[2], q[3]};
            uint32_t U[13];
#pragma unroll
            for (int o = 1; o <= 12; o++) U[o] = (o & 3) ? __funnelshift_r(W[o >> 2], W[(o >> 2) + 1], 8 * (o & 3)) : W[o >> 2];
#pragma unroll
            for (int i = 0; i < 8; i++) hs[rr][i] = __dp4a(U[i + 5], KB, __dp4a(U[i + 1], KA, 0u));
        }
        uint32_t o[8];
#pragma unroll
        for (int i = 0; i < 8; i++) o[i] = __byte_perm(hs[0][i], hs[1][i], 0x5410);   // lo16(row 2p) | lo16(row 2p+1) << 16
        uint4 *d4 = reinterpret_cast<uint4 *>(s_h + p * BTW + 8 * g);
        d4[0] = make_uint4(o[0], o[1], o[2], o[3]);
        d4[1] = make_uint4(o[4], o[5], o[6], o[7]);
    }
    __syncthreads();
    // vertical: item = (4-column group cg, segment of 4 output rows), as in k_blur.  The output address is formed once per item and
    // pinned in a register: left to itself the compiler re-derives it per row from the parameter bank (17 instructions per store).
    uint8_t *dstp = L.blur + (size_t)f * L.blur_fstride + (size_t)y0 * L.blur_pitch + x0;
    int bp = L.blur_pitch;
    asm volatile("" : "+l"(dstp), "+r"(bp));
    for (int it = threadIdx.x; it < 16 * nseg; it += 256) {
        const int cg = it & 15, seg = it >> 4;
        const int gx = x0 + 4 * cg;
        if (gx >= w) continue;
        uint8_t *orow = dstp + (4 * seg) * bp + 4 * cg;
        constexpr uint32_t EA = 18u | (34u << 8) | (48u << 16) | (56u << 24), EB = 48u | (34u << 8) | (18u << 16);
        constexpr uint32_t OA = (18u << 8) | (34u << 16) | (48u << 24), OB = 56u | (48u << 8) | (34u << 16) | (18u << 24);
        uint32_t Pq[5][4];
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const uint4 q = *reinterpret_cast<const uint4 *>(s_h + (2 * seg + k) * BTW + 4 * cg);
            Pq[k][0] = q.x; Pq[k][1] = q.y; Pq[k][2] = q.z; Pq[k][3] = q.w;
        }
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int gy = y0 + 4 * seg + r, pb = r >> 1;
            const uint32_t ka = (r & 1) ? OA : EA, kb = (r & 1) ? OB : EB;
            uint32_t acc[4];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                uint32_t a = __dp2a_lo(Pq[pb][c], ka, 32768u);
                a = __dp2a_hi(Pq[pb + 1][c], ka, a);
                a = __dp2a_lo(Pq[pb + 2][c], kb, a);
                acc[c] = __dp2a_hi(Pq[pb + 3][c], kb, a);
            }
            // byte 2 of every accumulator is the pixel (acc < 2^24): three permutes for four pixels
            const uint32_t px = __byte_perm(__byte_perm(acc[0], acc[1], 0x0062), __byte_perm(acc[2], acc[3], 0x0062), 0x5410);
            if (gy < h) *reinterpret_cast<uint32_t *>(orow + r * bp) = px;
        }
    }
}

int launch_blur(const LevelDev *h_levels, const BlurTile *d_tiles, int ntiles, int f0, int batch, cudaStream_t stream, const BlurTma *tma) {
    if (ntiles <= 0) return 0;
    dim3 grid(ntiles, batch);
    if (tma && tma->ok) {
        static_assert(sizeof(BlurTma::map) == sizeof(BlurTmaParams::map), "tensor map storage mismatch");
        BlurTmaParams M;
        memcpy(M.map, tma->map, sizeof(M.map));
        k_blur_tma<<<grid, 256, 0, stream>>>(M, make_table(h_levels), d_tiles, f0);
        return 1;
    }
    k_blur<<<grid, 256, 0, stream>>>(make_table(h_levels), d_tiles, f0);
    return 1;
}

// ---------------------------------------------------------------------------------------------------------------
// K2  FAST-9/16 score + per-cell NMS + per-cell threshold fallback.  One CTA per (cell, frame).
//
// score(p) = max(0, max over the 16 arcs of 9 ring pixels of max(min_k (v - r_k), min_k (r_k - v)))  (= cv's
// cornerScore + 1, independent of the threshold).  Corner at threshold t <=> score > t; response = score - 1.
// Per thread: 4 horizontally adjacent pixels x 2 rows.  The 16 ring windows are 4-byte slices of 3 staged words
// per row; odd pixels use the slices as two u16 lanes (pixel in the high byte, neighbour as harmless junk in the
// low byte), even pixels use the slices shifted by one byte.  Arc extrema use VIMNMX / VIMNMX3 .U16x2:
//   min over arcs of (max over arc) = min_i max3(Q2[i], Q2[i+2], min(r[2i], r[2i+9])),  Q2 = maxima of 4 ring pixels.
// ---------------------------------------------------------------------------------------------------------------
constexpr int FT_PITCH = 84;           // staged ROI row pitch in bytes (ROI <= 76 wide, +4 guard, multiple of 4)
constexpr int FT_ROWS = 80;
constexpr int FS_PITCH = 80;           // score tile pitch (interior <= 70, +1 border each side, padded)

__device__ __forceinline__ uint32_t umax2(uint32_t a, uint32_t b) { return __vmaxu2(a, b); }
__device__ __forceinline__ uint32_t umin2(uint32_t a, uint32_t b) { return __vminu2(a, b); }
__device__ __forceinline__ uint32_t umax3(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_u16x2(a, b, c); }
__device__ __forceinline__ uint32_t umin3(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_u16x2(a, b, c); }

// a + b - min(a, b) on the FMA pipe: two IMADs with multiplier 1 (an asm block keeps ptxas from folding them back into one
// ALU-pipe IADD3)
__constant__ uint32_t c_one = 1u;   // a multiplier ptxas cannot fold: keeps the two adds of pair_max as IMADs (FMA pipe)
template <int PM>
__device__ __forceinline__ uint32_t pair_max(uint32_t a, uint32_t b, uint32_t mn) {
    if (PM == 1 || PM == 4 || PM == 6) return umax2(a, b);   // one VIMNMX (ALU pipe): fewer instructions, more ALU-pipe work
    const uint32_t one = c_one;
    return (a + b * one) - mn * one;           // two IMAD (FMA pipe)
}

// r[0..15] ring lanes, v centre lanes (pixel value in the high byte of each u16 lane).  Returns the two scores as
// clean u16 lanes.
template <int PM>
__device__ __forceinline__ uint32_t fast_score_lanes(const uint32_t (&r)[16], uint32_t v) {
    uint32_t qx[8], qn[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        // max = a + b - min holds for the whole word (exact integer identity per lane, carries cancel): the additions
        // can issue on the FMA pipe (IMAD.IADD) while the ALU pipe, which bounds this kernel, does the min/max
        qn[j] = umin2(r[2 * j + 1], r[(2 * j + 2) & 15]);
        qx[j] = pair_max<PM>(r[2 * j + 1], r[(2 * j + 2) & 15], qn[j]);
    }
    uint32_t q2x[8], q2n[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        q2x[i] = umax2(qx[i], qx[(i + 1) & 7]);
        q2n[i] = umin2(qn[i], qn[(i + 1) & 7]);
    }
    uint32_t fx[8], fn[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t a = r[2 * i], b = r[(2 * i + 9) & 15];
        const uint32_t mn = umin2(a, b);
        fx[i] = umax3(q2x[i], q2x[(i + 2) & 7], mn);   // max over arc, smaller of the two arcs sharing 8 pixels
        fn[i] = umin3(q2n[i], q2n[(i + 2) & 7], pair_max<(PM == 2 ? 1 : PM)>(a, b, mn));
    }
    uint32_t min_arc_max = umin3(umin3(fx[0], fx[1], fx[2]), umin3(fx[3], fx[4], fx[5]), umin2(fx[6], fx[7]));
    uint32_t max_arc_min = umax3(umax3(fn[0], fn[1], fn[2]), umax3(fn[3], fn[4], fn[5]), umax2(fn[6], fn[7]));
    // high bytes -> clean lanes; dark = v - min_arc_max, bright = max_arc_min - v, score = max(dark, bright, 0)
    const uint32_t hv = __byte_perm(v, 0, 0x4341), hx = __byte_perm(min_arc_max, 0, 0x4341), hn = __byte_perm(max_arc_min, 0, 0x4341);
    const uint32_t dark = hv + 0x01000100u - hx;       // 256 + (v - M) per lane, never borrows
    const uint32_t bright = hn + 0x01000100u - hv;
    return umax3(dark, bright, 0x01000100u) & 0x00FF00FFu;
}

// w[7][4]: staged words of rows y-3..y+3, 16 bytes each; the four pixels of interest sit at bytes M+3 .. M+6 (M = 0..3 is
// the misalignment of the ROI inside its 16-byte aligned TMA box; the fallback kernel stages aligned ROIs, M = 0).
// funnel shift right by 8 k bits as IMAD.HI + IMAD on the FMA pipe: (lo * 2^(32-8k)) >> 32 + hi * 2^(32-8k); the multipliers come from
// the constant bank so that ptxas cannot turn the pair back into one ALU-pipe SHF
__constant__ uint32_t c_fsh[4] = {0u, 1u << 24, 1u << 16, 1u << 8};
template <int PM>
__device__ __forceinline__ uint32_t slice_r(uint32_t lo, uint32_t hi, int k) {
    if (PM >= 5) { const uint32_t c = c_fsh[k]; return __umulhi(lo, c) + hi * c; }
    return __funnelshift_r(lo, hi, 8 * k);
}

template <int M, int PM>
__device__ __forceinline__ uint32_t fast_score4(const uint32_t (&w)[7][4]) {
    // 4-byte slices at byte offset o of a row
#define SL(row, o) ((((o) + M) & 3) == 0 ? w[row][((o) + M) >> 2] : slice_r<PM>(w[row][((o) + M) >> 2], w[row][(((o) + M) >> 2) + 1], ((o) + M) & 3))
    uint32_t r[16];
    r[0] = SL(6, 3);  r[1] = SL(6, 4);  r[2] = SL(5, 5);  r[3] = SL(4, 6);
    r[4] = SL(3, 6);  r[5] = SL(2, 6);  r[6] = SL(1, 5);  r[7] = SL(0, 4);
    r[8] = SL(0, 3);  r[9] = SL(0, 2);  r[10] = SL(1, 1); r[11] = SL(2, 0);
    r[12] = SL(3, 0); r[13] = SL(4, 0); r[14] = SL(5, 1); r[15] = SL(6, 2);
    const uint32_t c = SL(3, 3);
#undef SL
    const uint32_t odd = fast_score_lanes<PM>(r, c);      // pixels 1, 3
    uint32_t re[16];
#pragma unroll
    for (int k = 0; k < 16; k++) re[k] = r[k] << 8;
    const uint32_t even = fast_score_lanes<PM>(re, c << 8);   // pixels 0, 2
    return even | (odd << 8);
}

// scores of a 4-pixel x 2-row item: p = first staged word of row 2s of the item, pitch in bytes
template <int M, int PM>
__device__ __forceinline__ void fast_item(const uint8_t *p, int pitch, uint32_t &sa, uint32_t &sb) {
    uint32_t w[8][4];
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const uint32_t *q = reinterpret_cast<const uint32_t *>(p + r * pitch);
        w[r][0] = q[0]; w[r][1] = q[1]; w[r][2] = q[2];
        w[r][3] = M >= 2 ? q[3] : 0u;   // a window of 12 bytes is enough for M < 2
    }
    uint32_t wa[7][4], wb[7][4];
#pragma unroll
    for (int r = 0; r < 7; r++)
#pragma unroll
        for (int k = 0; k < 4; k++) { wa[r][k] = w[r][k]; wb[r][k] = w[r + 1][k]; }
    sa = fast_score4<M, PM>(wa); sb = fast_score4<M, PM>(wb);
}

__global__ void __launch_bounds__(192) k_fast_cells(const __grid_constant__ LevelTable T, const CellRect *__restrict__ cells,
                                                    int ini_th, int min_th, int f0, int *__restrict__ overflow) {
    const LevelDev *lv = T.lv;   // level table in the kernel parameter (constant) bank: no dependent global loads
    __shared__ __align__(16) uint8_t s_roi[FT_ROWS * FT_PITCH];
    __shared__ __align__(16) uint8_t s_sc[(FT_ROWS - 4) * FS_PITCH];   // interior scores with a 1-px zero ring
    __shared__ uint32_t s_list[36 * 36 + 8];
    __shared__ int s_n, s_base;
    const CellRect cell = cells[blockIdx.x];
    const LevelDev &L = lv[cell.level];
    const int f = f0 + blockIdx.y;
    const int rw = cell.x1 - cell.x0, rh = cell.y1 - cell.y0;   // ROI
    const int iw = rw - 6, ih = rh - 6;                            // tested pixels
    const uint8_t *__restrict__ src = L.img + (size_t)f * L.img_fstride + (size_t)cell.y0 * L.pitch + cell.x0;
    // stage ROI rows 0..rh (one extra row for the row-pair overhang) and rw+8 columns (4-pixel group overhang): one
    // warp per row, one aligned 32-bit word pair per lane realigned with a funnel shift.  All of it lies inside the
    // level plane because a ROI ends at least 16 px before the right / bottom edge; the overhang is masked later.
    {
        const int nwords = min((rw + 8) >> 2, FT_PITCH / 4);
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
        if (lane < nwords) {
            const uint8_t *p0 = src + 4 * lane;
            constexpr int RB = 7;   // rows in flight per lane (6 warps x 7 rows covers a 42-row ROI in one round)
            if ((L.pitch & 3) == 0) {
                // every row has the same misalignment: one aligned base pointer, rows are pitch/4 words apart
                const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(p0) & 3), sh = 8 * mis;
                const uint32_t *q0 = reinterpret_cast<const uint32_t *>(p0 - mis);
                const int pw = L.pitch >> 2;
                for (int r0 = wid; r0 <= rh; r0 += nwarps * RB) {
                    uint32_t lo[RB], hi[RB];
#pragma unroll
                    for (int k = 0; k < RB; k++) {
                        const int r = r0 + k * nwarps;
                        if (r <= rh) { const uint32_t *q = q0 + r * pw; lo[k] = __ldg(q); hi[k] = __ldg(q + 1); }
                    }
#pragma unroll
                    for (int k = 0; k < RB; k++) {
                        const int r = r0 + k * nwarps;
                        if (r <= rh) *reinterpret_cast<uint32_t *>(s_roi + r * FT_PITCH + 4 * lane) = __funnelshift_r(lo[k], hi[k], sh);
                    }
                }
            } else {
                for (int r = wid; r <= rh; r += nwarps) {
                    const uint8_t *p = p0 + (size_t)r * L.pitch;
                    *reinterpret_cast<uint32_t *>(s_roi + r * FT_PITCH + 4 * lane) =
                        (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
                }
            }
        }
    }
    for (int i = threadIdx.x; i < (ih + 2) * FS_PITCH / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(s_sc)[i] = 0;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    // scores: item = (4-pixel group g, row pair s)
    const int ng = (iw + 3) >> 2, ns = (ih + 1) >> 1;
    for (int it = threadIdx.x; it < ng * ns; it += blockDim.x) {
        const int s = (int)(((uint32_t)it * (65536u / (uint32_t)ng + 1u)) >> 16), g = it - s * ng;
        uint32_t sa, sb;
        fast_item<0, 0>(s_roi + (2 * s) * FT_PITCH + 4 * g, FT_PITCH, sa, sb);
        // mask pixels beyond the interior (group / row-pair overhang)
        const int valid = iw - 4 * g;
        if (valid < 4) { const uint32_t m = 0xFFFFFFFFu >> (8 * (4 - valid)); sa &= m; sb &= m; }
        // score tile: interior pixel (x, y) at [(y+1)*FS_PITCH + 4 + x]  (column 3 and row 0 are the zero ring)
        *reinterpret_cast<uint32_t *>(s_sc + (2 * s + 1) * FS_PITCH + 4 + 4 * g) = sa;
        if (2 * s + 1 < ih) *reinterpret_cast<uint32_t *>(s_sc + (2 * s + 2) * FS_PITCH + 4 + 4 * g) = sb;
    }
    __syncthreads();
    // NMS (strict 8-neighbour maximum inside the cell, neighbours outside count 0) + count at the initial threshold.
    // Same 4-pixel items as the scores; u16 lanes again: max over the 3x3 ring = max3(T(x-1), T(x+1), V(x)) with
    // V = max(up, down), T = max(V, mid); strict compare via (s & 0xFF00) > (m | 0x00FF) per lane.
    int n_ini_local = 0;
    {
        const uint32_t thr = ((uint32_t)min_th << 8) | ((uint32_t)min_th << 24);
        const uint32_t inv_ng = 65536u / (uint32_t)ng + 1u;   // exact floor(it / ng) for it < 65536 / ... (it < 18 * 74)
        for (int it = threadIdx.x; it < ng * ih; it += blockDim.x) {
            const int y = (int)(((uint32_t)it * inv_ng) >> 16), g = it - y * ng;
            const uint32_t *ru = reinterpret_cast<const uint32_t *>(s_sc + y * FS_PITCH) + g;   // words g, g+1, g+2: x-4.., x.., x+4..
            const uint32_t *rm = ru + FS_PITCH / 4, *rd = rm + FS_PITCH / 4;
            const uint32_t c_m = rm[1];
            if (c_m == 0) continue;
            const uint32_t u0 = ru[0], u1 = ru[1], u2 = ru[2], m0 = rm[0], m2 = rm[2], d0 = rd[0], d1 = rd[1], d2 = rd[2];
            // columns x-1 / x+1 of the four pixels
            const uint32_t lu = __funnelshift_r(u0, u1, 24), lm = __funnelshift_r(m0, c_m, 24), ld = __funnelshift_r(d0, d1, 24);
            const uint32_t ruu = __funnelshift_r(u1, u2, 8), rmm = __funnelshift_r(c_m, m2, 8), rdd = __funnelshift_r(d1, d2, 8);
            // odd pixels (1, 3): high bytes of the lanes as they are
            uint32_t mo = umax3(umax3(lu, ld, lm), umax3(ruu, rdd, rmm), umax3(u1, d1, thr));
            // even pixels (0, 2): shift everything by one byte
            uint32_t me = umax3(umax3(lu << 8, ld << 8, lm << 8), umax3(ruu << 8, rdd << 8, rmm << 8), umax3(u1 << 8, d1 << 8, thr));
            const uint32_t so = c_m & 0xFF00FF00u, se = (c_m << 8) & 0xFF00FF00u;
            mo |= 0x00FF00FFu; me |= 0x00FF00FFu;
            const uint32_t fo = umax2(so, mo) ^ mo, fe = umax2(se, me) ^ me;   // non-zero lane <=> strict maximum above minTh
            // flagged pixels of this item (at most 2: neighbours cannot both be strict maxima)
            uint32_t mask = ((fe & 0xFFFFu) ? 1u : 0u) | ((fo & 0xFFFFu) ? 2u : 0u) | ((fe >> 16) ? 4u : 0u) | ((fo >> 16) ? 8u : 0u);
            if (mask == 0) continue;
            int slot = atomicAdd(&s_n, __popc(mask));
            while (mask) {
                const int k = __ffs(mask) - 1;
                mask &= mask - 1;
                const int m = (c_m >> (8 * k)) & 0xFF, x = 4 * g + k;
                // relative coordinates (x - 16, y - 16) of the reference's vToDistributeKeys entries
                const uint32_t xr = (uint32_t)(cell.x0 + 3 + x - kMinBorder), yr = (uint32_t)(cell.y0 + 3 + y - kMinBorder);
                s_list[slot++] = (yr << 20) | (xr << 8) | (uint32_t)(m - 1);
                if (m > ini_th) n_ini_local++;
            }
        }
    }
    const int n_ini_total = __syncthreads_count(n_ini_local > 0);
    // per-cell fallback: if any corner passes iniThFAST keep only those, else keep everything above minThFAST
    const int keep_th = n_ini_total > 0 ? ini_th : min_th;
    const int n_all = s_n;
    __syncthreads();
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    uint32_t mine[8];
    int nmine = 0;
    for (int i = threadIdx.x; i < n_all; i += blockDim.x) {
        const uint32_t c = s_list[i];
        if ((int)(c & 0xFF) + 1 > keep_th && nmine < 8) mine[nmine++] = c;
    }
    const int my_off = nmine ? atomicAdd(&s_n, nmine) : 0;
    __syncthreads();
    if (threadIdx.x == 0) s_base = s_n ? atomicAdd(&L.cand_count[f], s_n) : 0;
    __syncthreads();
    uint32_t *__restrict__ out = L.cand + (size_t)f * L.cand_cap;
    for (int k = 0; k < nmine; k++) {
        const int pos = s_base + my_off + k;
        if (pos < L.cand_cap) out[pos] = mine[k]; else *overflow = 1;
    }
}

// TMA-staged, persistent, warp-tiled variant (the one the pipeline runs; k_fast_cells above stays as the fallback for
// level-0 planes that do not meet the TMA alignment rules).  One WARP owns one cell at a time: it walks (cell, frame)
// items with a grid-wide stride, keeps the ROI of its next item in flight (cp.async.bulk.tensor into its second
// shared-memory stage, completion on its own mbarrier) and needs nothing but __syncwarp between scoring, NMS and
// the hand-over of the candidates - no CTA barrier anywhere, which is what bounded the CTA-per-cell kernels
// (ncu: 27 % of the samples sat on the two barriers per cell).
//   * A TMA box must start on a 16-byte boundary of the innermost dimension (measured: an unaligned start coordinate
//     raises "illegal instruction"), so the box starts at x0 & ~15 and the 4-pixel groups are laid out on absolute
//     multiples of 4 (window = pixel - 3): every shared-memory read stays an aligned word, at the price of up to 3
//     masked pixels in the first group of a row.
//   * Candidates above iniThFAST fill the warp's list from the front, the rest from the back: the per-cell threshold
//     fallback is then a choice of sub-array, appended to the level's global list with one atomic per cell.
// Scores, NMS rule and emitted candidate sets are identical to k_fast_cells (candidate order is free by design).
constexpr int FW_WARPS = 4;             // warps per CTA (all independent); 6 CTAs/SM at 85 registers

struct FastTmaParams {
    CUtensorMap map[kMaxLevels];        // level plane [frames][h][w], box = box_w x box_h x 1
    int box_w[kMaxLevels], box_h[kMaxLevels];
    uint32_t *cand[kMaxLevels];         // [frames][cand_cap]
    int *cand_count[kMaxLevels];        // [frames]
    int cand_cap[kMaxLevels];
    int stage_bytes;                    // per-warp shared memory: ROI stage (multiple of 128) ...
    int sc_pitch, sc_bytes;             // ... score tile (row pitch, size) ...
    int list_cap;                       // ... candidate list (u16 entries) ...
    int warp_bytes;                     // ... total per warp (multiple of 128)
};

// PM: how the 16 pair maxima per pass are formed -- 0: a + b - min as two IMAD (FMA pipe), 1: VIMNMX (ALU pipe), 2: the Q pairs by IMAD,
// the (r[2i], r[2i+9]) pairs by VIMNMX; 3 / 4: as 0 / 1 with the NMS flags taken from VIMNMX predicate outputs; 5 / 6: as 3 / 4 with the ring slices cut by IMAD.HI + IMAD
template <int PM>
__global__ void __launch_bounds__(FW_WARPS * 32, 6) k_fast_tma(const __grid_constant__ FastTmaParams P, const CellRect *__restrict__ cells,
                                                               int ncells, int total, int ini_th, int min_th, int f0,
                                                               int *__restrict__ overflow) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *wsm = smem + (size_t)warp * P.warp_bytes;
    uint8_t *s_roi = wsm;                                                         // stage_bytes
    uint8_t *s_sc = wsm + P.stage_bytes;                                          // sc_bytes
    uint16_t *s_list = reinterpret_cast<uint16_t *>(s_sc + P.sc_bytes);           // list_cap entries: tile row << 8 | tile byte column
    uint64_t *s_full = reinterpret_cast<uint64_t *>(s_list + P.list_cap);         // 1 mbarrier
    int *s_cnt = reinterpret_cast<int *>(s_full + 1);                             // [above iniTh, rest]
    const int scp = P.sc_pitch, lcap = P.list_cap;

    if (lane == 0) {
        tma_mbar_init(s_full, 1);
        tma_mbar_fence_init();
        s_cnt[0] = 0; s_cnt[1] = 0;
    }
    __syncwarp();
    const int stride = gridDim.x * FW_WARPS;
    int item = blockIdx.x * FW_WARPS + warp;
    if (item >= total) return;
    auto load_cell = [&](int it, int &fi) -> CellRect {
        CellRect c; c.level = -1; c.x0 = c.y0 = c.x1 = c.y1 = 0; fi = 0;
        if (it < total) { fi = it / ncells; c = cells[it - fi * ncells]; }
        return c;
    };
    auto issue = [&](const CellRect &c, int fi) {      // lane 0 only
        tma_mbar_expect_tx(s_full, (uint32_t)(P.box_w[c.level] * P.box_h[c.level]));
        tma_load_3d(s_roi, &P.map[c.level], c.x0 & ~15, c.y0, f0 + fi, s_full);
    };
    int fi_cur, fi_nxt;
    CellRect cur = load_cell(item, fi_cur), nxt = load_cell(item + stride, fi_nxt);
    if (lane == 0) issue(cur, fi_cur);
    const uint32_t thr = ((uint32_t)min_th << 8) | ((uint32_t)min_th << 24);
    for (int q = 0; item < total; item += stride, q++) {
        int fi_nn;
        const CellRect nn = load_cell(item + 2 * stride, fi_nn);            // descriptor prefetch, consumed next iteration
        const int f = f0 + fi_cur, level = cur.level;
        const int rp = P.box_w[level];                                      // staged row pitch in bytes
        const int iw = cur.x1 - cur.x0 - 6, ih = cur.y1 - cur.y0 - 6;       // tested pixels
        // groups of 4 pixels on absolute multiples of 4: window = bytes X-3 .. X+8 of the first pixel X of a group
        const int lead_px = cur.x0 & 3;                                     // first tested column sits at lane lead_px of group 0
        const int wbase = (cur.x0 & ~3) - (cur.x0 & ~15);                   // byte offset of group 0's window in the box
        const int ng = (lead_px + iw + 3) >> 2, ns = (ih + 1) >> 1;
        const uint32_t inv_ng = 65536u / (uint32_t)ng + 1u;
        // zero ring of the score tile for this geometry: rows 0, ih+1 (and ih+2 for the row-pair overhang), words 0 and
        // ng+1 of the rows between
        for (int i = lane; i < ng + 2; i += 32) {
            reinterpret_cast<uint32_t *>(s_sc)[i] = 0;
            reinterpret_cast<uint32_t *>(s_sc + (ih + 1) * scp)[i] = 0;
            reinterpret_cast<uint32_t *>(s_sc + (ih + 2) * scp)[i] = 0;
        }
        for (int y = 1 + lane; y <= ih; y += 32) {
            reinterpret_cast<uint32_t *>(s_sc + y * scp)[0] = 0;
            reinterpret_cast<uint32_t *>(s_sc + y * scp)[ng + 1] = 0;
        }
        tma_mbar_wait(s_full, (uint32_t)q & 1u);
        // scores: work item = (group g, row pair s)
        for (int wi = lane; wi < ng * ns; wi += 32) {
            const int s2 = (int)(((uint32_t)wi * inv_ng) >> 16), g = wi - s2 * ng;
            uint32_t sa, sb;
            fast_item<0, PM>(s_roi + (2 * s2) * rp + wbase + 4 * g, rp, sa, sb);
            // mask pixels outside the tested columns (leading lanes of group 0, trailing lanes of the last group)
            const int lo = lead_px - 4 * g, hi = lead_px + iw - 4 * g;       // valid lanes: lo <= k < hi
            uint32_t m = 0xFFFFFFFFu;
            if (lo > 0) m &= 0xFFFFFFFFu << (8 * lo);
            if (hi < 4) m &= 0xFFFFFFFFu >> (8 * (4 - hi));
            sa &= m; sb &= m;
            if (2 * s2 + 1 >= ih) sb = 0;
            // score tile: lane k of group g on tested row y at [(y+1)*scp + 4 + 4g + k]  (word 0 and row 0 are the zero ring)
            *reinterpret_cast<uint32_t *>(s_sc + (2 * s2 + 1) * scp + 4 + 4 * g) = sa;
            *reinterpret_cast<uint32_t *>(s_sc + (2 * s2 + 2) * scp + 4 + 4 * g) = sb;
        }
        __syncwarp();
        // the ROI stage is free again: start the next item's TMA now; it lands while this item is suppressed and emitted
        if (lane == 0 && nxt.level >= 0) issue(nxt, fi_nxt);
        // NMS (strict 8-neighbour maximum inside the cell, neighbours outside count 0), same (group, row pair) items:
        // per tile row the 3-wide horizontal maximum H and the left/right maximum LR (with minTh folded in) are formed
        // once in u16 lanes (odd pixels as they are, even pixels shifted up a byte); ring(y) = max3(H(y-1), H(y+1), LR(y)).
        for (int wi = lane; wi < ng * ns; wi += 32) {
            const int s2 = (int)(((uint32_t)wi * inv_ng) >> 16), g = wi - s2 * ng;
            const uint8_t *t0 = s_sc + (2 * s2) * scp + 4 * g;               // tile rows 2s .. 2s+3 = tested rows 2s-1 .. 2s+2
            uint32_t mid[4], Ho[4], He[4], LRo[4], LRe[4];
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const uint32_t *p = reinterpret_cast<const uint32_t *>(t0 + r * scp);
                const uint32_t w0 = p[0], w1 = p[1], w2 = p[2];
                const uint32_t l = __funnelshift_r(w0, w1, 24), rr = __funnelshift_r(w1, w2, 8);
                mid[r] = w1;
                Ho[r] = umax3(l, w1, rr); LRo[r] = umax3(l, rr, thr);
                const uint32_t le = l << 8, me = w1 << 8, re = rr << 8;
                He[r] = umax3(le, me, re); LRe[r] = umax3(le, re, thr);
            }
#pragma unroll
            for (int c = 1; c <= 2; c++) {
                const uint32_t c_m = mid[c];
                if (c_m == 0) continue;
                const uint32_t mo = umax3(Ho[c - 1], Ho[c + 1], LRo[c]) | 0x00FF00FFu;
                const uint32_t me = umax3(He[c - 1], He[c + 1], LRe[c]) | 0x00FF00FFu;
                const uint32_t so = c_m & 0xFF00FF00u, se = (c_m << 8) & 0xFF00FF00u;
                uint32_t mask;
                if (PM >= 3) {       // strict maximum above minTh <=> NOT (ring maximum >= score): VIMNMX with predicate outputs
                    bool oh, ol, eh, el;
                    __vibmax_u16x2(mo, so, &oh, &ol);
                    __vibmax_u16x2(me, se, &eh, &el);
                    mask = (el ? 0u : 1u) | (ol ? 0u : 2u) | (eh ? 0u : 4u) | (oh ? 0u : 8u);
                } else {
                    const uint32_t fo = umax2(so, mo) ^ mo, fe = umax2(se, me) ^ me;   // non-zero lane <=> strict maximum above minTh
                    mask = ((fe & 0xFFFFu) ? 1u : 0u) | ((fo & 0xFFFFu) ? 2u : 0u) | ((fe >> 16) ? 4u : 0u) | ((fo >> 16) ? 8u : 0u);
                }
                while (mask) {
                    const int k = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const int m = (c_m >> (8 * k)) & 0xFF;
                    const uint16_t e = (uint16_t)(((2 * s2 + c) << 8) | (4 + 4 * g + k));   // tile row, tile byte column
                    if (m > ini_th) s_list[atomicAdd(&s_cnt[0], 1)] = e;
                    else s_list[lcap - 1 - atomicAdd(&s_cnt[1], 1)] = e;
                }
            }
        }
        __syncwarp();
        // per-cell fallback: if any corner passes iniThFAST keep only those, else keep everything above minThFAST
        {
            const int n_ini = s_cnt[0], n_low = s_cnt[1];
            const int n = n_ini > 0 ? n_ini : n_low;
            const uint16_t *src = s_list + (n_ini > 0 ? 0 : lcap - n_low);
            if (n > 0) {
                int base = 0;
                if (lane == 0) base = atomicAdd(P.cand_count[level] + f, n);
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                const int cap = P.cand_cap[level];
                uint32_t *__restrict__ out = P.cand[level] + (size_t)f * cap;
                // relative coordinates (x - 16, y - 16) of the reference's vToDistributeKeys entries
                const int xrel = cur.x0 + 3 - lead_px - 4 - kMinBorder, yrel = cur.y0 + 3 - 1 - kMinBorder;
                for (int i = lane; i < n; i += 32) {
                    const uint32_t e = src[i], ty = e >> 8, tx = e & 0xFF;
                    const uint32_t score = s_sc[ty * scp + tx];
                    const uint32_t v = ((uint32_t)(yrel + (int)ty) << 20) | ((uint32_t)(xrel + (int)tx) << 8) | (score - 1u);
                    if (base + i < cap) out[base + i] = v; else *overflow = 1;
                }
            }
            __syncwarp();
            if (lane == 0) { s_cnt[0] = 0; s_cnt[1] = 0; }
            __syncwarp();
        }
        cur = nxt; fi_cur = fi_nxt; nxt = nn; fi_nxt = fi_nn;
    }
}

int launch_fast(const LevelDev *h_levels, const CellRect *d_cells, int ncells, int f0, int batch,
                int ini_th, int min_th, int *d_overflow, cudaStream_t stream, const FastTma *tma, int sm_count) {
    if (ncells <= 0) return 0;
    if (tma && tma->ok) {
        static_assert(sizeof(FastTma::map) == sizeof(FastTmaParams::map), "tensor map storage mismatch");
        FastTmaParams P;
        memcpy(P.map, tma->map, sizeof(P.map));
        int stage = 128;
        for (int l = 0; l < kMaxLevels; l++) {
            P.box_w[l] = tma->box_w[l]; P.box_h[l] = tma->box_h[l];
            P.cand[l] = h_levels[l].cand; P.cand_count[l] = h_levels[l].cand_count; P.cand_cap[l] = h_levels[l].cand_cap;
            stage = max(stage, (tma->box_w[l] * tma->box_h[l] + 8 + 127) / 128 * 128);   // + 8: the last window may read past its row
        }
        P.stage_bytes = stage;
        P.sc_pitch = 4 * ((tma->max_iw + 3 + 3) / 4 + 2);
        P.sc_bytes = (P.sc_pitch * (tma->max_ih + 4) + 8 + 15) / 16 * 16;
        P.list_cap = (((tma->max_iw + 1) / 2) * ((tma->max_ih + 1) / 2) + 8 + 3) / 4 * 4;
        P.warp_bytes = (stage + P.sc_bytes + P.list_cap * 2 + 8 + 2 * 4 + 127) / 128 * 128;
        const size_t smem = (size_t)FW_WARPS * P.warp_bytes;
        // function attributes are per device: one slot per device ordinal (a process may hold handles on several GPUs)
        static size_t configured_[kMaxDevices], last_[kMaxDevices];
        static int per_sm_[kMaxDevices];   // resident CTAs per SM: the persistent grid is exactly one wave
        const int dv = current_device_slot();
        std::lock_guard<std::mutex> lock(g_attr_mutex);
        size_t &configured = configured_[dv], &last = last_[dv];
        int &per_sm = per_sm_[dv];
        static const int pm_env = [] { const char *e = getenv("ORBX_FAST_PM"); return e ? atoi(e) : 3; }();   // measured: 0 225-227, 1 238, 2 232, 3 223-225, 4 236, 5 230, 6 240 us
        typedef void (*FastKernel)(const FastTmaParams, const CellRect *, int, int, int, int, int, int *);
        const FastKernel fn = pm_env == 1 ? k_fast_tma<1> : pm_env == 2 ? k_fast_tma<2> : pm_env == 3 ? k_fast_tma<3> : pm_env == 4 ? k_fast_tma<4> : pm_env == 5 ? k_fast_tma<5> : pm_env == 6 ? k_fast_tma<6> : k_fast_tma<0>;
        if (smem > configured) {
            cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            configured = smem;
        }
        if (smem != last) {
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, FW_WARPS * 32, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
            last = smem;
        }
        const int total = ncells * batch;
        const int want = (total + FW_WARPS - 1) / FW_WARPS;
        static const int cap_env = [] { const char *e = getenv("ORBX_FAST_CTAS"); return e ? atoi(e) : 0; }();
        const int resident = cap_env > 0 && cap_env < per_sm ? cap_env : per_sm;
        const int grid = want < sm_count * resident ? want : sm_count * resident;
        fn<<<grid, FW_WARPS * 32, smem, stream>>>(P, d_cells, ncells, total, ini_th, min_th, f0, d_overflow);
        return 1;
    }
    dim3 grid(ncells, batch);
    k_fast_cells<<<grid, 192, 0, stream>>>(make_table(h_levels), d_cells, ini_th, min_th, f0, d_overflow);
    return 1;
}

// ---------------------------------------------------------------------------------------------------------------
// K3  DistributeOctTree.  One CTA per (frame, level).
//
// The reference's list algorithm is replayed breadth-first with block-wide scans.  Its quadrants are fixed by the
// root box (ceil-halving), so a key's path is a function of its integer coordinates: the host tabulates it
// (LevelDev::xbin/ybin, depth0 levels, Morton order).  One histogram pass over the candidates then gives the key
// count of EVERY node down to depth0 as a difference of two prefix sums; deeper nodes (only reached in sparse,
// clustered frames) are split by partitioning their slice of a bin-sorted copy of the keys.  The winner of a leaf
// (max response, first in the reference's emission order on ties) comes from a per-bin atomicMax of
// (score << 24 | ~order).  List order, the (size, creation order) sort of the final phase and its early break are
// reproduced with prefix sums over list positions.
// ---------------------------------------------------------------------------------------------------------------
struct __align__(8) QNode {
    int16_t ulx, uly, brx, bry;
    int lo, hi;            // key range (positions in bin-sorted order)
    uint32_t prefix;       // root << 2*depth | Morton path
    int depth;
};

constexpr int OT_THREADS = 256;       // threads per (frame, level) problem when many problems run side by side
constexpr int OT_THREADS_WIDE = 512;   // ... for single-frame calls and for large problems (1080p, the 6250-feature initialisation extractor): the
                                       // breadth-first replay is a chain of block-wide passes over up to 4096 nodes.  Measured (quadtree stage,
                                       // us, 256 / 512 / 1024 threads): 640x480 x1 27 / 25 / 27, 1280x800 x1 46 / 39 / 41, 6250 features 86 / 60 / 62,
                                       // 1080p x16 112 / 67 / 68 -- but 64 x 640x480 39 / 52 and 7 k frames/s off the four-lane step

// exclusive scan of a[0..n) in place; returns the total to every thread.  s_warp: NT / 32 ints of scratch.
template <int NT>
__device__ int block_excl_scan(int *a, int n, int *s_warp) {
    const int tid = threadIdx.x, per = (n + NT - 1) / NT;
    const int b = tid * per, e = min(b + per, n);
    int sum = 0;
    for (int i = b; i < e; i++) sum += a[i];
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if ((tid & 31) >= o) inc += t; }
    __syncthreads();           // s_warp may still be read from a previous call
    if ((tid & 31) == 31) s_warp[tid >> 5] = inc;
    __syncthreads();
    // offsets of the warps: one scan of the (at most 32) warp totals, done redundantly by every warp
    int wv = (tid & 31) < NT / 32 ? s_warp[tid & 31] : 0, winc = wv;
#pragma unroll
    for (int o = 1; o < NT / 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, winc, o); if ((tid & 31) >= o) winc += t; }
    const int woff = __shfl_sync(0xFFFFFFFFu, winc - wv, tid >> 5), total = __shfl_sync(0xFFFFFFFFu, winc, NT / 32 - 1);
    int run = woff + inc - sum;
    for (int i = b; i < e; i++) { const int v = a[i]; a[i] = run; run += v; }
    __syncthreads();
    return total;
}

struct OctShared {
    int len, phase, finish, need_sorted, sorted_done, cand_n, rstar, ctot;
};

__device__ __forceinline__ void child_box(const QNode &p, int q, QNode &c) {
    const int mx = p.ulx + ((p.brx - p.ulx + 1) >> 1), my = p.uly + ((p.bry - p.uly + 1) >> 1);   // ceil(half)
    c.ulx = (q & 1) ? mx : p.ulx; c.brx = (q & 1) ? p.brx : mx;
    c.uly = (q & 2) ? my : p.uly; c.bry = (q & 2) ? p.bry : my;
    c.prefix = p.prefix * 4u + (uint32_t)q;
    c.depth = p.depth + 1;
}

// split points s[0..4] of node p (s[0]=lo, s[4]=hi): children q own [s[q], s[q+1]).
__device__ void split_node(const LevelDev &L, const QNode &p, const int *bin_start, uint32_t *sorted, int *s, bool sorted_ok,
                           int *need_sorted) {
    s[0] = p.lo; s[4] = p.hi;
    if (p.depth < L.depth0) {
        const int sh = 2 * (L.depth0 - p.depth - 1);
        const uint32_t b = p.prefix * 4u;
        s[1] = bin_start[(b + 1) << sh]; s[2] = bin_start[(b + 2) << sh]; s[3] = bin_start[(b + 3) << sh];
        return;
    }
    if (!sorted_ok) { *need_sorted = 1; s[1] = s[2] = s[3] = p.lo; return; }
    // below the tabulated depth: 4-way in-place partition of this node's slice (small by construction)
    const int mx = p.ulx + ((p.brx - p.ulx + 1) >> 1), my = p.uly + ((p.bry - p.uly + 1) >> 1);
    int c0 = 0, c1 = 0, c2 = 0;
    for (int k = p.lo; k < p.hi; k++) {
        const uint32_t key = sorted[k];
        const int x = (key >> 8) & 0xFFF, y = key >> 20;
        const int q = (x < mx ? 0 : 1) | (y < my ? 0 : 2);
        c0 += q == 0; c1 += q == 1; c2 += q == 2;
    }
    s[1] = p.lo + c0; s[2] = s[1] + c1; s[3] = s[2] + c2;
    int cur0 = s[0], cur1 = s[1], cur2 = s[2], cur3 = s[3];
    for (int b = 0; b < 3; b++) {
        const int end = s[b + 1];
        int &cur = b == 0 ? cur0 : b == 1 ? cur1 : cur2;
        while (cur < end) {
            const uint32_t key = sorted[cur];
            const int x = (key >> 8) & 0xFFF, y = key >> 20;
            const int q = (x < mx ? 0 : 1) | (y < my ? 0 : 2);
            if (q == b) { cur++; continue; }
            int &dst = q == 1 ? cur1 : q == 2 ? cur2 : cur3;   // q > b always here
            const uint32_t other = sorted[dst];
            sorted[dst] = key; sorted[cur] = other; dst++;
        }
    }
}

template <int NT>
__global__ void __launch_bounds__(NT) k_octree(const __grid_constant__ LevelTable T, int nlevels, int cap_nodes,
                                                       int f0, int *__restrict__ overflow) {
    const LevelDev *lv = T.lv;   // level table in the kernel parameter (constant) bank: no dependent global loads
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int level = blockIdx.x, f = f0 + blockIdx.y;
    const LevelDev &L = lv[level];
    const int tid = threadIdx.x;
    __shared__ OctShared S;
    __shared__ int s_warp[NT / 32];
    if (L.n_ini <= 0 || L.nbins <= 0) { if (tid == 0) L.sel_count[f] = 0; return; }
    int n = L.cand_count[f];
    if (n > L.cand_cap) n = L.cand_cap;
    const int N = L.quota;
    // shared carve-up
    QNode *cur = reinterpret_cast<QNode *>(smem_raw);
    QNode *nxt = cur + cap_nodes;
    int *bin_start = reinterpret_cast<int *>(nxt + cap_nodes);             // [nbins + 1]
    uint32_t *best = reinterpret_cast<uint32_t *>(bin_start + L.nbins + 1);   // [nbins]
    int *a_cc = reinterpret_cast<int *>(best + L.nbins);      // [cap] children count -> offsets
    int *a_kp = a_cc + cap_nodes;                              // [cap] kept / processed flags -> offsets
    int *a_s1 = a_kp + cap_nodes, *a_s2 = a_s1 + cap_nodes, *a_s3 = a_s2 + cap_nodes;   // split points
    int *a_rank = a_s3 + cap_nodes;                            // [cap] processing rank of a candidate / order
    int *a_ord = a_rank + cap_nodes;                           // [cap] list position by rank
    const uint32_t *__restrict__ cand = L.cand + (size_t)f * L.cand_cap;
    uint32_t *sorted = L.sorted + (size_t)f * L.cand_cap;

    for (int i = tid; i < L.nbins; i += NT) { bin_start[i] = 0; best[i] = 0; }
    if (tid == 0) { bin_start[L.nbins] = 0; S.len = 0; S.phase = 0; S.finish = 0; S.need_sorted = 0; S.sorted_done = 0; }
    __syncthreads();
    // pass A: histogram + per-bin winner.  Four candidates per thread are in flight (candidate word, then its four LUT
    // entries) before the shared-memory atomics of any of them: the loop is bound by the dependent global loads.
    {
        const uint32_t *__restrict__ xbin = L.xbin, *__restrict__ ybin = L.ybin, *__restrict__ xord = L.xord, *__restrict__ yord = L.yord;
        for (int k0 = tid; k0 < n; k0 += 4 * NT) {
            uint32_t c[4], bb[4], oo[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { const int k = k0 + u * NT; c[u] = k < n ? __ldg(cand + k) : 0u; }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t xr = (c[u] >> 8) & 0xFFF, yr = c[u] >> 20;
                bb[u] = __ldg(xbin + xr) | __ldg(ybin + yr);
                oo[u] = __ldg(xord + xr) + __ldg(yord + yr);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (k0 + u * NT < n) {
                    atomicAdd(&bin_start[bb[u]], 1);
                    atomicMax(&best[bb[u]], ((c[u] & 0xFFu) << 24) | (0xFFFFFFu - oo[u]));
                }
            }
        }
    }
    __syncthreads();
    block_excl_scan<NT>(bin_start, L.nbins + 1, s_warp);   // bin_start[b] = first key of bin b, bin_start[nbins] = n
    // roots (push_back order), empty ones erased
    if (tid == 0) {
        int len = 0;
        for (int r = 0; r < L.n_ini; r++) {
            QNode q;
            q.ulx = (int16_t)L.root_ulx[r]; q.brx = (int16_t)L.root_brx[r]; q.uly = 0; q.bry = (int16_t)L.reg_h;
            q.lo = bin_start[r << (2 * L.depth0)]; q.hi = bin_start[(r + 1) << (2 * L.depth0)];
            q.prefix = (uint32_t)r; q.depth = 0;
            if (q.hi > q.lo) cur[len++] = q;
        }
        S.len = len; S.cand_n = 0;
    }
    __syncthreads();

    while (true) {
        const int len = S.len;
        const bool final_phase = S.phase != 0;
        // region of nodes to divide: every non-single node (full round) or the candidates created by the last step
        const int region = final_phase ? S.cand_n : len;
        if (tid == 0) S.need_sorted = 0;
        __syncthreads();
        const bool sorted_ok = S.sorted_done != 0;
        for (int i = tid; i < len; i += NT) {
            int cc = 0;
            a_kp[i] = 0;
            if (i < region) {
                const QNode p = cur[i];
                if (p.hi - p.lo > 1) {
                    int s[5];
                    split_node(L, p, bin_start, sorted, s, sorted_ok, &S.need_sorted);
                    a_s1[i] = s[1]; a_s2[i] = s[2]; a_s3[i] = s[3];
                    cc = (s[1] > s[0]) + (s[2] > s[1]) + (s[3] > s[2]) + (s[4] > s[3]);
                }
            }
            a_cc[i] = cc;       // 0 => not divided in this step
        }
        __syncthreads();
        if (S.need_sorted) {
            // the tree wants to go below depth0: build the bin-sorted key copy once, then redo this step
            uint32_t *cursor = L.bin_cursor + (size_t)f * L.nbins;
            for (int i = tid; i < L.nbins; i += NT) cursor[i] = 0;
            __syncthreads();
            for (int k = tid; k < n; k += NT) {
                const uint32_t c = cand[k];
                const uint32_t b = L.xbin[(c >> 8) & 0xFFF] | L.ybin[c >> 20];
                sorted[bin_start[b] + atomicAdd(&cursor[b], 1u)] = c;
            }
            __threadfence_block();
            __syncthreads();
            if (tid == 0) S.sorted_done = 1;
            __syncthreads();
            continue;
        }

        int new_len, ctot, expandable = 0;
        if (!final_phase) {
            // ---- full round: every divided node is replaced by its non-empty children (pushed to the front in
            //      creation order => reversed), single-key nodes keep their relative order behind them
            for (int i = tid; i < len; i += NT) a_kp[i] = (a_cc[i] == 0) ? 1 : 0;
            __syncthreads();
            ctot = block_excl_scan<NT>(a_cc, len, s_warp);
            const int kept = block_excl_scan<NT>(a_kp, len, s_warp);
            new_len = ctot + kept;
            if (new_len > cap_nodes) { if (tid == 0) { *overflow = 2; L.sel_count[f] = 0; } return; }
            for (int i = tid; i < len; i += NT) {
                const QNode p = cur[i];
                if (p.hi - p.lo <= 1) { nxt[ctot + a_kp[i]] = p; continue; }
                const int s[5] = {p.lo, a_s1[i], a_s2[i], a_s3[i], p.hi};
                int k = a_cc[i];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    if (s[q + 1] > s[q]) {
                        QNode c; child_box(p, q, c); c.lo = s[q]; c.hi = s[q + 1];
                        nxt[ctot - 1 - k] = c; k++;
                        if (c.hi - c.lo > 1) expandable++;
                    }
                }
            }
        } else {
            // ---- final phase: candidates (count > 1, created by the previous step = list positions [0, region))
            //      are divided largest first, ties by creation order descending = list position ascending, until
            //      the list holds N nodes
            for (int i = tid; i < region; i += NT) {
                int rank = -1;
                if (a_cc[i] > 0) {
                    const int sz = cur[i].hi - cur[i].lo;
                    rank = 0;
                    for (int j = 0; j < region; j++) {
                        if (a_cc[j] > 0) {
                            const int sj = cur[j].hi - cur[j].lo;
                            rank += (sj > sz) || (sj == sz && j < i);
                        }
                    }
                }
                a_rank[i] = rank;
            }
            __syncthreads();
            // number of candidates
            for (int i = tid; i < region; i += NT) if (a_rank[i] >= 0) a_ord[a_rank[i]] = i;
            if (tid == 0) { S.rstar = 0x7FFFFFFF; S.ctot = 0; }
            __syncthreads();
            // total candidates via scan of flags
            for (int i = tid; i < region; i += NT) a_kp[i] = a_rank[i] >= 0 ? 1 : 0;
            __syncthreads();
            const int ncand = block_excl_scan<NT>(a_kp, region, s_warp);
            // growth in processing order: a_s? reused as scratch is not possible (split points live there) -> use a_kp
            for (int r = tid; r < ncand; r += NT) a_kp[r] = a_cc[a_ord[r]] - 1;
            __syncthreads();
            block_excl_scan<NT>(a_kp, ncand, s_warp);          // a_kp[r] = growth before rank r
            for (int r = tid; r < ncand; r += NT) {
                const int after = len + a_kp[r] + a_cc[a_ord[r]] - 1;
                if (after >= N) atomicMin(&S.rstar, r);
            }
            __syncthreads();
            const int rstar = min(S.rstar, ncand - 1);     // last processed rank (-1 if no candidates)
            const int nproc = rstar + 1;
            // children offsets in processing order
            for (int r = tid; r < ncand; r += NT) a_kp[r] = (r < nproc) ? a_cc[a_ord[r]] : 0;
            __syncthreads();
            ctot = block_excl_scan<NT>(a_kp, ncand, s_warp);   // a_kp[r] = children created before rank r
            // a_rank[i] (list position) -> child offset, or -1 if the node is not processed
            for (int i = tid; i < len; i += NT) {
                int v = -1;
                if (i < region && a_rank[i] >= 0 && a_rank[i] < nproc) v = a_kp[a_rank[i]];
                a_ord[i] = v;      // a_ord no longer needed as rank->position map
            }
            __syncthreads();
            for (int i = tid; i < len; i += NT) a_kp[i] = a_ord[i] >= 0 ? 1 : 0;
            __syncthreads();
            block_excl_scan<NT>(a_kp, len, s_warp);            // a_kp[i] = processed nodes before list position i
            new_len = len - nproc + ctot;
            if (new_len > cap_nodes) { if (tid == 0) { *overflow = 2; L.sel_count[f] = 0; } return; }
            for (int i = tid; i < len; i += NT) {
                const QNode p = cur[i];
                if (a_ord[i] < 0) { nxt[ctot + i - a_kp[i]] = p; continue; }
                const int s[5] = {p.lo, a_s1[i], a_s2[i], a_s3[i], p.hi};
                int k = a_ord[i];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    if (s[q + 1] > s[q]) {
                        QNode c; child_box(p, q, c); c.lo = s[q]; c.hi = s[q + 1];
                        nxt[ctot - 1 - k] = c; k++;
                    }
                }
            }
        }
        __syncthreads();
        // reduce `expandable` (full round only)
        if (!final_phase) {
            if (tid == 0) S.ctot = 0;
            __syncthreads();
            if (expandable) atomicAdd(&S.ctot, expandable);
            __syncthreads();
            expandable = S.ctot;
        }
        __syncthreads();
        if (tid == 0) {
            const int prev = S.len;
            S.len = new_len;
            S.cand_n = ctot;                 // nodes created by this step sit at list positions [0, ctot)
            if (new_len >= N || new_len == prev) S.finish = 1;
            else if (!final_phase && new_len + 3 * expandable > N) S.phase = 1;
        }
        { QNode *t = cur; cur = nxt; nxt = t; }
        __syncthreads();
        if (S.finish) break;
    }

    // ---- leaves -> keypoints: best response per node, reference emission order on ties
    const int len = S.len;
    uint32_t *__restrict__ sel = L.sel + (size_t)f * L.out_cap;
    if (len > L.out_cap) { if (tid == 0) { *overflow = 3; L.sel_count[f] = 0; } return; }
    for (int i = tid; i < len; i += NT) {
        const QNode p = cur[i];
        uint32_t bv = 0;
        if (p.depth <= L.depth0) {
            const int sh = 2 * (L.depth0 - p.depth);
            const uint32_t b0 = p.prefix << sh, b1 = (p.prefix + 1u) << sh;
            for (uint32_t b = b0; b < b1; b++) bv = max(bv, best[b]);
        } else {
            for (int k = p.lo; k < p.hi; k++) {
                const uint32_t c = sorted[k];
                const uint32_t xr = (c >> 8) & 0xFFF, yr = c >> 20;
                bv = max(bv, ((c & 0xFFu) << 24) | (0xFFFFFFu - (L.xord[xr] + L.yord[yr])));
            }
        }
        const uint32_t ord = 0xFFFFFFu - (bv & 0xFFFFFFu), score = bv >> 24;
        const uint32_t cellid = ord / L.ord_cell_area, rem = ord - cellid * L.ord_cell_area;
        const uint32_t yin = rem / (uint32_t)L.wcell, xin = rem - yin * (uint32_t)L.wcell;
        const uint32_t ci = cellid / L.ord_ncols, cj = cellid - ci * L.ord_ncols;
        const uint32_t x = cj * (uint32_t)L.wcell + xin + 3u + kMinBorder, y = ci * (uint32_t)L.hcell + yin + 3u + kMinBorder;
        sel[i] = (y << 20) | (x << 8) | score;
    }
    if (tid == 0) L.sel_count[f] = len;
}

static size_t octree_smem_bytes(int nbins, int cap) {
    return (size_t)2 * cap * sizeof(QNode) + (size_t)(2 * nbins + 1) * 4 + (size_t)7 * cap * 4 + 16;
}

int launch_octree(const LevelDev *h_levels, int nlevels, int f0, int batch, int *d_overflow,
                  cudaStream_t stream) {
    // one launch for all levels: size the node arrays for the largest quota and the bins for the finest table
    int cap = 0, nbins = 0;
    for (int l = 0; l < nlevels; l++) {
        if (h_levels[l].out_cap + 8 > cap) cap = h_levels[l].out_cap + 8;
        if (h_levels[l].nbins > nbins) nbins = h_levels[l].nbins;
    }
    cap = (cap + 1) & ~1;
    const size_t smem = octree_smem_bytes(nbins, cap);
    static size_t configured_[kMaxDevices];
    size_t &configured = configured_[current_device_slot()];
    {
        std::lock_guard<std::mutex> lock(g_attr_mutex);
        if (smem > configured) {
            cudaFuncSetAttribute(k_octree<OT_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(k_octree<OT_THREADS_WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            configured = smem;
        }
    }
    dim3 grid(nlevels, batch);
    // ORBX_OCTREE_WIDE=0/1 overrides the choice
    static const int wide_env = getenv("ORBX_OCTREE_WIDE") ? atoi(getenv("ORBX_OCTREE_WIDE")) : -1;
    const bool big = (long long)h_levels[0].w * h_levels[0].h >= 1500000 || h_levels[0].quota >= 1000;
    const bool wide = wide_env >= 0 ? wide_env != 0 : (batch <= 4 || big);
    if (wide) k_octree<OT_THREADS_WIDE><<<grid, OT_THREADS_WIDE, smem, stream>>>(make_table(h_levels), nlevels, cap, f0, d_overflow);
    else k_octree<OT_THREADS><<<grid, OT_THREADS, smem, stream>>>(make_table(h_levels), nlevels, cap, f0, d_overflow);
    return 1;
}

// ---------------------------------------------------------------------------------------------------------------
// K7  output slots.  One CTA per frame: walks levels 0..L-1 in list order, scales to image coordinates and assigns
//     the reference's slot (lapping-area keypoints fill from the back, the rest from the front).
// ---------------------------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NT) k_finalize(const __grid_constant__ LevelTable T, int nlevels, int total_out_cap, int lap0,
                                                  int lap1, KeypointRec *__restrict__ kp, int cap, int *__restrict__ slot,
                                                  uint2 *__restrict__ items, int *__restrict__ n_out, int *__restrict__ mono_out, int f0,
                                                  int *__restrict__ overflow) {
    const LevelDev *lv = T.lv;   // level table in the kernel parameter (constant) bank: no dependent global loads
    const int f = f0 + blockIdx.x, tid = threadIdx.x;
    __shared__ int s_cnt[kMaxLevels + 1];
    __shared__ int s_warp[NT / 32];
    __shared__ int s_run;
    if (tid < 32) {   // level offsets: the counts of all levels are fetched at once (lane = level), then one warp scan
        const int v = tid < nlevels ? lv[tid].sel_count[f] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (tid >= o) inc += t; }
        if (tid < nlevels) s_cnt[tid] = inc - v;
        if (tid == nlevels - 1) s_cnt[nlevels] = inc;
        if (tid == 0) s_run = 0;
    }
    __syncthreads();
    const int ntot = s_cnt[nlevels];
    if (ntot > cap) {
        if (tid == 0) { *overflow = 4; n_out[f] = 0; mono_out[f] = 0; }
        for (int i = tid; i < total_out_cap; i += NT) items[(size_t)f * total_out_cap + i] = make_uint2(0xFFFFFFFFu, 0u);
        return;
    }
    for (int i = ntot + tid; i < total_out_cap; i += NT) items[(size_t)f * total_out_cap + i] = make_uint2(0xFFFFFFFFu, 0u);
    const float flap0 = (float)lap0, flap1 = (float)lap1;
    // chunks of NT keypoints in sequence order; running count of lapping-area keypoints carried across chunks
    for (int base = 0; base < ntot; base += NT) {
        const int i = base + tid;
        int l = 0, inlap = 0;
        float x = 0.f, y = 0.f, resp = 0.f;
        uint32_t c = 0;
        if (i < ntot) {
            while (i >= s_cnt[l + 1]) l++;
            c = lv[l].sel[(size_t)f * lv[l].out_cap + (i - s_cnt[l])];
            x = (float)((c >> 8) & 0xFFF); y = (float)(c >> 20); resp = (float)(c & 0xFF);
            if (l != 0) { x = __fmul_rn(x, lv[l].scale); y = __fmul_rn(y, lv[l].scale); }
            inlap = (x >= flap0 && x <= flap1) ? 1 : 0;
        }
        // block exclusive scan of inlap
        int inc = inlap;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if ((tid & 31) >= o) inc += t; }
        if ((tid & 31) == 31) s_warp[tid >> 5] = inc;
        __syncthreads();
        int wv = (tid & 31) < NT / 32 ? s_warp[tid & 31] : 0, winc = wv;
#pragma unroll
        for (int o = 1; o < NT / 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, winc, o); if ((tid & 31) >= o) winc += t; }
        const int woff = __shfl_sync(0xFFFFFFFFu, winc - wv, tid >> 5), tot = __shfl_sync(0xFFFFFFFFu, winc, NT / 32 - 1);
        const int before = s_run + woff + inc - inlap;     // lapping keypoints before i
        if (i < ntot) {
            const int sl = inlap ? (ntot - 1 - before) : (i - before);
            KeypointRec r;
            r.x = x; r.y = y; r.size = lv[l].kp_size; r.angle = -1.f; r.response = resp; r.octave = l; r.class_id = -1;
            kp[(size_t)f * cap + sl] = r;
            slot[(size_t)f * total_out_cap + lv[l].out_base + (i - s_cnt[l])] = sl;
            items[(size_t)f * total_out_cap + i] = make_uint2(((uint32_t)l << 24) | ((c >> 20) << 12) | ((c >> 8) & 0xFFFu), (uint32_t)sl);
        }
        __syncthreads();
        if (tid == 0) s_run += tot;
        __syncthreads();
    }
    if (tid == 0) { n_out[f] = ntot; mono_out[f] = ntot - s_run; }
}

int launch_finalize(const LevelDev *h_levels, int nlevels, int f0, int batch, int total_out_cap, int lap0, int lap1,
                    KeypointRec *d_kp, int cap, int *d_slot, uint2 *d_items, int *d_n, int *d_mono, int *d_overflow, cudaStream_t stream) {
    // a handful of frames: 1024 threads each (the chunks of a frame are a serial chain); a full batch: 256, so that the CTAs fit beside
    // the machine-filling kernels of the other frame ranges
    if (batch <= 8) k_finalize<1024><<<batch, 1024, 0, stream>>>(make_table(h_levels), nlevels, total_out_cap, lap0, lap1, d_kp, cap, d_slot, d_items, d_n, d_mono, f0, d_overflow);
    else k_finalize<256><<<batch, 256, 0, stream>>>(make_table(h_levels), nlevels, total_out_cap, lap0, lap1, d_kp, cap, d_slot, d_items, d_n, d_mono, f0, d_overflow);
    return 1;
}

// ---------------------------------------------------------------------------------------------------------------
// K4 + K6  one warp per keypoint: IC_Angle moments by warp reduction, cv::fastAtan2 in individually rounded fp32,
//          then 256 steered BRIEF tests (8 per lane) on the blurred plane.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float s = 57.29577951308232f;   // (float)(180 / CV_PI)
    const float p1 = __fmul_rn(0.9997878412794807f, s), p3 = __fmul_rn(-0.3258083974640975f, s);
    const float p5 = __fmul_rn(0.1555786518463281f, s), p7 = __fmul_rn(-0.04432655554792128f, s);
    const float eps = 2.220446049250313e-16f;   // (float)DBL_EPSILON
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0.f) a = __fsub_rn(180.f, a);
    if (y < 0.f) a = __fsub_rn(360.f, a);
    return a;
}

__device__ __forceinline__ void warp_ic_moments(const uint8_t *__restrict__ center, int pitch, int lane, int &m01_out, int &m10_out) {
    // lane <-> column u = lane - 15; all 31 row loads of a lane are independent (fully unrolled, predicated by the
    // circular patch mask umax[|v|] = 15,15,15,15,14,14,14,13,13,12,11,10,9,8,6,3)
    constexpr int UM[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
    const int u = lane - kHalfPatch, au = u < 0 ? -u : u;
    int sum = 0, m01 = 0;
    const uint8_t *p = center + u;
#pragma unroll
    for (int v = -kHalfPatch; v <= kHalfPatch; v++) {
        const int d = UM[v < 0 ? -v : v];
        const int val = (au <= d) ? (int)__ldg(p + v * pitch) : 0;
        sum += val; m01 += v * val;
    }
    int m10 = u * sum;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { m10 += __shfl_xor_sync(0xFFFFFFFFu, m10, o); m01 += __shfl_xor_sync(0xFFFFFFFFu, m01, o); }
    m01_out = m01; m10_out = m10;
}

__device__ __forceinline__ float warp_ic_angle(const uint8_t *__restrict__ center, int pitch, int lane) {
    int m01, m10;
    warp_ic_moments(center, pitch, lane, m01, m10);
    return fast_atan2_deg((float)m01, (float)m10);
}

// cosf / sinf of the reference are modelled as the correctly rounded fp32 value of the fp64 result (DESIGN.md "trig rule")
__device__ __forceinline__ void steer_trig(float angle_deg, float &a, float &b) {
    const float factor = 0.017453292519943295f;   // (float)(CV_PI / 180.f)
    const float ang = __fmul_rn(angle_deg, factor);
    double sd, cd;
    sincos((double)ang, &sd, &cd);
    a = __double2float_rn(cd); b = __double2float_rn(sd);
}

__device__ __forceinline__ uint8_t warp_brief_byte(const uint8_t *__restrict__ center, int pitch, float a, float b, int lane) {
    // (measured: a float4 pattern table and a running row pointer for the moments both made this kernel slower, 81 -> 119 us)
    const char4 *pat = reinterpret_cast<const char4 *>(g_pattern) + lane * 8;
    uint32_t val = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const char4 p = pat[j];
        const float x0 = (float)p.x, y0 = (float)p.y, x1 = (float)p.z, y1 = (float)p.w;
        const int ix0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
        const int iy0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
        const int ix1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
        const int iy1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
        const int t0 = center[iy0 * pitch + ix0], t1 = center[iy1 * pitch + ix1];
        val |= (uint32_t)(t0 < t1) << j;
    }
    return (uint8_t)val;
}

// One warp handles DG consecutive keypoint slots: the moments of keypoint j end up in lane j, so that fastAtan2 and
// the fp64 sincos of the steering angle are evaluated once per keypoint in separate lanes (the fp64 pipe is narrow;
// evaluating them warp-wide per keypoint made this kernel DP-bound) and then broadcast for the 256 tests.
constexpr int DG = 1;   // measured: 8 slots per warp is slower (87 -> 107 us): the kernel is bound by the scattered BRIEF gathers, not by fp64
__global__ void __launch_bounds__(256) k_describe(const __grid_constant__ LevelTable T, int nlevels, int total_out_cap,
                                                  const int *__restrict__ slot, KeypointRec *__restrict__ kp,
                                                  uint8_t *__restrict__ desc, int cap, int f0) {
    const LevelDev *lv = T.lv;   // level table in the kernel parameter (constant) bank: no dependent global loads
    const int f = f0 + blockIdx.y, lane = threadIdx.x & 31;
    const int item0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * DG;
    if (item0 >= total_out_cap) return;
    int lvl[DG], px[DG], py[DG];
    int my01 = 0, my10 = 0;
    int l = 0;
#pragma unroll
    for (int j = 0; j < DG; j++) {
        const int item = item0 + j;
        lvl[j] = -1; px[j] = py[j] = 0;
        if (item >= total_out_cap) continue;
        while (l + 1 < nlevels && item >= lv[l + 1].out_base) l++;
        const LevelDev &L = lv[l];
        const int idx = item - L.out_base;
        if (idx >= L.sel_count[f]) continue;
        const uint32_t c = L.sel[(size_t)f * L.out_cap + idx];
        lvl[j] = l; px[j] = (c >> 8) & 0xFFF; py[j] = c >> 20;
        int m01, m10;
        warp_ic_moments(L.img + (size_t)f * L.img_fstride + (size_t)py[j] * L.pitch + px[j], L.pitch, lane, m01, m10);
        if (lane == j) { my01 = m01; my10 = m10; }
    }
    float angle = 0.f, a = 1.f, b = 0.f;
    if (lane < DG) {
        angle = fast_atan2_deg((float)my01, (float)my10);
        steer_trig(angle, a, b);
    }
#pragma unroll
    for (int j = 0; j < DG; j++) {
        const float aj = __shfl_sync(0xFFFFFFFFu, a, j), bj = __shfl_sync(0xFFFFFFFFu, b, j), angj = __shfl_sync(0xFFFFFFFFu, angle, j);
        if (lvl[j] < 0) continue;
        const LevelDev &L = lv[lvl[j]];
        const uint8_t *bl = L.blur + (size_t)f * L.blur_fstride + (size_t)py[j] * L.blur_pitch + px[j];
        const uint8_t byte = warp_brief_byte(bl, L.blur_pitch, aj, bj, lane);
        const int sl = slot[(size_t)f * total_out_cap + item0 + j];
        desc[((size_t)f * cap + sl) * 32 + lane] = byte;
        if (lane == 0) kp[(size_t)f * cap + sl].angle = angj;
    }
}

// explicit shared-space byte load: generic pointers into shared memory cost an address-space conversion per access
__device__ __forceinline__ uint32_t lds_u8(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}

// warp_brief_byte on a patch in shared memory (center = shared-space address of the keypoint, rows 64 bytes apart)
__device__ __forceinline__ uint8_t warp_brief_byte_smem(uint32_t center, float a, float b, int lane) {
    const char4 *pat = reinterpret_cast<const char4 *>(g_pattern) + lane * 8;
    uint32_t val = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const char4 p = pat[j];
        const float x0 = (float)p.x, y0 = (float)p.y, x1 = (float)p.z, y1 = (float)p.w;
        const int ix0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
        const int iy0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
        const int ix1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
        const int iy1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
        const uint32_t t0 = lds_u8(center + (uint32_t)(iy0 * 64 + ix0)), t1 = lds_u8(center + (uint32_t)(iy1 * 64 + ix1));
        val |= (uint32_t)(t0 < t1) << j;
    }
    return (uint8_t)val;
}

// warp_brief_byte on a patch in shared memory with the lane's 8 pattern tests already in fp32 registers (pq[j] = x0, y0, x1, y1) and
// cvRound as the fp32 magic-number addition: for |x| < 2^22, fadd.rn(x, 1.5 * 2^23) has rint(x) (ties to even, as cvRound) in its low
// mantissa bits, so bits(x + M) = bits(M) + rint(x).  center65 = shared-space address of the keypoint minus 65 * bits(M): the byte
// offset iy * 64 + ix then comes out of one IMAD on the raw bit patterns (all arithmetic modulo 2^32).  No I2F / F2I in the loop:
// conversions issue at a quarter of the FMA rate and were ~28 of the kernel's 71 us (profiles/k_describe_tma_r01_summary.txt).
__device__ __forceinline__ uint8_t warp_brief_byte_regs(uint32_t center65, float a, float b, const float4 (&pq)[8]) {
    const float M = 12582912.f;
    uint32_t val = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const float4 p = pq[j];
        const uint32_t ix0 = __float_as_uint(__fadd_rn(__fsub_rn(__fmul_rn(p.x, a), __fmul_rn(p.y, b)), M));
        const uint32_t iy0 = __float_as_uint(__fadd_rn(__fadd_rn(__fmul_rn(p.x, b), __fmul_rn(p.y, a)), M));
        const uint32_t ix1 = __float_as_uint(__fadd_rn(__fsub_rn(__fmul_rn(p.z, a), __fmul_rn(p.w, b)), M));
        const uint32_t iy1 = __float_as_uint(__fadd_rn(__fadd_rn(__fmul_rn(p.z, b), __fmul_rn(p.w, a)), M));
        const uint32_t t0 = lds_u8(center65 + iy0 * 64u + ix0), t1 = lds_u8(center65 + iy1 * 64u + ix1);
        val |= (uint32_t)(t0 < t1) << j;
    }
    return (uint8_t)val;
}

// TMA-staged, persistent variant (the one the pipeline runs when every plane meets the TMA alignment rules).  One warp owns one
// keypoint at a time and walks the (frame, slot) items with a grid stride.  The 31-row patch of the un-blurred level (IC_Angle)
// and the 39-row patch of the blurred level (the rotated test points lie within radius 18.4 of the keypoint) are brought into
// the warp's shared memory by cp.async.bulk.tensor, two keypoints deep, so the 31 moment loads and the 16 scattered BRIEF
// gathers per lane become LDS with immediate / 32-bit offsets: k_describe spends half of its instructions on 64-bit address
// arithmetic and is bound by the scattered gathers in L1.  Arithmetic and results are identical to k_describe.
constexpr int DT_WARPS = 4;
constexpr int DT_BOXW = 64, DT_IMG_ROWS = 31, DT_BLUR_ROWS = 39, DT_R = 19;
constexpr int DT_IMG_BYTES = DT_BOXW * DT_IMG_ROWS, DT_BLUR_BYTES = DT_BOXW * DT_BLUR_ROWS;      // 1984 + 2496 bytes moved by TMA
constexpr int DT_BLUR_OFS = 2048;                                                               // TMA destinations are 128-byte aligned
constexpr int DT_STAGE = 4608;                                                                  // 2048 + 2496, rounded to 128
constexpr int DT_WARP_BYTES = 2 * DT_STAGE + 128;                                               // + mbarriers

struct DescTmaParams {
    CUtensorMap img[kMaxLevels], blur[kMaxLevels];   // level planes [frames][h][w]; boxes 64 x 31 and 64 x 39
};

struct DescItem { int level, x, y, sl, f; };         // level < 0: nothing to do for this item

__global__ void __launch_bounds__(DT_WARPS * 32, 6) k_describe_tma(const __grid_constant__ DescTmaParams P, const __grid_constant__ LevelTable T,
                                                                   int nlevels, int total_out_cap, int nframes, const uint2 *__restrict__ items,
                                                                   KeypointRec *__restrict__ kp, uint8_t *__restrict__ desc, int cap, int f0) {
    extern __shared__ __align__(128) uint8_t smem[];
    const LevelDev *lv = T.lv;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *wsm = smem + (size_t)warp * DT_WARP_BYTES;
    uint64_t *s_full = reinterpret_cast<uint64_t *>(wsm + 2 * DT_STAGE);
    if (lane == 0) { tma_mbar_init(&s_full[0], 1); tma_mbar_init(&s_full[1], 1); tma_mbar_fence_init(); }
    __syncwarp();
    const int total = nframes * total_out_cap, stride = gridDim.x * DT_WARPS;
    // work list written by k_finalize: one 8-byte record per keypoint (level, position, output slot), dense per frame
    auto load_item = [&](int i) -> DescItem {
        DescItem d; d.level = -1; d.x = d.y = d.sl = d.f = 0;
        if (i >= total) return d;
        const int fi = i / total_out_cap, k = i - fi * total_out_cap, f = f0 + fi;
        const uint2 r = __ldg(items + (size_t)f * total_out_cap + k);
        if (r.x == 0xFFFFFFFFu) return d;
        d.level = (int)(r.x >> 24); d.x = (int)(r.x & 0xFFFu); d.y = (int)((r.x >> 12) & 0xFFFu); d.f = f; d.sl = (int)r.y;
        return d;
    };
    auto issue = [&](const DescItem &d, int st) {        // lane 0 only
        uint8_t *dst = wsm + st * DT_STAGE;
        const int xa = (d.x - DT_R) & ~15;               // a TMA box starts on a 16-byte boundary
        tma_mbar_expect_tx(&s_full[st], (uint32_t)(DT_IMG_BYTES + DT_BLUR_BYTES));
        tma_load_3d(dst, &P.img[d.level], xa, d.y - kHalfPatch, d.f, &s_full[st]);
        tma_load_3d(dst + DT_BLUR_OFS, &P.blur[d.level], xa, d.y - DT_R, d.f, &s_full[st]);
    };
    int i_cur = blockIdx.x * DT_WARPS + warp;
    DescItem A = load_item(i_cur), B = load_item(i_cur + stride);
    if (lane == 0) { if (A.level >= 0) issue(A, 0); if (B.level >= 0) issue(B, 1); }
    uint32_t phase = 0;                                   // bit st = parity to wait for on stage st
    constexpr int UM[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
    const int u = lane - kHalfPatch, au = u < 0 ? -u : u;
    // the lane's 8 BRIEF tests (bits 8 lane .. 8 lane + 7) as fp32, once per persistent warp
    float4 pq[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const char4 p = reinterpret_cast<const char4 *>(g_pattern)[lane * 8 + j];
        pq[j] = make_float4((float)p.x, (float)p.y, (float)p.z, (float)p.w);
    }
    for (int q = 0; i_cur < total; q++, i_cur += stride) {
        const int st = q & 1;
        const DescItem Cn = load_item(i_cur + 2 * stride);                 // descriptor prefetch for the item after next
        if (A.level >= 0) {
            tma_mbar_wait(&s_full[st], (phase >> st) & 1u);
            phase ^= 1u << st;
            const uint8_t *ip = wsm + st * DT_STAGE, *bp = ip + DT_BLUR_OFS;
            const int xo = A.x - ((A.x - DT_R) & ~15);                      // keypoint column inside the box: 19 .. 34
            // IC_Angle moments: lane <-> column u, 31 rows at immediate offsets
            int sum = 0, m01 = 0;
            const uint32_t pc = tma_smem_u32(ip) + (uint32_t)(xo + u);
#pragma unroll
            for (int v = -kHalfPatch; v <= kHalfPatch; v++) {
                const int d = UM[v < 0 ? -v : v];
                const int val = (au <= d) ? (int)lds_u8(pc + (uint32_t)((v + kHalfPatch) * DT_BOXW)) : 0;
                sum += val; m01 += v * val;
            }
            int m10 = u * sum;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { m10 += __shfl_xor_sync(0xFFFFFFFFu, m10, o); m01 += __shfl_xor_sync(0xFFFFFFFFu, m01, o); }
            float angle = 0.f, a = 1.f, b = 0.f;
            if (lane == 0) {
                angle = fast_atan2_deg((float)m01, (float)m10);
                steer_trig(angle, a, b);
            }
            angle = __shfl_sync(0xFFFFFFFFu, angle, 0); a = __shfl_sync(0xFFFFFFFFu, a, 0); b = __shfl_sync(0xFFFFFFFFu, b, 0);
            const uint8_t byte = warp_brief_byte_regs(tma_smem_u32(bp) + (uint32_t)(DT_R * DT_BOXW + xo) - 65u * 0x4B400000u, a, b, pq);
            desc[((size_t)A.f * cap + A.sl) * 32 + lane] = byte;
            if (lane == 0) kp[(size_t)A.f * cap + A.sl].angle = angle;
        }
        __syncwarp();
        if (lane == 0 && Cn.level >= 0) issue(Cn, st);                     // the stage has been consumed: refill it
        A = B; B = Cn;
    }
}

int launch_describe(const LevelDev *h_levels, int nlevels, int f0, int batch, int total_out_cap, const int *d_slot, const uint2 *d_items,
                    KeypointRec *d_kp, uint8_t *d_desc, int cap, cudaStream_t stream, const DescTma *tma, int sm_count) {
    if (tma && tma->ok) {
        static_assert(sizeof(DescTma::img) == sizeof(DescTmaParams::img), "tensor map storage mismatch");
        DescTmaParams P;
        memcpy(P.img, tma->img, sizeof(P.img));
        memcpy(P.blur, tma->blur, sizeof(P.blur));
        const size_t smem = (size_t)DT_WARPS * DT_WARP_BYTES;
        static int per_sm_[kMaxDevices];
        int &per_sm = per_sm_[current_device_slot()];
        std::unique_lock<std::mutex> lock(g_attr_mutex);
        if (!per_sm) {
            cudaFuncSetAttribute(k_describe_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_describe_tma, DT_WARPS * 32, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
        }
        const int total = batch * total_out_cap;
        const int want = (total + DT_WARPS - 1) / DT_WARPS;
        const int grid = want < sm_count * per_sm ? want : sm_count * per_sm;
        k_describe_tma<<<grid, DT_WARPS * 32, smem, stream>>>(P, make_table(h_levels), nlevels, total_out_cap, batch, d_items, d_kp, d_desc, cap, f0);
        return 1;
    }
    dim3 grid((total_out_cap + 8 * DG - 1) / (8 * DG), batch);
    k_describe<<<grid, 256, 0, stream>>>(make_table(h_levels), nlevels, total_out_cap, d_slot, d_kp, d_desc, cap, f0);
    return 1;
}

__global__ void __launch_bounds__(256) k_describe_points(const uint8_t *__restrict__ img, const uint8_t *__restrict__ blur, int pitch,
                                                         const float *__restrict__ xy, int n, const float *__restrict__ angle_in,
                                                         float *__restrict__ angle_out, uint8_t *__restrict__ desc) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    const int x = __float2int_rn(xy[2 * i]), y = __float2int_rn(xy[2 * i + 1]);
    float angle;
    if (angle_in) angle = angle_in[i];
    else angle = warp_ic_angle(img + (size_t)y * pitch + x, pitch, lane);
    if (angle_out && lane == 0) angle_out[i] = angle;
    if (desc && blur) {
        float a, b;
        steer_trig(angle, a, b);
        desc[(size_t)i * 32 + lane] = warp_brief_byte(blur + (size_t)y * pitch + x, pitch, a, b, lane);
    }
}

int launch_describe_points(const uint8_t *d_img, const uint8_t *d_blur, int pitch, const float *d_xy, int n,
                           const float *d_angle_in, float *d_angle_out, uint8_t *d_desc, cudaStream_t stream) {
    if (n <= 0) return 0;
    k_describe_points<<<(n + 7) / 8, 256, 0, stream>>>(d_img, d_blur, pitch, d_xy, n, d_angle_in, d_angle_out, d_desc);
    return 1;
}

}  // namespace orbx

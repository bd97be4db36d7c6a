// K5 (tensor-core variant)  GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) of every pyramid level with the VERTICAL pass on the tensor
// cores and the horizontal pass in registers.
//
// Replaces the per-level cv::GaussianBlur of UPSTREAM ORB-SLAM3 ORBextractor::operator() (kernel
// [18,34,48,56,48,34,18]/256 per axis, out = (sum sum + 2^15) >> 16, exact).  Same bytes as k_blur_tma / k_blur (orbx_kernels.cu); what
// changes is who does the arithmetic:
//   * a tile is 122 output rows x 96 output columns of one level of one frame; its input box (128 rows x 128 columns at
//     (x0 - 16, y0 - 3), zero fill outside the plane) arrives by ONE TMA load with SWIZZLE_128B -- which is exactly the canonical MN-major
//     operand layout of tcgen05.mma (rows of 128 bytes = 128 image columns, 8-row groups 1024 bytes apart): the image tile is the B operand
//     as it lies in shared memory;
//   * the A operand is the constant 128 x 128 band matrix A[m][k] = w[k - m] (K-major, SWIZZLE_128B), built once per CTA;
//   * D[m][n] = sum_k A[m][k] B[k][n] = the vertical 7-tap sums (<= 65280, int32 in TMEM) by four tcgen05.mma kind::i8 (u8 x u8, K = 32);
//   * an epilogue thread owns one output row (TMEM lane): it reads its column sums, packs neighbours into u16 pairs, forms each output
//     pixel with four IDP.2A (the horizontal taps), rounds and writes 16-byte pieces of its row;
//   * REFLECT_101 rows / columns are patched in the swizzled tile before the MMA reads it (fence.proxy.async in between).
// About 6.5 instructions per pixel against 19 for k_blur_tma; tools/blur_tc_probe.cu is the stand-alone version with its own CPU check.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "orbx_dev.h"
#include "orbx_tma.cuh"

namespace orbx {

extern std::mutex g_attr_mutex;

namespace {

constexpr int TR = kBlurTcTileH, TC = kBlurTcTileW;   // 122 x 96 output tile
constexpr int BOX = 128;                              // input box: 128 bytes x 128 rows
constexpr uint32_t A_BYTES = 16384, B_BYTES = 16384;
constexpr int NACC = 2;                               // accumulator stages (128 TMEM columns each)
constexpr int NBMAX = 4;                              // input tiles in flight (template parameter NB <= NBMAX)
constexpr int THREADS = 192;                          // warp 0: TMA + MMA issue, warp 1: edge patches, warps 2..5: epilogue (TMEM lane quarter = warp % 4).
                                                      // Twelve epilogue warps (one per lane quarter and 32-column chunk) were measured slower: 45 -> 52 us;
                                                      // requesting the next chunk's sums before working on the current one: 56 -> 55 us alone but 155
                                                      // registers and 189 -> 186 k frames/s with four batches in flight

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(bar)) : "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major SWIZZLE_128B: rows of 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t desc_k_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major SWIZZLE_128B: rows (one per k) of 128 B = 128 MN elements, 8-row groups 1024 B apart (SBO), 128-element MN blocks LBO apart
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
#define TMEM_LD8(taddr, v)                                                                                               \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"                          \
                 : "=r"((v)[0]), "=r"((v)[1]), "=r"((v)[2]), "=r"((v)[3]), "=r"((v)[4]), "=r"((v)[5]), "=r"((v)[6]), "=r"((v)[7]) \
                 : "r"(taddr) : "memory")
#define TMEM_LD32(taddr, v)                                                                                              \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                               \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                               \
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"               \
                 : "=r"((v)[0]), "=r"((v)[1]), "=r"((v)[2]), "=r"((v)[3]), "=r"((v)[4]), "=r"((v)[5]), "=r"((v)[6]), "=r"((v)[7]),       \
                   "=r"((v)[8]), "=r"((v)[9]), "=r"((v)[10]), "=r"((v)[11]), "=r"((v)[12]), "=r"((v)[13]), "=r"((v)[14]), "=r"((v)[15]), \
                   "=r"((v)[16]), "=r"((v)[17]), "=r"((v)[18]), "=r"((v)[19]), "=r"((v)[20]), "=r"((v)[21]), "=r"((v)[22]), "=r"((v)[23]), \
                   "=r"((v)[24]), "=r"((v)[25]), "=r"((v)[26]), "=r"((v)[27]), "=r"((v)[28]), "=r"((v)[29]), "=r"((v)[30]), "=r"((v)[31]) \
                 : "r"(taddr) : "memory")


// byte address of (row i, byte column b) of a SWIZZLE_128B tile with 128-byte rows
__device__ __forceinline__ uint32_t sw(int i, int b) { return (uint32_t)(i * 128 + ((((b >> 4) ^ (i & 7)) << 4) | (b & 15))); }

struct Ctl {
    uint64_t full[NBMAX], patched[NBMAX], empty[NBMAX], acc_full[NACC], acc_empty[NACC];
    uint32_t tmem_base, pad;
};

struct BlurTcParams {
    CUtensorMap map[kMaxLevels];       // un-blurred level plane [frames][h][w], box 128 x 128 x 1, SWIZZLE_128B
    CUtensorMap omap[kMaxLevels];      // blurred level plane, box 96 x 122 x 1: the finished tile leaves by one TMA store, clipped at the plane's edges
    int w[kMaxLevels], h[kMaxLevels];
};
constexpr uint32_t O_BYTES = (TR * TC + 127) / 128 * 128;   // one output tile, rows of 96 bytes

// Persistent CTAs; NB input boxes in flight (the box of tile i + NB - 1 is requested while tile i is multiplied: a box takes over a
// microsecond to arrive, a tile's MMAs a tenth of that), two accumulator stages between the MMA issuer and the epilogue warps.
// (One tile per CTA with 128 TMEM columns -- short-lived CTAs that slot in anywhere -- was measured slower: 57 vs 45 us.)
template <int NB>
__global__ void __launch_bounds__(THREADS, 1) k_blur_tc(const __grid_constant__ BlurTcParams P, const BlurTile *__restrict__ tiles, int ntiles_frame,
                                                        int total, int f0) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (s32(smem_raw) & 1023u)) & 1023u);
    uint8_t *sA = smem, *sB = smem + A_BYTES;
    uint8_t *sO = smem + A_BYTES + NB * B_BYTES;
    Ctl &S = *reinterpret_cast<Ctl *>(sO + 2 * O_BYTES);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

    // A[m][k] = w[k - m] for m < 122, 0 <= k - m <= 6 (K-major, 128-byte rows, 16-byte pieces swizzled by the row)
    for (int i = threadIdx.x; i < (int)(A_BYTES / 16); i += THREADS) reinterpret_cast<uint4 *>(sA)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    if (threadIdx.x < TR) {
        const int m = threadIdx.x;
        const uint8_t W7[7] = {18, 34, 48, 56, 48, 34, 18};
#pragma unroll
        for (int t = 0; t < 7; t++) sA[sw(m, m + t)] = W7[t];
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < NB; s++) { tma_mbar_init(&S.full[s], 1); tma_mbar_init(&S.patched[s], 1); tma_mbar_init(&S.empty[s], 1); }
        for (int s = 0; s < NACC; s++) { tma_mbar_init(&S.acc_full[s], 1); tma_mbar_init(&S.acc_empty[s], 4); }
        tma_mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&S.tmem_base)), "n"(128 * NACC) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the band matrix was written by ordinary stores
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = S.tmem_base;
    const int first = blockIdx.x, step = gridDim.x;
    // item -> (tile of the level list, frame): frames are the slow index, so that the CTAs of one wave work on neighbouring tiles
    auto tile_of = [&](int it, int &f) -> BlurTile { f = it / ntiles_frame; return tiles[it - f * ntiles_frame]; };

    if (warp == 0) {
        // ===== TMA + MMA issue (whole warp in the loop, one elected lane issues) =====
        const uint32_t idesc = (2u << 4) | (1u << 16) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // s32 accumulate, u8 x u8, B MN-major
        const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
        auto request = [&](int ti, int sb) {
            int f; const BlurTile t = tile_of(ti, f);
            if (elect_one()) {
                tma_mbar_expect_tx(&S.full[sb], BOX * BOX);
                tma_load_3d(sB + sb * B_BYTES, &P.map[t.level], t.tx * TC - 16, t.ty * TR - 3, f0 + f, &S.full[sb]);
            }
        };
        for (int p = 0; p < NB - 1; p++)
            if (first + p * step < total) request(first + p * step, p);
        uint32_t it = 0;
        for (int ti = first; ti < total; ti += step, it++) {
            const int sb = it % NB, pb = (it / NB) & 1, sa = it % NACC;
            if (ti + (NB - 1) * step < total) {  // box of tile it + NB - 1 into the stage tile it - 1 used, as soon as its MMAs have retired
                const int s2 = (it + NB - 1) % NB;
                if (it >= 1) tma_mbar_wait(&S.empty[s2], ((it - 1) / NB) & 1);
                request(ti + (NB - 1) * step, s2);
            }
            tma_mbar_wait(&S.patched[sb], pb);                                          // box landed and its edges are reflected
            if (it >= NACC) tma_mbar_wait(&S.acc_empty[sa], ((it - NACC) / NACC) & 1);  // accumulator stage read out
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const uint32_t d = tb + sa * 128;
                const uint64_t da = desc_k_sw128(s32(sA)), db = desc_mn_sw128(s32(sB + sb * B_BYTES));
#pragma unroll
                for (int ks = 0; ks < 4; ks++)       // K = 32 per step: 32 bytes along an A row, 32 rows (4 KB) of the box
                    umma_i8(d, da + 2 * ks, db + 256 * ks, idesc, ks ? 1u : 0u);
                umma_commit(&S.empty[sb]);
                umma_commit(&S.acc_full[sa]);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===== edge patches: REFLECT_101 rows, then columns, in the swizzled tile =====
        uint32_t it = 0;
        for (int ti = first; ti < total; ti += step, it++) {
            const int s = it % NB, ph = (it / NB) & 1;
            int f; const BlurTile t = tile_of(ti, f);
            const int w = P.w[t.level], h = P.h[t.level], x0 = t.tx * TC, y0 = t.ty * TR;
            uint8_t *B = sB + s * B_BYTES;
            tma_mbar_wait(&S.full[s], ph);
            const bool top = y0 == 0, bottom = y0 - 3 + BOX > h, left = x0 == 0, right = x0 + 112 > w;
            if (top || bottom) {
                // box row i holds image row y0 - 3 + i; 16-byte pieces keep their logical column, the swizzle depends on the row
                for (int j = lane; j < 6 * 8; j += 32) {
                    const int k = j >> 3, c = j & 7;
                    int dr = -1, sr = -1;
                    if (k < 3) { if (top) { dr = 2 - k; sr = 4 + k; } }                                       // rows -1-k <- rows 1+k
                    else if (bottom) { const int kk = k - 3; dr = h + kk - (y0 - 3); sr = h - 2 - kk - (y0 - 3); if (dr >= BOX || sr < 0) dr = -1; }
                    if (dr >= 0) *reinterpret_cast<uint4 *>(B + sw(dr, 16 * c)) = *reinterpret_cast<const uint4 *>(B + sw(sr, 16 * c));
                }
                __syncwarp();
            }
            if (left || right) {
                for (int i = lane; i < BOX; i += 32) {
                    if (left) { B[sw(i, 15)] = B[sw(i, 17)]; B[sw(i, 14)] = B[sw(i, 18)]; B[sw(i, 13)] = B[sw(i, 19)]; }   // x = -1,-2,-3 <- 1,2,3
                    if (right) {
                        const int c = w - x0 + 16;                                                            // box column of image column w
#pragma unroll
                        for (int k = 0; k < 3; k++) if (c + k < BOX) B[sw(i, c + k)] = B[sw(i, c - 2 - k)];   // x = w + k <- w - 2 - k
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.patched[s]);
        }
    } else {
        // ===== epilogue: thread = output row (TMEM lane), horizontal taps in registers =====
        const int quarter = warp & 3, row = quarter * 32 + lane;
        constexpr uint32_t W01 = 18u | (34u << 8) | (48u << 16) | (56u << 24), W45 = 48u | (34u << 8) | (18u << 16);
        uint32_t it = 0;
        for (int ti = first; ti < total; ti += step, it++) {
            const int s = it % NACC, ph = (it / NACC) & 1;
            int f; const BlurTile t = tile_of(ti, f);
            const int w = P.w[t.level], x0 = t.tx * TC;
            const int nchunk = min(3, (w - x0 + 31) >> 5);      // 32-column chunks that hold image columns
            uint8_t *out = sO + (it & 1) * O_BYTES;
            // the TMA store that read this output stage two tiles ago must have finished reading shared memory
            if (warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            asm volatile("bar.sync 1, 128;" ::: "memory");
            tma_mbar_wait(&S.acc_full[s], ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int c = 0; c < nchunk; c++) {
                // outputs j = 32c .. 32c+31 (image column x0 + j) need the column sums of box columns j + 13 .. j + 19
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(s * 128 + 32 * c + 13);
                uint32_t V[40];
                TMEM_LD32(taddr, V);
                TMEM_LD8(taddr + 32, (V + 32));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (c == nchunk - 1) {               // last read of this accumulator stage
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&S.acc_empty[s]);
                }
                uint32_t Pp[19], Q[19];
#pragma unroll
                for (int k = 0; k < 19; k++) { Pp[k] = __byte_perm(V[2 * k], V[2 * k + 1], 0x5410); Q[k] = __byte_perm(V[2 * k + 1], V[2 * k + 2], 0x5410); }
                uint32_t px[8];
#pragma unroll
                for (int g = 0; g < 8; g++) {
                    uint32_t a[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const int j = 4 * g + e, k = j >> 1;
                        const uint32_t *R = (j & 1) ? Q : Pp;
                        uint32_t acc = __dp2a_lo(R[k], W01, 32768u);
                        acc = __dp2a_hi(R[k + 1], W01, acc);
                        acc = __dp2a_lo(R[k + 2], W45, acc);
                        a[e] = __dp2a_hi(R[k + 3], W45, acc);
                    }
                    px[g] = __byte_perm(__byte_perm(a[0], a[1], 0x0062), __byte_perm(a[2], a[3], 0x0062), 0x5410);
                }
                if (row < TR) {
                    uint4 *o4 = reinterpret_cast<uint4 *>(out + row * TC + 32 * c);
                    o4[0] = make_uint4(px[0], px[1], px[2], px[3]);
                    o4[1] = make_uint4(px[4], px[5], px[6], px[7]);
                }
            }
            // the tile leaves by one TMA store; rows / columns beyond the plane are clipped by the tensor map
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (warp == 2 && lane == 0) {
                asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                             ::"l"(&P.omap[t.level]), "r"(x0), "r"(t.ty * TR), "r"(f0 + f), "r"(s32(out)) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        if (warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(128 * NACC) : "memory");
    }
}

}  // namespace

int launch_blur_tc(const LevelDev *h_levels, const BlurTc &C, int f0, int batch, cudaStream_t stream, int sm_count) {
    if (C.ntiles <= 0 || batch <= 0) return 0;
    static_assert(sizeof(BlurTc::map) == sizeof(BlurTcParams::map), "tensor map storage mismatch");
    BlurTcParams P;
    memcpy(P.map, C.map, sizeof(P.map));
    memcpy(P.omap, C.omap, sizeof(P.omap));
    for (int l = 0; l < kMaxLevels; l++) { P.w[l] = h_levels[l].w; P.h[l] = h_levels[l].h; }
    // Measured on 64 x 640x480 (blur stage alone / four batches in flight): 2 input stages, 1 CTA per SM 56 us / 189.4 k frames/s; 2 CTAs per SM
    // 43 us / 186.5 k; 3 stages change nothing (the epilogue warps, not the box loads, set the pace); k_blur_tma 47 us / 185.8 k.  One CTA per
    // SM leaves room for the other frame ranges' kernels; 1080p-class frames run in fewer, longer ranges and take two.
    static const int nb = [] { const char *e = getenv("ORBX_BLUR_TC_STAGES"); const int v = e ? atoi(e) : 2; return v < 2 ? 2 : v > NBMAX ? NBMAX : v; }();
    static const int per_sm_env = getenv("ORBX_BLUR_TC_CTAS") ? atoi(getenv("ORBX_BLUR_TC_CTAS")) : 0;   // 256 TMEM columns each
    const int per_sm = per_sm_env > 0 ? per_sm_env : ((long long)h_levels[0].w * h_levels[0].h >= 1500000 ? 2 : 1);
    const size_t smem = A_BYTES + (size_t)nb * B_BYTES + 2 * O_BYTES + sizeof(Ctl) + 1024;
    static bool configured_[kMaxDevices];
    {
        std::lock_guard<std::mutex> lock(g_attr_mutex);
        bool &configured = configured_[current_device_slot()];
        if (!configured) {
            const int big = (int)(A_BYTES + NBMAX * B_BYTES + 2 * O_BYTES + sizeof(Ctl) + 1024);
            if (cudaFuncSetAttribute(k_blur_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big) != cudaSuccess ||
                cudaFuncSetAttribute(k_blur_tc<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, big) != cudaSuccess ||
                cudaFuncSetAttribute(k_blur_tc<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, big) != cudaSuccess) {
                cudaGetLastError();
                return 0;                 // the caller takes k_blur_tma instead
            }
            configured = true;
        }
    }
    const int total = C.ntiles * batch;
    const int grid = total < sm_count * per_sm ? total : sm_count * per_sm;
    if (nb == 2) k_blur_tc<2><<<grid, THREADS, smem, stream>>>(P, C.d_tiles, C.ntiles, total, f0);
    else if (nb == 3) k_blur_tc<3><<<grid, THREADS, smem, stream>>>(P, C.d_tiles, C.ntiles, total, f0);
    else k_blur_tc<4><<<grid, THREADS, smem, stream>>>(P, C.d_tiles, C.ntiles, total, f0);
    return 1;
}

}  // namespace orbx

// Brute-force Hamming kNN (k = 2) on the 5th-generation tensor cores, FP4 variant (tcgen05.mma kind::mxf4, sm_100a).
//
// Same contraction as orbx_knn_tc.cu -- a 256-bit descriptor becomes 256 values in {-1, +1}, q . d = 256 - 2 * Hamming(q, d) -- but
// the values are E2M1 nibbles (+1.0 = 0x2, -1.0 = 0xA: exact in FP4), 128 bytes per descriptor instead of 256, and the tiles run
// through the block-scaled FP4 MMA, which issues at twice the int8 rate (ncu on the int8 kernel: tensor pipe 79 % busy on the
// 10 M-row shard, i.e. that kernel sits at its pipe's ceiling).  Every block scale is 1.0: the scale-factor columns of TMEM are
// filled with UE8M0 0x7F once, so their layout never matters.  Accumulators are fp32 (exact: integers up to 256).
//     D[128 queries x 224 rows] = Q[128 x 256] * DB[224 x 256]^T      (4 MMAs of K = 64 per tile, accumulators in TMEM)
// Persistent CTAs (one per SM): a CTA owns database tiles t = cta, cta + grid, ...; per database tile it streams all
// query tiles (<= 16 x 128 queries).  Warp roles: warp 0 = TMA producer (cp.async.bulk.tensor, SWIZZLE_128B K-major
// tiles), warp 1 = MMA issuer (8 x tcgen05.mma of K = 32 per tile, tcgen05.commit -> mbarrier), warps 2..5 = epilogue
// (tcgen05.ld 32x32b.x64, one TMEM lane = one query per thread).  The epilogue keeps a running top-2 per query in
// shared memory as (H << 23 | row-in-CTA) keys; a group of 8 dot products only enters the 3-op top-2 update when its
// maximum beats the current second best, so the steady state costs ~0.6 ALU ops per pair and the kernel is bound by
// the tensor pipe.  Two TMEM accumulator stages (2 x 256 columns) overlap MMA and epilogue.
// Results per CTA go to partial[cta][query][2] = (H << 32 | global row) and are merged by k_knn2_merge.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <mutex>
#include <cstdlib>
#include <string>

#include "orbx_dev.h"

namespace orbx {
namespace fp4 {

constexpr int TC_QT = 128;            // queries per tile (UMMA M)
constexpr int TC_DT = 224;            // database rows per tile (UMMA N): 2 x 224 accumulator columns + 64 scale-factor columns = all of TMEM
constexpr int TC_K = 128;             // expanded descriptor length in bytes (256 E2M1 nibbles): exactly one 128-byte swizzle row
constexpr int TC_MAX_QTILES = 16;     // 2048 queries per launch
constexpr int TC_EPI_PARTS = 4;         // epilogue warps per TMEM lane quarter: each takes TC_DT / 4 = 56 accumulator columns
constexpr int TC_THREADS = 64 + 4 * TC_EPI_PARTS * 32;   // warp 0 TMA, warp 1 MMA, then 4 x TC_EPI_PARTS epilogue warps
constexpr uint32_t TC_SENTINEL = 0x7FFFFFu;   // row field of a seeded (virtual) key
constexpr int TC_NA = 4;               // query-tile stages: a stage is refilled only after its MMAs retire, and the refill takes one TMA
                                       // latency from L2 -- two stages left the tensor pipe waiting for query tiles
constexpr uint32_t TC_A_BYTES = TC_QT * TC_K;          // 16 KB per query tile (one 128 x 128 B swizzle box)
constexpr uint32_t TC_B_BYTES = TC_DT * TC_K;          // 28 KB per database tile (one 224 x 128 B box)
constexpr uint32_t TC_SF_COL = 2 * TC_DT;              // TMEM columns 448 .. 511: scale factors (A at 448, B at 480)
constexpr unsigned long long TC_NONE64 = ~0ull;

// ---- expansion: packed bits -> E2M1 nibbles (+1.0 = 0x2, -1.0 = 0xA), 8 output bytes per thread ---------------------------
__global__ void __launch_bounds__(256) k_expand_fp4(const uint8_t *__restrict__ bits, long long nrows, long long nrows_pad,
                                                    uint8_t *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one thread = 16 bits = 2 input bytes
    if (i >= nrows_pad * 16) return;
    const long long row = i >> 4;
    uint32_t w[2] = {0, 0};
    if (row < nrows) {
        const uint32_t b = (uint32_t)bits[2 * i] | ((uint32_t)bits[2 * i + 1] << 8);
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const uint32_t v = ((b >> k) & 1u) ? 0x2u : 0xAu;     // bit k of byte j is descriptor bit 8j + k (LSB first); element k in nibble k
            w[k >> 3] |= v << (4 * (k & 7));
        }
    }
    reinterpret_cast<uint2 *>(out)[i] = make_uint2(w[0], w[1]);   // padded rows are all zero (+0.0: dot = 0)
}

// ---- PTX helpers -----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one lane of a converged warp; the role loops below run warp-uniform (so their addresses and descriptors stay in uniform registers and
// each tcgen05 / TMA instruction is a single issue, not a per-thread loop) and only the issue itself is predicated on the elected lane
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (rows of 128 B, 8-row groups 1024 B apart)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_mxf4(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate, uint32_t tsfa,
                                          uint32_t tsfb) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(tsfa), "r"(tsfb) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
#define TMEM_LD64(taddr, v)                                                                                              \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 "                                                               \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                               \
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "                      \
                 "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "                      \
                 "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"               \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),       \
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), \
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), \
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), \
                   "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), \
                   "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), \
                   "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), \
                   "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63]) \
                 : "r"(taddr) : "memory")

#define TMEM_LD32(taddr, v)                                                                                              \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                               \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                               \
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"               \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),       \
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), \
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), \
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) \
                 : "r"(taddr) : "memory")
#define TMEM_LD8(taddr, v)                                                                                               \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"                          \
                 : "=r"((v)[0]), "=r"((v)[1]), "=r"((v)[2]), "=r"((v)[3]), "=r"((v)[4]), "=r"((v)[5]), "=r"((v)[6]), "=r"((v)[7]) \
                 : "r"(taddr) : "memory")
#define TMEM_LD16(taddr, v)                                                                                              \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                               \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"                         \
                 : "=r"((v)[0]), "=r"((v)[1]), "=r"((v)[2]), "=r"((v)[3]), "=r"((v)[4]), "=r"((v)[5]), "=r"((v)[6]), "=r"((v)[7]), \
                   "=r"((v)[8]), "=r"((v)[9]), "=r"((v)[10]), "=r"((v)[11]), "=r"((v)[12]), "=r"((v)[13]), "=r"((v)[14]), "=r"((v)[15]) \
                 : "r"(taddr) : "memory")

struct TcShared {
    uint64_t full_a[TC_NA], empty_a[TC_NA], full_b[2], empty_b[2], tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
    uint32_t top[TC_MAX_QTILES][TC_EPI_PARTS][TC_QT][2];   // [query tile][column part][row]: running (H << 23 | row-in-CTA) keys
};

__global__ void __launch_bounds__(TC_THREADS, 1) k_knn2_fp4(const __grid_constant__ CUtensorMap map_q,
                                                           const __grid_constant__ CUtensorMap map_db, int nq, int nqt,
                                                           long long nrows, int tile0, int ntiles, long long row_offset, int qsplit,
                                                           const unsigned long long *__restrict__ seed,
                                                           unsigned long long *__restrict__ partial, unsigned int *gthr) {
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment: round the dynamic shared-memory base up (1 KB of slack is requested)
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    // carve-up: A stages (TC_NA x 16 KB), B stages (2 x 28 KB), then barriers + top-2 state
    uint8_t *smem_a = smem;
    uint8_t *smem_b = smem + TC_NA * TC_A_BYTES;
    TcShared &S = *reinterpret_cast<TcShared *>(smem + TC_NA * TC_A_BYTES + 2 * TC_B_BYTES);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp index, provably uniform
    // qsplit consecutive CTAs share one database-tile slot and split the query tiles between them (the seeding pass: one query tile per
    // CTA, so that its latency is one tile, not sixteen); qsplit = 1 is the main pass: every CTA takes all query tiles of its tiles
    const int cta = blockIdx.x / qsplit, G = gridDim.x / qsplit, qpart = blockIdx.x % qsplit;
    const int q_per = (nqt + qsplit - 1) / qsplit, q_lo = qpart * q_per, q_hi = min(nqt, q_lo + q_per);
    const int my_tiles = (cta < ntiles && q_lo < q_hi) ? (ntiles - cta + G - 1) / G : 0;

    // running top-2 state; with a seed (top-2 of an earlier pass over lower rows) only strictly closer rows can matter,
    // so both slots start at the virtual key (seed second-best distance, sentinel row)
    for (int i = threadIdx.x; i < TC_MAX_QTILES * TC_EPI_PARTS * TC_QT; i += TC_THREADS) {
        const int q = i / (TC_EPI_PARTS * TC_QT), r = i % TC_QT, qi = q * TC_QT + r;
        uint32_t v = 0xFFFFFFFFu;
        if (seed && qi < nq) {
            const unsigned long long s2 = seed[2 * (size_t)qi + 1];
            if (s2 != TC_NONE64) v = ((uint32_t)(s2 >> 32) << 23) | TC_SENTINEL;
        }
        (&S.top[0][0][0][0])[2 * i] = v; (&S.top[0][0][0][0])[2 * i + 1] = v;
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_NA; s++) { mbar_init(&S.full_a[s], 1); mbar_init(&S.empty_a[s], 1); }
        for (int s = 0; s < 2; s++) {
            mbar_init(&S.full_b[s], 1); mbar_init(&S.empty_b[s], 1);
            mbar_init(&S.tmem_full[s], 1); mbar_init(&S.tmem_empty[s], 4 * TC_EPI_PARTS);   // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // TMEM: all 512 columns (two 256-column accumulator stages)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&S.tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = S.tmem_base;
    // every block scale is 1.0: fill the 64 scale-factor columns of all 128 lanes with UE8M0 0x7F (each epilogue warp its lane quarter)
    if (warp >= 2 && warp < 6) {
        const uint32_t one4 = 0x7F7F7F7Fu;
        const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + TC_SF_COL;
#pragma unroll
        for (int c = 0; c < 64; c += 8)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr + (uint32_t)c), "r"(one4) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    if (warp == 0) {
        // ===== TMA producer (whole warp in the loop, one elected lane issues) =====
        uint32_t it = 0;
        for (int k = 0; k < my_tiles; k++) {
            const int t = tile0 + cta + k * G, bs = k & 1;
            mbar_wait(&S.empty_b[bs], ((k >> 1) & 1) ^ 1);
            if (elect_one()) {
                mbar_expect_tx(&S.full_b[bs], TC_B_BYTES);
                tma_load_2d(smem_b + bs * TC_B_BYTES, &map_db, 0, t * TC_DT, &S.full_b[bs]);
            }
            for (int q = q_lo; q < q_hi; q++, it++) {
                const int as = it % TC_NA;
                mbar_wait(&S.empty_a[as], ((it / TC_NA) & 1) ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&S.full_a[as], TC_A_BYTES);
                    tma_load_2d(smem_a + as * TC_A_BYTES, &map_q, 0, q * TC_QT, &S.full_a[as]);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer (whole warp in the loop, one elected lane issues) =====
        // block-scaled instruction descriptor (cute::UMMA::InstrDescriptorBlockScaled): A = B = E2M1 (MXF4 format 1) at bits 7 / 10, both
        // K-major, N >> 3 at bit 17, scale format UE8M0 at bit 23, M >> 4 at bit 24, scale-factor ids 0, K = 64 per instruction
        const uint32_t idesc = (1u << 7) | (1u << 10) | ((uint32_t)(TC_DT >> 3) << 17) | (1u << 23) | ((uint32_t)(TC_QT >> 4) << 24);
        const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t tsfa = tb + TC_SF_COL, tsfb = tb + TC_SF_COL + 32;
        uint32_t it = 0;
        for (int k = 0; k < my_tiles; k++) {
            const int bs = k & 1;
            mbar_wait(&S.full_b[bs], (k >> 1) & 1);
            const uint64_t db = umma_desc_sw128(smem_u32(smem_b + bs * TC_B_BYTES));
            for (int q = q_lo; q < q_hi; q++, it++) {
                const int as = it % TC_NA, acc = it & 1;
                mbar_wait(&S.full_a[as], (it / TC_NA) & 1);
                mbar_wait(&S.tmem_empty[acc], ((it >> 1) & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d = tb + acc * TC_DT;
                const uint64_t da = umma_desc_sw128(smem_u32(smem_a + as * TC_A_BYTES));
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 4; ks++)   // 64 elements = 32 bytes of K per MMA: advance the start address by 2 x 16 B
                        umma_mxf4(d, da + 2 * ks, db + 2 * ks, idesc, ks ? 1u : 0u, tsfa, tsfb);
                    umma_commit(&S.empty_a[as]);        // A stage reusable once these MMAs retire
                    umma_commit(&S.tmem_full[acc]);     // accumulator ready for the epilogue
                    if (q == q_hi - 1) umma_commit(&S.empty_b[bs]);
                }
                __syncwarp();
            }
        }
    } else {
        // ===== epilogue: warps 2 ..; TMEM lane quarter = warp % 4, column part = (warp - 2) / 4 =====
        const int quarter = warp & 3, half = (warp - 2) >> 2;   // `half` = column part 0 .. TC_EPI_PARTS - 1
        const int row_in_tile = quarter * 32 + lane;
        uint32_t it = 0;
        for (int k = 0; k < my_tiles; k++) {
            const int t = tile0 + cta + k * G;
            const long long tile_row0 = (long long)t * TC_DT;
            const int valid_cols = (int)min((long long)TC_DT, nrows - tile_row0);
            for (int q = q_lo; q < q_hi; q++, it++) {
                const int acc = it & 1;
                uint32_t k1 = S.top[q][half][row_in_tile][0], k2 = S.top[q][half][row_in_tile][1];
                const uint32_t k2_in = k2;
                // thresholds shared between the CTAs: gthr[query] = the smallest second-best distance any CTA has published.  A row of the final
                // top 2 has a distance <= that (ties with rows of other CTAs are decided by row number later, so "<=" here, "<" against this
                // CTA's own second best, whose rows come first).  The load is in flight during the wait for the accumulator.
                unsigned int *gq = gthr + q * TC_QT + row_in_tile;
                const float thr_g = (float)(255 - 2 * (int)min(__ldcg(gq), 511u));
                float thr = fmaxf((float)(256 - 2 * (int)(k2 >> 23)), thr_g);   // a dot product must exceed this to enter the top 2
                mbar_wait(&S.tmem_full[acc], (it >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                constexpr int PW = TC_DT / TC_EPI_PARTS;     // 56 columns per warp: one x32, one x16 and one x8 load
                const uint32_t tcol = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * TC_DT + half * PW);
                uint32_t v[PW];
                TMEM_LD32(tcol, v);
                TMEM_LD16(tcol + 32, (v + 32));
                TMEM_LD8(tcol + 48, (v + 48));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                // the accumulator stage is free as soon as every epilogue warp holds its columns in registers: the next-but-one MMA
                // runs under the scan below
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&S.tmem_empty[acc]);
                // fast path: one three-input max tree over the warp's 56 accumulators (fp32, exact integers) and one branch; only a
                // row whose maximum beats the threshold looks at its groups of 8
                float gm[PW / 8];
#pragma unroll
                for (int g = 0; g < PW / 8; g++) {
                    const float m0 = fmaxf(fmaxf(__uint_as_float(v[8 * g]), __uint_as_float(v[8 * g + 1])), __uint_as_float(v[8 * g + 2]));
                    const float m1 = fmaxf(fmaxf(__uint_as_float(v[8 * g + 3]), __uint_as_float(v[8 * g + 4])), __uint_as_float(v[8 * g + 5]));
                    gm[g] = fmaxf(fmaxf(__uint_as_float(v[8 * g + 6]), __uint_as_float(v[8 * g + 7])), fmaxf(m0, m1));
                }
                const float mall = fmaxf(fmaxf(fmaxf(fmaxf(gm[0], gm[1]), gm[2]), fmaxf(fmaxf(gm[3], gm[4]), gm[5])), gm[6]);
                if (mall > thr) {
#pragma unroll
                    for (int g = 0; g < PW / 8; g++) {
                        if (gm[g] > thr) {
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                const int col = half * PW + 8 * g + j;
                                if (col < valid_cols) {
                                    // H = 128 - dot / 2 (an integer 0 .. 256) lands in the low mantissa bits of H + 2^23: one FFMA instead of a
                                    // float -> int conversion (quarter rate); the shift by 23 drops the exponent bits
                                    const uint32_t hb = __float_as_uint(__fmaf_rn(__uint_as_float(v[8 * g + j]), -0.5f, 128.f + 8388608.f));
                                    const uint32_t key = (hb << 23) | (uint32_t)(k * TC_DT + col);
                                    k2 = min(k2, max(k1, key)); k1 = min(k1, key);
                                }
                            }
                            thr = fmaxf((float)(256 - 2 * (int)(k2 >> 23)), thr_g);
                        }
                    }
                    if (k2 != k2_in && (k2 & 0x7FFFFFu) != TC_SENTINEL) atomicMin(gq, k2 >> 23);
                }
                S.top[q][half][row_in_tile][0] = k1; S.top[q][half][row_in_tile][1] = k2;
            }
        }
        // merge the two column halves and write this CTA's partial result
        asm volatile("bar.sync 1, %0;" ::"n"(4 * TC_EPI_PARTS * 32) : "memory");
        if (half == 0) {
            for (int q = 0; q < nqt; q++) {
                const int qi = q * TC_QT + row_in_tile;
                if (qi < nq) {
                    uint32_t k1 = S.top[q][0][row_in_tile][0], k2 = S.top[q][0][row_in_tile][1];
#pragma unroll
                    for (int pp = 1; pp < TC_EPI_PARTS; pp++)
#pragma unroll
                        for (int j = 0; j < 2; j++) {
                            const uint32_t key = S.top[q][pp][row_in_tile][j];
                            k2 = min(k2, max(k1, key)); k1 = min(k1, key);
                        }
                    unsigned long long *o = partial + ((size_t)blockIdx.x * nq + qi) * 2;
                    const uint32_t keys[2] = {k1, k2};
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        const uint32_t key = keys[j], local = key & 0x7FFFFFu;
                        if (key == 0xFFFFFFFFu || local == TC_SENTINEL) { o[j] = TC_NONE64; continue; }
                        const uint32_t lt = local / (uint32_t)TC_DT, lc = local - lt * (uint32_t)TC_DT;
                        const long long grow = row_offset + (long long)(tile0 + cta + (int)lt * G) * TC_DT + lc;
                        o[j] = ((unsigned long long)(key >> 23) << 32) | (unsigned long long)grow;
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// 2D u8 tensor [rows][128 B], box = 128 B x box_rows, SWIZZLE_128B
static bool make_map(CUtensorMap *map, const void *base, long long rows, int box_rows, std::string &err) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) { err = "cuTensorMapEncodeTiled entry point not available"; return false; }
    cuuint64_t dims[2] = {(cuuint64_t)TC_K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)TC_K};
    cuuint32_t box[2] = {(cuuint32_t)TC_K, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { err = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")"; return false; }
    return true;
}

size_t knn_fp4_smem_bytes() { return TC_NA * TC_A_BYTES + 2 * TC_B_BYTES + sizeof(TcShared) + 1024; }
int knn_fp4_max_queries() { return TC_MAX_QTILES * TC_QT; }
long long knn_fp4_padded_rows(long long nrows) { return (nrows + TC_DT - 1) / TC_DT * TC_DT; }
int knn_fp4_padded_queries(int nq) { return (nq + TC_QT - 1) / TC_QT * TC_QT; }

int launch_expand_fp4(const uint8_t *d_bits, long long nrows, long long nrows_pad, uint8_t *d_out, cudaStream_t stream) {
    const long long nthreads = nrows_pad * 16;
    if (nthreads <= 0) return 0;
    k_expand_fp4<<<(unsigned)((nthreads + 255) / 256), 256, 0, stream>>>(d_bits, nrows, nrows_pad, d_out);
    return 1;
}

// d_qe: expanded queries [nq_pad][128 B], d_dbe: expanded database [rows_pad][128 B]; partial: [sm_count + 1][nq][2].
// Two passes when the shard has more tiles than CTAs: pass A takes the top-2 over the first `grid` tiles (merged into
// partial slot `grid` by the caller-supplied merge), pass B covers the rest seeded with pass A's second-best distances,
// which keeps its epilogue on the cheap filter path.  Returns launches issued (0 = failure, err set);
// *nparts_out = number of partial blocks to merge at the end.
int launch_knn2_fp4(const uint8_t *d_qe, int nq, const uint8_t *d_dbe, long long nrows, long long row_offset, int sm_count,
                   unsigned long long *d_partial, int *nparts_out, void (*merge)(const unsigned long long *, int, int, unsigned long long *, cudaStream_t),
                   cudaStream_t stream, std::string &err) {
    const int nqt = (nq + TC_QT - 1) / TC_QT;
    if (nqt < 1 || nqt > TC_MAX_QTILES) { err = "tensor-core kNN handles 1..2048 queries per launch"; return 0; }
    const long long ntiles_ll = (nrows + TC_DT - 1) / TC_DT;
    if (ntiles_ll < 1 || ntiles_ll > (1ll << 30)) { err = "bad database size"; return 0; }
    const int ntiles = (int)ntiles_ll;
    const int grid = ntiles < sm_count ? ntiles : sm_count;
    if ((long long)((ntiles + grid - 1) / grid + 1) * TC_DT >= (1ll << 23) - TC_DT) { err = "database shard too large for the packed key"; return 0; }
    CUtensorMap mq, mdb;
    if (!make_map(&mq, d_qe, (long long)nqt * TC_QT, TC_QT, err)) return 0;
    if (!make_map(&mdb, d_dbe, ntiles_ll * TC_DT, TC_DT, err)) return 0;
    static bool configured_[kMaxDevices];
    static std::mutex attr_mutex;
    std::lock_guard<std::mutex> attr_lock(attr_mutex);
    bool &configured = configured_[current_device_slot()];
    const size_t smem = knn_fp4_smem_bytes();
    if (!configured) {
        if (cudaFuncSetAttribute(k_knn2_fp4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            err = "cannot reserve shared memory for the tensor-core kNN kernel"; cudaGetLastError(); return 0;
        }
        configured = true;
    }
    unsigned int *gthr = reinterpret_cast<unsigned int *>(d_partial + (size_t)(sm_count + 1) * nq * 2);
    cudaMemsetAsync(gthr, 0xFF, (size_t)TC_MAX_QTILES * TC_QT * sizeof(unsigned int), stream);
    if (ntiles <= 2 * grid) {   // small shard: one pass
        k_knn2_fp4<<<grid, TC_THREADS, smem, stream>>>(mq, mdb, nq, nqt, nrows, 0, ntiles, row_offset, 1, nullptr, d_partial, gthr);
        *nparts_out = grid;
        return 1;
    }
    unsigned long long *seed = d_partial + (size_t)grid * nq * 2;
    // seeding pass: (query tile, database tile) pairs spread over the CTAs, one pair each -- t1 tiles x nqt query tiles <= grid
    static const bool split_seed = getenv("ORBX_KNN_SEED_SPLIT") && atoi(getenv("ORBX_KNN_SEED_SPLIT")) != 0;
    const bool sp = split_seed && grid / nqt > 0;
    const int t1 = sp ? grid / nqt : 1, qs = sp ? nqt : 1, g1 = sp ? t1 * nqt : grid;
    const int first = sp ? t1 : grid;
    k_knn2_fp4<<<g1, TC_THREADS, smem, stream>>>(mq, mdb, nq, nqt, nrows, 0, first, row_offset, qs, nullptr, d_partial, gthr);
    merge(d_partial, g1, nq, seed, stream);
    k_knn2_fp4<<<grid, TC_THREADS, smem, stream>>>(mq, mdb, nq, nqt, nrows, first, ntiles - first, row_offset, 1, seed, d_partial, gthr);
    *nparts_out = grid + 1;
    return 3;
}

}  // namespace fp4

size_t knn_fp4_smem_bytes() { return fp4::knn_fp4_smem_bytes(); }
int knn_fp4_max_queries() { return fp4::knn_fp4_max_queries(); }
long long knn_fp4_padded_rows(long long nrows) { return fp4::knn_fp4_padded_rows(nrows); }
int knn_fp4_padded_queries(int nq) { return fp4::knn_fp4_padded_queries(nq); }
int launch_expand_fp4(const uint8_t *d_bits, long long nrows, long long nrows_pad, uint8_t *d_out, cudaStream_t stream) {
    return fp4::launch_expand_fp4(d_bits, nrows, nrows_pad, d_out, stream);
}
int launch_knn2_fp4(const uint8_t *d_qe, int nq, const uint8_t *d_dbe, long long nrows, long long row_offset, int sm_count, unsigned long long *d_partial,
                    int *nparts_out, void (*merge)(const unsigned long long *, int, int, unsigned long long *, cudaStream_t), cudaStream_t stream,
                    std::string &err) {
    return fp4::launch_knn2_fp4(d_qe, nq, d_dbe, nrows, row_offset, sm_count, d_partial, nparts_out, merge, stream, err);
}

}  // namespace orbx

// Probe for the next step on k_blur: the same 7x7 fixed-point Gaussian with its halo tile staged
//   A) by aligned word loads at clamped columns + shared-memory edge patches (what send_slam_b200/csrc/orbx_kernels.cu does), or
//   B) by ONE TMA box per tile (cp.async.bulk.tensor.3d, 96 x 118 bytes starting at x0 - 16, zero fill outside the plane) followed by
//      the REFLECT_101 patches for rows and columns in shared memory,
// and the four accumulators of a pixel quad packed with byte permutes.  Both variants are checked against a CPU restatement of the
// filter inside this file (no oracle, no library) and timed on [frames][h][w] planes.
// NOT VERIFIED ON A GPU YET: written at the end of round 1 after the GPU budget was spent; it compiles for sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I send_slam_b200/csrc -o tools/_build/blur_tma_probe tools/blur_tma_probe.cu
// Run:   tools/_build/blur_tma_probe [w h frames]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "orbx_tma.cuh"
using namespace orbx;

constexpr int TW = 64, TH = 112;            // output tile
constexpr int PA = 80;                      // staged row pitch, variant A: columns x0-4 .. x0+75
constexpr int PB = 96;                      // staged row pitch, variant B: columns x0-16 .. x0+79 (TMA boxes start on 16 bytes)
constexpr int ROWS = TH + 6;

__device__ __forceinline__ int reflect1(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

// Horizontal + vertical pass on a staged tile.  `col0` = staged byte column of image column x0 - 4 (0 for A, 12 for B).
template <int PITCH>
__device__ __forceinline__ void blur_passes(const uint8_t *s_in, uint32_t *s_h, int col0, int srows, int nseg, int x0, int y0, int w, int h,
                                            uint8_t *__restrict__ dst, int dpitch) {
    const int npairs = srows >> 1;
    for (int it = threadIdx.x; it < npairs * 8; it += 256) {
        const int p = it >> 3, g = it & 7;
        constexpr uint32_t KA = 18u | (34u << 8) | (48u << 16) | (56u << 24), KB = 48u | (34u << 8) | (18u << 16);
        uint32_t hs[2][8];
#pragma unroll
        for (int rr = 0; rr < 2; rr++) {
            const uint32_t *q = reinterpret_cast<const uint32_t *>(s_in + (2 * p + rr) * PITCH + col0 + 8 * g);
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
        uint4 *d4 = reinterpret_cast<uint4 *>(s_h + p * TW + 8 * g);
        d4[0] = make_uint4(o[0], o[1], o[2], o[3]);
        d4[1] = make_uint4(o[4], o[5], o[6], o[7]);
    }
    __syncthreads();
    for (int it = threadIdx.x; it < 16 * nseg; it += 256) {
        const int cg = it & 15, seg = it >> 4;
        const int gx = x0 + 4 * cg;
        if (gx >= w) continue;
        constexpr uint32_t EA = 18u | (34u << 8) | (48u << 16) | (56u << 24), EB = 48u | (34u << 8) | (18u << 16);
        constexpr uint32_t OA = (18u << 8) | (34u << 16) | (48u << 24), OB = 56u | (48u << 8) | (34u << 16) | (18u << 24);
        uint32_t P[5][4];
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const uint4 q = *reinterpret_cast<const uint4 *>(s_h + (2 * seg + k) * TW + 4 * cg);
            P[k][0] = q.x; P[k][1] = q.y; P[k][2] = q.z; P[k][3] = q.w;
        }
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int gy = y0 + 4 * seg + r, pb = r >> 1;
            const uint32_t ka = (r & 1) ? OA : EA, kb = (r & 1) ? OB : EB;
            uint32_t acc[4];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                uint32_t a = __dp2a_lo(P[pb][c], ka, 32768u);
                a = __dp2a_hi(P[pb + 1][c], ka, a);
                a = __dp2a_lo(P[pb + 2][c], kb, a);
                acc[c] = __dp2a_hi(P[pb + 3][c], kb, a);
            }
            // byte 2 of every accumulator is the pixel (acc < 2^24): three permutes for four pixels
            const uint32_t px = __byte_perm(__byte_perm(acc[0], acc[1], 0x0062), __byte_perm(acc[2], acc[3], 0x0062), 0x5410);
            if (gy < h) *reinterpret_cast<uint32_t *>(dst + (size_t)gy * dpitch + gx) = px;
        }
    }
}

// Variant A: word-load staging (planes of ordinary size: w, h >= 16, 4-byte aligned rows, >= 16 bytes of slack behind each row)
__global__ void __launch_bounds__(256) k_blur_a(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int w, int h, int pitch, size_t fstride,
                                                int ntx) {
    __shared__ __align__(16) uint8_t s_in[ROWS * PA];
    __shared__ __align__(16) uint32_t s_h[(ROWS / 2) * TW];
    const int tx = blockIdx.x % ntx, ty = blockIdx.x / ntx, f = blockIdx.y;
    const int x0 = tx * TW, y0 = ty * TH;
    src += (size_t)f * fstride; dst += (size_t)f * fstride;
    const int nseg = (min(TH, h - y0) + 3) >> 2, srows = 4 * nseg + 6;
    const int lastw = (w - 1) & ~3;
    const int tr = threadIdx.x / (PA / 4), wc = threadIdx.x - tr * (PA / 4);
    if (tr < 12) {
        const int cx = min(max(x0 - 4 + 4 * wc, 0), lastw);
        for (int r0 = tr; r0 < srows; r0 += 60) {
            uint32_t v[5];
#pragma unroll
            for (int k = 0; k < 5; k++) {
                const int r = r0 + 12 * k;
                if (r < srows) v[k] = __ldg(reinterpret_cast<const uint32_t *>(src + (size_t)reflect1(y0 - 3 + r, h) * pitch + cx));
            }
#pragma unroll
            for (int k = 0; k < 5; k++) {
                const int r = r0 + 12 * k;
                if (r < srows) reinterpret_cast<uint32_t *>(s_in)[r * (PA / 4) + wc] = v[k];
            }
        }
    }
    const bool left = x0 == 0, right = x0 + 76 > w;
    if (left || right) {
        __syncthreads();
        for (int r = threadIdx.x; r < srows; r += 256) {
            uint8_t *row = s_in + r * PA;
            if (left) { row[1] = row[7]; row[2] = row[6]; row[3] = row[5]; }
            if (right) {
                const int c = w - x0 + 4;
#pragma unroll
                for (int k = 0; k < 3; k++) if (c + k < PA) row[c + k] = row[c - 2 - k];
            }
        }
    }
    __syncthreads();
    blur_passes<PA>(s_in, s_h, 0, srows, nseg, x0, y0, w, h, dst, pitch);
}

// Variant B: one TMA box per tile.  Box = 96 x 118 x 1 at (x0 - 16, y0 - 3, f); bytes outside the plane arrive as 0 and are
// replaced by their REFLECT_101 sources: rows first (whole staged rows), then the three columns beyond a vertical edge.
__global__ void __launch_bounds__(256) k_blur_b(const __grid_constant__ CUtensorMap map, uint8_t *__restrict__ dst, int w, int h, int pitch,
                                                size_t fstride, int ntx) {
    __shared__ __align__(128) uint8_t s_in[ROWS * PB];
    __shared__ __align__(16) uint32_t s_h[(ROWS / 2) * TW];
    __shared__ __align__(8) uint64_t bar;
    const int tx = blockIdx.x % ntx, ty = blockIdx.x / ntx, f = blockIdx.y;
    const int x0 = tx * TW, y0 = ty * TH;
    dst += (size_t)f * fstride;
    const int nseg = (min(TH, h - y0) + 3) >> 2, srows = 4 * nseg + 6;
    if (threadIdx.x == 0) { tma_mbar_init(&bar, 1); tma_mbar_fence_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        tma_mbar_expect_tx(&bar, (uint32_t)(PB * ROWS));
        tma_load_3d(s_in, &map, x0 - 16, y0 - 3, f, &bar);
    }
    tma_mbar_wait(&bar, 0);
    // rows outside the plane <- their reflections (staged row r holds image row y0 - 3 + r)
    const bool top = y0 == 0, bottom = y0 - 3 + srows > h;
    if (top || bottom) {
        for (int i = threadIdx.x; i < 3 * (PB / 4); i += 256) {
            const int k = i / (PB / 4), wq = i - k * (PB / 4);
            uint32_t *S = reinterpret_cast<uint32_t *>(s_in);
            if (top) S[(2 - k) * (PB / 4) + wq] = S[(4 + k) * (PB / 4) + wq];                 // rows -1-k <- rows 1+k
            if (bottom) {
                const int r = h + k - (y0 - 3), s = h - 2 - k - (y0 - 3);                        // row h+k <- row h-2-k
                if (r < srows && s >= 0) S[r * (PB / 4) + wq] = S[s * (PB / 4) + wq];
            }
        }
        __syncthreads();
    }
    const bool left = x0 == 0, right = x0 + 76 > w;
    if (left || right) {
        for (int r = threadIdx.x; r < srows; r += 256) {
            uint8_t *row = s_in + r * PB + 12;                                                   // row[c]: image column x0 - 4 + c
            if (left) { row[1] = row[7]; row[2] = row[6]; row[3] = row[5]; }
            if (right) {
                const int c = w - x0 + 4;
#pragma unroll
                for (int k = 0; k < 3; k++) if (c + k < PB - 12) row[c + k] = row[c - 2 - k];
            }
        }
    }
    __syncthreads();
    blur_passes<PB>(s_in, s_h, 12, srows, nseg, x0, y0, w, h, dst, pitch);
}

static void cpu_blur(const uint8_t *src, uint8_t *dst, int w, int h, int pitch) {
    static const int K[7] = {18, 34, 48, 56, 48, 34, 18};
    auto refl = [](int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); };
    std::vector<uint32_t> H((size_t)w * h);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            uint32_t a = 0;
            for (int k = 0; k < 7; k++) a += K[k] * src[(size_t)y * pitch + refl(x + k - 3, w)];
            H[(size_t)y * w + x] = a;
        }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            uint32_t a = 32768;
            for (int k = 0; k < 7; k++) a += K[k] * H[(size_t)refl(y + k - 3, h) * w + x];
            dst[(size_t)y * pitch + x] = (uint8_t)(a >> 16);
        }
}

int main(int argc, char **argv) {
    const int w = argc > 1 ? atoi(argv[1]) : 640, h = argc > 2 ? atoi(argv[2]) : 480, F = argc > 3 ? atoi(argv[3]) : 64;
    const int pitch = (w + 31) / 32 * 32;
    const size_t fstride = (size_t)pitch * h, bytes = fstride * F + 64;
    std::vector<uint8_t> hsrc(bytes), want(bytes), got(bytes);
    for (size_t i = 0; i < bytes; i++) hsrc[i] = (uint8_t)((i * 2654435761u) >> 11);
    uint8_t *d_src, *d_a, *d_b;
    cudaMalloc(&d_src, bytes); cudaMalloc(&d_a, bytes); cudaMalloc(&d_b, bytes);
    cudaMemcpy(d_src, hsrc.data(), bytes, cudaMemcpyHostToDevice);
    CUtensorMap map;
    if (!tma_make_plane_map(&map, d_src, w, h, F, (size_t)pitch, fstride, PB, ROWS)) { printf("tensor map refused\n"); return 1; }
    const int ntx = (w + TW - 1) / TW, nty = (h + TH - 1) / TH;
    dim3 grid(ntx * nty, F);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](int variant, int reps) {
        cudaEventRecord(e0);
        for (int i = 0; i < reps; i++) {
            if (variant == 0) k_blur_a<<<grid, 256>>>(d_src, d_a, w, h, pitch, fstride, ntx);
            else k_blur_b<<<grid, 256>>>(map, d_b, w, h, pitch, fstride, ntx);
        }
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant %c: %s\n", 'A' + variant, cudaGetErrorString(e)); exit(1); }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        return 1e3f * ms / reps;
    };
    for (int f = 0; f < F; f += (F > 4 ? F / 4 : 1)) cpu_blur(hsrc.data() + f * fstride, want.data() + f * fstride, w, h, pitch);
    for (int variant = 0; variant < 2; variant++) {
        run(variant, 3);
        cudaMemcpy(got.data(), variant ? d_b : d_a, bytes, cudaMemcpyDeviceToHost);
        long bad = 0;
        for (int f = 0; f < F; f += (F > 4 ? F / 4 : 1))
            for (int y = 0; y < h; y++)
                for (int x = 0; x < w; x++) bad += got[f * fstride + (size_t)y * pitch + x] != want[f * fstride + (size_t)y * pitch + x];
        const float us = run(variant, 20);
        printf("variant %c (%s): %ld mismatching pixels, %.1f us per %d x %dx%d launch = %.2f us/Mpx\n", 'A' + variant,
               variant ? "TMA box + patches" : "word loads + patches", bad, us, F, w, h, us / (1e-6 * w * h * F));
    }
    return 0;
}

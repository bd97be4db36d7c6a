// Probe: which (rank, box) shapes of an un-swizzled u8 TMA tile load work on this GPU.  Build: nvcc -arch=sm_100a -I../send_slam_b200/csrc
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "orbx_tma.cuh"
using namespace orbx;

__device__ __forceinline__ void tma_load_2d_(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(tma_smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(tma_smem_u32(bar)) : "memory");
}

__global__ void k_probe(const __grid_constant__ CUtensorMap m, int rank, int x, int y, int z, int bytes, uint8_t *out) {
    __shared__ __align__(128) uint8_t s[128 * 80];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) { tma_mbar_init(&bar, 1); tma_mbar_fence_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        tma_mbar_expect_tx(&bar, bytes);
        if (rank == 3) tma_load_3d(s, &m, x, y, z, &bar); else tma_load_2d_(s, &m, x, y, &bar);
    }
    tma_mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = s[i];
}

int main(int argc, char **argv) {
    const int rank = argc > 1 ? atoi(argv[1]) : 3, swz = argc > 2 ? atoi(argv[2]) : 0, bw = argc > 3 ? atoi(argv[3]) : 48, bh = argc > 4 ? atoi(argv[4]) : 45;
    const int W = 267, H = 200, P = 288, F = 3;
    std::vector<uint8_t> h((size_t)P * H * F);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 2654435761u >> 13);
    uint8_t *d, *o; cudaMalloc(&d, h.size()); cudaMalloc(&o, 128 * 80);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    PFN_tmaEncodeTiled enc = tma_encode_fn();
    CUtensorMap m;
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)(rank == 3 ? H : H * F), (cuuint64_t)F};
    cuuint64_t strides[2] = {(cuuint64_t)P, (cuuint64_t)P * H};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1u};
    cuuint32_t estr[3] = {1u, 1u, 1u};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     (CUtensorMapSwizzle)swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("rank %d swz %d box %dx%d: encode failed %d\n", rank, swz, bw, bh, (int)r); return 0; }
    const int x = argc > 5 ? atoi(argv[5]) : 37, y = 21, z = 1;
    k_probe<<<1, 128>>>(m, rank, x, rank == 3 ? y : y + z * H, z, bw * bh, o);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("rank %d swz %d box %dx%d: %s\n", rank, swz, bw, bh, cudaGetErrorString(e)); return 1; }
    std::vector<uint8_t> rr(bw * bh); cudaMemcpy(rr.data(), o, rr.size(), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int j = 0; j < bh; j++) for (int i = 0; i < bw; i++) {
        const int gx = x + i, gy = y + j;
        const uint8_t want = (gx < W && gy < H) ? h[((size_t)z * H + gy) * P + gx] : 0;
        bad += rr[j * bw + i] != want;
    }
    printf("rank %d swz %d box %dx%d: ok, %d mismatches\n", rank, swz, bw, bh, bad);
    return 0;
}

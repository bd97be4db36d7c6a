// mbarrier / TMA (cp.async.bulk.tensor) helpers for the byte-tile kernels (sm_100a).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace orbx {

typedef CUresult (*PFN_tmaEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                       const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_tmaEncodeTiled tma_encode_fn() {
    static PFN_tmaEncodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = (PFN_tmaEncodeTiled)p;
    }
    return fn;
}

// u8 tensor [frames][rows][width] with byte strides (pitch, frame_stride); box = box_w x box_h x 1, no swizzle, OOB -> 0.
// Returns false when the geometry does not meet the TMA rules (16-byte aligned base and strides) or the driver refuses.
inline bool tma_make_plane_map(CUtensorMap *map, const void *base, int width, int rows, int frames, size_t pitch, size_t frame_stride,
                               int box_w, int box_h) {
    PFN_tmaEncodeTiled enc = tma_encode_fn();
    if (!enc) return false;
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (pitch & 15) || (frame_stride & 15) || (box_w & 15) || box_w > 256 || box_h > 256) return false;
    cuuint64_t dims[3] = {(cuuint64_t)width, (cuuint64_t)rows, (cuuint64_t)frames};
    cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)frame_stride};
    cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1u};
    cuuint32_t estr[3] = {1u, 1u, 1u};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// The same plane with SWIZZLE_128B and a 128-byte-wide box: the box lands in shared memory as the canonical MN-major (or K-major) operand
// tile of tcgen05.mma (128-byte rows, 16-byte pieces XOR-ed with the row number modulo 8).
inline bool tma_make_plane_map_sw128(CUtensorMap *map, const void *base, int width, int rows, int frames, size_t pitch, size_t frame_stride,
                                     int box_h) {
    PFN_tmaEncodeTiled enc = tma_encode_fn();
    if (!enc) return false;
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (pitch & 15) || (frame_stride & 15) || box_h > 256) return false;
    cuuint64_t dims[3] = {(cuuint64_t)width, (cuuint64_t)rows, (cuuint64_t)frames};
    cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)frame_stride};
    cuuint32_t box[3] = {128u, (cuuint32_t)box_h, 1u};
    cuuint32_t estr[3] = {1u, 1u, 1u};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t tma_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tma_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void tma_mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tma_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tma_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
        ::"r"(tma_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(tma_smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(tma_smem_u32(bar)) : "memory");
}
#endif

}  // namespace orbx

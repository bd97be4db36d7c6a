// Device-side descriptors shared by the extraction kernels and the host runtime.
#pragma once
#include <cstddef>
#include <cstdint>

#include <vector_types.h>

#include "orbx_plan.h"

namespace orbx {

// per-device bookkeeping of kernel attributes (cudaFuncSetAttribute is per device)
constexpr int kMaxDevices = 64;
int current_device_slot();

// One pyramid level as the kernels see it.  Planes are [batch][h][pitch] u8; level 0 may alias caller memory.
struct LevelDev {
    uint8_t *img;              // un-blurred plane (FAST, IC_Angle read this)
    uint8_t *blur;             // 7x7 Gaussian of img (steered BRIEF reads this)
    size_t img_fstride;        // bytes between consecutive frames in img
    size_t blur_fstride;
    int w, h, pitch, blur_pitch;
    int padded;                // img is a workspace plane with slack behind its rows (0 when level 0 aliases caller memory)
    const ResizeTap *xtap, *ytap;                   // taps from level l-1
    const uint32_t *xpack;                          // x taps as ofs << 16 | c1 (c0 = 2048 - c1), padded to 4; may be null
    const uint32_t *xbin, *ybin, *xord, *yord;      // quadtree / order LUTs, indexed by (x-16), (y-16)
    int reg_w, reg_h, n_ini, depth0, nbins, quota, out_cap;
    int root_ulx[kMaxRoots], root_brx[kMaxRoots];
    uint32_t ord_cell_area, ord_ncols;
    int wcell, hcell;
    uint32_t *cand;            // [batch][cand_cap]   (y-16)<<20 | (x-16)<<8 | score
    uint32_t *sorted;          // [batch][cand_cap]   bin-sorted copy, built only when the tree goes below depth0
    uint32_t *bin_cursor;      // [batch][nbins]      scratch for that copy
    int *cand_count;           // [batch]
    int cand_cap;
    uint32_t *sel;             // [batch][out_cap]    y<<20 | x<<8 | score, level coordinates, quadtree list order
    int *sel_count;            // [batch]
    int out_base;              // prefix sum of out_cap over lower levels
    float scale, kp_size;
};

constexpr int kBlurTileW = 64, kBlurTileH = 112;
struct BlurTile { int16_t level, tx, ty, pad; };   // one kBlurTileW x kBlurTileH output tile of the Gaussian pass

struct KeypointRec {           // == orbx_keypoint == cv::KeyPoint
    float x, y, size, angle, response;
    int32_t octave, class_id;
};

// launch wrappers (orbx_kernels.cu); all asynchronous on `stream`, return the number of kernel launches issued.
// They process frames [f0, f0 + batch) of the workspace.  h_levels = HOST copy of the level table (kMaxLevels entries): it is
// passed to the kernels by value in the parameter bank (no device copy exists).
int launch_gray(const uint8_t *d_src, size_t src_fstride, int src_pitch, int format, int shift, uint8_t *d_dst, size_t dst_fstride,
                int dst_pitch, int w, int h, int f0, int batch, cudaStream_t stream);
int launch_resize(const LevelDev *h_levels, int level, int f0, int batch, cudaStream_t stream);
// One launch of the fused pyramid kernel (orbx_plan.h: ConePlan): levels src + 1 .. src + nl - 1 from level src.
struct ConeLaunch {
    alignas(64) unsigned char map[128];   // tensor map of the source level plane, box = box_w x box_h
    const ConeLevel *d_tiles;             // [ntiles][nl] on the device
    int src, nl, ntiles, box_w, box_h, pitch, buf0_bytes, buf1_bytes;
    bool ok;
};
int launch_pyramid_cone(const LevelDev *h_levels, const ConeLaunch &cl, int f0, int batch, cudaStream_t stream);
// Tensor maps of the un-blurred level planes for the TMA-staged Gaussian (box 96 x 118); ok = every level has one and is >= 16 x 16.
struct BlurTma {
    alignas(64) unsigned char map[kMaxLevels][128];
    bool level_ok[kMaxLevels];
    bool ok;
};
int launch_blur(const LevelDev *h_levels, const BlurTile *d_tiles, int ntiles, int f0, int batch, cudaStream_t stream, const BlurTma *tma = nullptr);
// Tensor-core Gaussian (orbx_blur_tc.cu): 96 x 122 output tiles, swizzled tensor maps of the un-blurred planes (box 128 x 128), its own tile list
constexpr int kBlurTcTileW = 96, kBlurTcTileH = 122;
struct BlurTc {
    alignas(64) unsigned char map[kMaxLevels][128];    // un-blurred planes, swizzled 128 x 128 box (input tiles)
    alignas(64) unsigned char omap[kMaxLevels][128];   // blurred planes, 96 x 122 box (output tiles)
    bool level_ok[kMaxLevels];
    bool ok;
    const BlurTile *d_tiles;
    int ntiles;
};
int launch_blur_tc(const LevelDev *h_levels, const BlurTc &C, int f0, int batch, cudaStream_t stream, int sm_count);
// Tensor maps of the level planes for the TMA-staged FAST kernel (host side: orbx_api.cu builds them; 128 bytes each,
// stored opaquely so that this header does not need <cuda.h>).
struct FastTma {
    alignas(64) unsigned char map[kMaxLevels][128];
    int box_w[kMaxLevels], box_h[kMaxLevels];
    int max_iw, max_ih;         // largest tested-pixel extent of any cell (sizes the per-warp score tile and list)
    bool level_ok[kMaxLevels];
    bool ok;                    // every level has a valid map
};
int launch_fast(const LevelDev *h_levels, const CellRect *d_cells, int ncells, int f0, int batch,
                int ini_th, int min_th, int *d_overflow, cudaStream_t stream, const FastTma *tma, int sm_count);
// Tensor maps + geometry of the pair-plane FAST kernel (orbx_fast2.cu): one box of box_w[l] x box_h bytes per chunk of `ch` tested
// cell rows; `pitch` = pair-plane pitch in words (one of the kernel's template instantiations).
struct Fast2Tma {
    alignas(64) unsigned char map[kMaxLevels][128];
    int box_w[kMaxLevels];
    int box_h, ch, rows, stage_bytes, pitch;   // ch: chunk height asked for; rows: tallest chunk that occurs (box_h = rows + 6)
    int max_np, max_iw, max_ih;
    bool level_ok[kMaxLevels];
    bool ok;                    // every level has a valid map and a kernel instantiation fits
};
int fast2_pick_pitch(int min_words);   // smallest instantiated pair-plane pitch >= min_words, 0 if none
int launch_fast2(const LevelDev *h_levels, const CellRect *d_cells, int ncells, int f0, int batch, int ini_th, int min_th, int *d_overflow,
                 cudaStream_t stream, const Fast2Tma *tma, int sm_count);
int launch_octree(const LevelDev *h_levels, int nlevels, int f0, int batch, int *d_overflow,
                  cudaStream_t stream);
// d_items [frames][total_out_cap]: per frame the dense list of its keypoints in sequence order, .x = level << 24 | y << 12 | x (level
// coordinates), .y = output slot; 0xFFFFFFFF in .x behind the last one.  The descriptor kernel walks it instead of the level tables.
int launch_finalize(const LevelDev *h_levels, int nlevels, int f0, int batch, int total_out_cap, int lap0, int lap1,
                    KeypointRec *d_kp, int cap, int *d_slot, uint2 *d_items, int *d_n, int *d_mono, int *d_overflow, cudaStream_t stream);
// Tensor maps of the un-blurred and blurred level planes for the TMA-staged descriptor kernel (boxes 64 x 31 and 64 x 39).
struct DescTma {
    alignas(64) unsigned char img[kMaxLevels][128];
    alignas(64) unsigned char blur[kMaxLevels][128];
    bool level_ok[kMaxLevels];
    bool ok;
};
int launch_describe(const LevelDev *h_levels, int nlevels, int f0, int batch, int total_out_cap, const int *d_slot, const uint2 *d_items,
                    KeypointRec *d_kp, uint8_t *d_desc, int cap, cudaStream_t stream, const DescTma *tma, int sm_count);
// Frame post-extraction steps (orbx_frame.cu): cv::undistortPoints of Frame::UndistortKeyPoints and the 64 x 48 feature grid
constexpr int kGridCols = 64, kGridRows = 48;
struct CameraDev { float fx, fy, cx, cy, k1, k2, p1, p2, k3; };   // == orbx_camera
int launch_undistort_xy(const float *d_xy, int n, const CameraDev &cam, float *d_out, cudaStream_t stream);
int launch_frame_grid(const KeypointRec *d_kp, const int *d_n, int n_one, int batch, int cap, const CameraDev &cam, const float *bounds4,
                      KeypointRec *d_kp_un, int32_t *d_cell_start, int32_t *d_cell_items, cudaStream_t stream);
// stand-alone stage launchers for the debug / parity entry points
int launch_describe_points(const uint8_t *d_img, const uint8_t *d_blur, int pitch, const float *d_xy, int n,
                           const float *d_angle_in, float *d_angle_out, uint8_t *d_desc, cudaStream_t stream);
void upload_constants();   // pattern + umax to __constant__/__device__ memory (once per process per device)
}  // namespace orbx

#include <string>
namespace orbx {
// tensor-core kNN (orbx_knn_tc.cu)
size_t knn_tc_smem_bytes();
int knn_tc_max_queries();
long long knn_tc_padded_rows(long long nrows);
int knn_tc_padded_queries(int nq);
int launch_expand_pm1(const uint8_t *d_bits, long long nrows, long long nrows_pad, int8_t *d_out, cudaStream_t stream);
int launch_knn2_tc(const int8_t *d_qe, int nq, const int8_t *d_dbe, long long nrows, long long row_offset, int sm_count,
                   unsigned long long *d_partial, int *nparts_out,
                   void (*merge)(const unsigned long long *, int, int, unsigned long long *, cudaStream_t), cudaStream_t stream,
                   std::string &err);
// FP4 variant (orbx_knn_fp4.cu): descriptors expanded to E2M1 nibbles (128 bytes per row), tcgen05.mma kind::mxf4 with unit block scales
size_t knn_fp4_smem_bytes();
int knn_fp4_max_queries();
long long knn_fp4_padded_rows(long long nrows);
int knn_fp4_padded_queries(int nq);
int launch_expand_fp4(const uint8_t *d_bits, long long nrows, long long nrows_pad, uint8_t *d_out, cudaStream_t stream);
int launch_knn2_fp4(const uint8_t *d_qe, int nq, const uint8_t *d_dbe, long long nrows, long long row_offset, int sm_count,
                    unsigned long long *d_partial, int *nparts_out,
                    void (*merge)(const unsigned long long *, int, int, unsigned long long *, cudaStream_t), cudaStream_t stream,
                    std::string &err);
}  // namespace orbx

namespace orbx {
constexpr int kMaxMatchPairs = 64;   // (query frame, train frame) pairs per launch of the batched grid search (kernel-parameter table)
int match_windowed_grid_batch_device(cudaStream_t stream, int npairs, const int32_t *pair_q, const int32_t *pair_t, int cap, const uint8_t *d_q_desc,
                                     const float *d_q_uvr, const int32_t *d_q_levels, const int32_t *d_n, const KeypointRec *d_t_kp,
                                     const uint8_t *d_t_desc, const int32_t *d_cell_start, const int32_t *d_cell_items, const float *bounds4,
                                     int32_t *d_best_idx, int32_t *d_best_dist, int32_t *d_second_idx, int32_t *d_second_dist);
int match_windowed_grid_device(cudaStream_t stream, const uint8_t *d_q_desc, const float *d_q_uvr, const int32_t *d_q_levels, int nq,
                               const KeypointRec *d_t_kp, const uint8_t *d_t_desc, const int32_t *d_cell_start, const int32_t *d_cell_items,
                               const float *bounds4, int32_t *d_best_idx, int32_t *d_best_dist, int32_t *d_second_idx, int32_t *d_second_dist);
}  // namespace orbx

struct orbx_keypoint;
namespace orbx {
// matching launchers that run on an extractor handle's stream (orbx_match.cu)
int match_distance_batch(int device, cudaStream_t stream, const uint8_t *a, const uint8_t *b, int n, int32_t *dist_out,
                         std::string &err, long long &launches);
int match_windowed(int device, cudaStream_t stream, const uint8_t *q_desc, const float *q_uvr, const int32_t *q_levels, int nq,
                   const orbx_keypoint *t_kp, const uint8_t *t_desc, int nt, const float *bounds4, int32_t *best_idx,
                   int32_t *best_dist, int32_t *second_idx, int32_t *second_dist, std::string &err, long long &launches);

}  // namespace orbx

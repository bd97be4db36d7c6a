// Frame post-extraction steps on the device: what UPSTREAM ORB-SLAM3 src/Frame.cc does right after
// ORBextractor::operator() returns and right before the matchers run --
//   Frame::UndistortKeyPoints   cv::undistortPoints(pts, pts, K, distCoef, cv::Mat(), K): 5 fixed-point iterations of the
//                               radial-tangential model in double precision (camera from the calibration message,
//                               slam_backends/orb_slam_3/orbslam3_mono_networked.cc:173-176)
//   Frame::AssignFeaturesToGrid Frame::PosInGrid on the 64 x 48 grid over the undistorted image bounds; cell lists in
//                               push_back (= keypoint index) order
// The grid is produced as CSR (cell_start[64*48+1], cell_items[n]), which is the index Frame::GetFeaturesInArea walks.
// Arithmetic: double, one rounding per operation (the library is compiled --fmad=false), same operation order as
// OpenCV's cvUndistortPointsInternal, so the CPU checker reproduces it bit for bit.
#include <cuda_runtime.h>

#include <cstdint>

#include "orbx_dev.h"

namespace orbx {

// cvUndistortPointsInternal for one point (criteria = MAX_ITER 5, no tilt, R = I, P = K); k = k1 k2 p1 p2 k3
__device__ __forceinline__ void undistort_point(float u_in, float v_in, const CameraDev &c, float &xo, float &yo) {
    const double fx = c.fx, fy = c.fy, cx = c.cx, cy = c.cy;
    const double k0 = c.k1, k1 = c.k2, k2 = c.p1, k3 = c.p2, k4 = c.k3;
    const double ifx = 1.0 / fx, ify = 1.0 / fy;
    const double u = u_in, v = v_in;
    double x = (u - cx) * ifx, y = (v - cy) * ify;
    const double x0 = x, y0 = y;
#pragma unroll 1
    for (int j = 0; j < 5; j++) {
        const double r2 = x * x + y * y;
        const double icdist = (1 + ((0.0 * r2 + 0.0) * r2 + 0.0) * r2) / (1 + ((k4 * r2 + k1) * r2 + k0) * r2);
        if (icdist < 0) { x = (u - cx) * ifx; y = (v - cy) * ify; break; }
        const double deltaX = 2 * k2 * x * y + k3 * (r2 + 2 * x * x) + 0.0 * r2 + 0.0 * r2 * r2;
        const double deltaY = k2 * (r2 + 2 * y * y) + 2 * k3 * x * y + 0.0 * r2 + 0.0 * r2 * r2;
        x = (x0 - deltaX) * icdist;
        y = (y0 - deltaY) * icdist;
    }
    const double xx = fx * x + 0.0 * y + cx, yy = 0.0 * x + fy * y + cy, ww = 1.0 / (0.0 * x + 0.0 * y + 1.0);
    xo = (float)(xx * ww); yo = (float)(yy * ww);
}

__global__ void __launch_bounds__(128) k_undistort_xy(const float *__restrict__ xy, int n, CameraDev cam, float *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x, y;
    undistort_point(xy[2 * i], xy[2 * i + 1], cam, x, y);
    out[2 * i] = x; out[2 * i + 1] = y;
}

// One CTA per frame: undistort, PosInGrid, counting sort into CSR; a cell's segment is put into keypoint-index order by the
// thread that owns the cell (segments hold a handful of entries).
constexpr int FG_THREADS = 1024, FG_CELLS = kGridCols * kGridRows;
__global__ void __launch_bounds__(FG_THREADS) k_frame_grid(const KeypointRec *__restrict__ kp, const int *__restrict__ n_in, int n_one, int cap,
                                                           CameraDev cam, float minX, float minY, float maxX, float maxY,
                                                           KeypointRec *__restrict__ kp_un, int32_t *__restrict__ cell_start,
                                                           int32_t *__restrict__ cell_items) {
    __shared__ int s_cnt[FG_CELLS + 1];
    __shared__ int s_cur[FG_CELLS];
    __shared__ int s_warp[FG_THREADS / 32];
    const int f = blockIdx.x, tid = threadIdx.x;
    const int n = min(n_in ? n_in[f] : n_one, cap);
    kp += (size_t)f * cap; kp_un += (size_t)f * cap; cell_items += (size_t)f * cap; cell_start += (size_t)f * (FG_CELLS + 1);
    for (int c = tid; c <= FG_CELLS; c += FG_THREADS) s_cnt[c] = 0;
    __syncthreads();
    // Frame::PosInGrid: mfGridElementWidthInv = 64 / (mnMaxX - mnMinX) in float, round() = half away from zero
    const float invW = (float)kGridCols / (maxX - minX), invH = (float)kGridRows / (maxY - minY);
    const bool distorted = cam.k1 != 0.0f;                 // Frame::UndistortKeyPoints: mDistCoef.at<float>(0) == 0.0 -> copy
    for (int i = tid; i < n; i += FG_THREADS) {
        KeypointRec r = kp[i];
        if (distorted) undistort_point(r.x, r.y, cam, r.x, r.y);
        kp_un[i] = r;
        const int px = (int)roundf(__fmul_rn(__fsub_rn(r.x, minX), invW)), py = (int)roundf(__fmul_rn(__fsub_rn(r.y, minY), invH));
        if (px >= 0 && px < kGridCols && py >= 0 && py < kGridRows) atomicAdd(&s_cnt[px * kGridRows + py], 1);
    }
    __syncthreads();
    // exclusive scan of the 3072 counts (3 per thread)
    {
        const int b = tid * 3;
        const int a0 = s_cnt[b], a1 = s_cnt[b + 1], a2 = s_cnt[b + 2];
        const int sum = a0 + a1 + a2;
        int inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if ((tid & 31) >= o) inc += t; }
        if ((tid & 31) == 31) s_warp[tid >> 5] = inc;
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < (tid >> 5); w++) woff += s_warp[w];
        const int ex = woff + inc - sum;
        __syncthreads();
        s_cnt[b] = ex; s_cnt[b + 1] = ex + a0; s_cnt[b + 2] = ex + a0 + a1;
        s_cur[b] = ex; s_cur[b + 1] = ex + a0; s_cur[b + 2] = ex + a0 + a1;
        if (tid == FG_THREADS - 1) s_cnt[FG_CELLS] = ex + sum;
    }
    __syncthreads();
    for (int i = tid; i < n; i += FG_THREADS) {
        const KeypointRec r = kp_un[i];
        const int px = (int)roundf(__fmul_rn(__fsub_rn(r.x, minX), invW)), py = (int)roundf(__fmul_rn(__fsub_rn(r.y, minY), invH));
        if (px >= 0 && px < kGridCols && py >= 0 && py < kGridRows) cell_items[atomicAdd(&s_cur[px * kGridRows + py], 1)] = i;
    }
    __syncthreads();
    for (int c = tid; c < FG_CELLS; c += FG_THREADS) {
        const int lo = s_cnt[c], hi = s_cnt[c + 1];
        for (int a = lo + 1; a < hi; a++) {            // insertion sort: push_back order = ascending keypoint index
            const int v = cell_items[a];
            int b = a - 1;
            while (b >= lo && cell_items[b] > v) { cell_items[b + 1] = cell_items[b]; b--; }
            cell_items[b + 1] = v;
        }
    }
    for (int c = tid; c <= FG_CELLS; c += FG_THREADS) cell_start[c] = s_cnt[c];
}

int launch_undistort_xy(const float *d_xy, int n, const CameraDev &cam, float *d_out, cudaStream_t stream) {
    if (n <= 0) return 0;
    k_undistort_xy<<<(n + 127) / 128, 128, 0, stream>>>(d_xy, n, cam, d_out);
    return 1;
}

int launch_frame_grid(const KeypointRec *d_kp, const int *d_n, int n_one, int batch, int cap, const CameraDev &cam, const float *bounds4,
                      KeypointRec *d_kp_un, int32_t *d_cell_start, int32_t *d_cell_items, cudaStream_t stream) {
    if (batch <= 0) return 0;
    k_frame_grid<<<batch, FG_THREADS, 0, stream>>>(d_kp, d_n, n_one, cap, cam, bounds4[0], bounds4[1], bounds4[2], bounds4[3], d_kp_un,
                                                   d_cell_start, d_cell_items);
    return 1;
}

}  // namespace orbx

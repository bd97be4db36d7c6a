// Host-side geometry plan of the ORB front end: everything that depends only on (extractor parameters, frame size).
// Restates the constructor / ComputePyramid / cell-grid / DistributeOctTree set-up arithmetic of UPSTREAM
// ORB-SLAM3 src/ORBextractor.cc (not under /root/reference; built per slam_backends/orb_slam_3/CMakeLists.txt:52,
// parameters from orbslam3_mono_networked.cc:193-206) as lookup tables the kernels index with integers, so that
// no floating-point geometry is recomputed on the device.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace orbx {

constexpr int kPatchSize = 31;
constexpr int kHalfPatch = 15;
constexpr int kEdge = 19;        // EDGE_THRESHOLD
constexpr int kMinBorder = 16;   // EDGE_THRESHOLD - 3
constexpr int kMaxLevels = 16;
constexpr int kPitchAlign = 32;
constexpr int kMaxBins = 4096;   // quadtree histogram bins per (frame, level) problem
constexpr int kMaxRoots = 16;   // quadtree roots of a level = round(width / height) of its tested region: aspect ratios up to 16.5 : 1

struct ExtractorParams {
    int nfeatures = 1000;
    float scale_factor_f = 1.2f;
    int nlevels = 8;
    int ini_th = 20, min_th = 7;
    // derived (constructor tables)
    float scale[kMaxLevels], inv_scale[kMaxLevels], sigma2[kMaxLevels], inv_sigma2[kMaxLevels];
    int quota[kMaxLevels];
    int umax[kHalfPatch + 2];
};
// Fills the derived tables; returns false on unsupported parameters.
bool init_params(ExtractorParams &p, int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th);

// one bilinear tap pair of cv::resize INTER_LINEAR (11-bit coefficients)
struct ResizeTap {
    int16_t ofs;      // source index of the first tap
    int16_t c0, c1;   // weights of src[ofs], src[min(ofs+1, n-1)]
    int16_t ofs1;     // clamped index of the second tap
};
static_assert(sizeof(ResizeTap) == 8, "ResizeTap must be 8 bytes");

struct CellRect {     // one FAST cell: ROI [x0,x1) x [y0,y1) in level coordinates; tested pixels = ROI shrunk by 3
    int16_t level, x0, y0, x1, y1, pad0, pad1, pad2;
};
static_assert(sizeof(CellRect) == 16, "CellRect must be 16 bytes");

struct LevelPlan {
    int w = 0, h = 0, pitch = 0;
    size_t plane_bytes = 0;          // pitch * h
    // resize taps from level l-1 (empty for level 0)
    std::vector<ResizeTap> xtap, ytap;
    std::vector<uint32_t> xpack;      // compact x taps for the fast kernel; empty if the taps do not fit that form
    // FAST cell grid
    int ncols = 0, nrows = 0, wcell = 0, hcell = 0;
    int first_cell = 0, ncells = 0;  // range in Plan::cells
    int cand_cap = 0;                // worst-case number of FAST candidates of this level
    // quadtree
    int reg_w = 0, reg_h = 0;        // maxBorderX-minBorderX, maxBorderY-minBorderY
    int n_ini = 0;                   // number of root nodes
    int depth0 = 0;                  // histogram depth D0
    int nbins = 0;                   // n_ini * 4^D0
    int root_ulx[kMaxRoots], root_brx[kMaxRoots];
    int quota = 0, out_cap = 0;      // N and max(N+3, 4*n_ini)
    // per relative coordinate (x - 16): low 16 bits bin part (root << 2*D0 | x path bits at even positions),
    // per relative coordinate (y - 16): y path bits at odd positions; ord parts add up to the canonical
    // emission order of the reference's cell loop (cell row-major, then y, then x).
    std::vector<uint32_t> xbin, ybin, xord, yord;
    uint32_t ord_cell_area = 0, ord_ncols = 0;   // to invert ord -> (x, y)
    float kp_size = 0.f;             // (int)(31 * scale)
};

struct Plan {
    int width = 0, height = 0, nlevels = 0;
    std::vector<LevelPlan> lv;
    std::vector<CellRect> cells;     // all levels, level-major
    int total_out_cap = 0;           // sum of out_cap
    size_t algorithmic_bytes(int nkeypoints) const;   // 5S - P0 - P(L-1) + 1321 N
};

void build_resize_taps(int dst, int src, bool is_x, std::vector<ResizeTap> &taps);

// Builds the plan; returns false and sets err on unsupported geometry.
bool build_plan(const ExtractorParams &p, int width, int height, Plan &plan, std::string &err);

}  // namespace orbx

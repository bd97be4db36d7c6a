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

// Fused pyramid ("cone" tiling, k_pyramid_cone): the frame is cut into gx x gy tiles; ONE CTA carries a tile down all levels in shared
// memory.  At level l it computes the region R_l = its own part of the level (what it writes to global memory; the own parts partition
// the level, interior x boundaries on multiples of 4) plus everything the next level's region reads through its bilinear taps, so no CTA
// ever waits for another one.  The source level enters as one TMA box per tile.  Deep cones are wasteful (the halo a level needs grows by
// about two pixels of every level above it), so a pyramid runs as a few launches of up to four levels each.
struct ConeLevel {
    int16_t rx0, ry0, rw, rh;        // region held in shared memory: columns [rx0, rx0 + rw) (both multiples of 4), rows [ry0, ry0 + rh)
    int16_t ox0, ox1, oy0, oy1;      // the part written to global memory; level 0 entry: rx0 / ry0 = origin of the TMA box
};
static_assert(sizeof(ConeLevel) == 16, "ConeLevel must be 16 bytes");
struct ConePlan {
    bool ok = false;
    int src = 0, last = 0;           // this launch reads level `src` (TMA) and produces levels src + 1 .. last; nl = last - src + 1 entries per tile
    int gx = 0, gy = 0, ntiles = 0;
    int box_w = 0, box_h = 0;        // TMA box of level 0: bytes per row (multiple of 16), rows
    int pitch = 0;                   // row pitch in bytes of the level >= 1 regions (multiple of 4, >= widest region + 16)
    int buf0_bytes = 0, buf1_bytes = 0;   // ping-pong buffers: buf0 = TMA box, then the even levels; buf1 = the odd levels
    std::vector<ConeLevel> lv;       // [tile][last - src + 1]: entry k describes level src + k (entry 0: the TMA box)
};

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
    std::vector<ConePlan> cones;     // fused pyramid launches (levels 1..4 from level 0, 5..7 from level 4, ...); empty: per-level kernels
    size_t algorithmic_bytes(int nkeypoints) const;   // 5S - P0 - P(L-1) + 1321 N
};

bool build_cone_plan(const Plan &plan, int src_level, int last_level, ConePlan &cone);
void build_resize_taps(int dst, int src, bool is_x, std::vector<ResizeTap> &taps);

// Builds the plan; returns false and sets err on unsupported geometry.
bool build_plan(const ExtractorParams &p, int width, int height, Plan &plan, std::string &err);

}  // namespace orbx

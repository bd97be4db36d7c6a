// C ABI of the extractor (include/orbx.h): host runtime around the kernels of orbx_kernels.cu.
// One handle = one device, one stream, one workspace sized for (frame size, batch).  No CPU fallback: every entry
// point that computes anything launches the CUDA kernels or fails with ORBX_E_CUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/orbx.h"
#include "orbx_dev.h"
#include "orbx_plan.h"
#include "../../include/orbx_wire.h"
#include "orbx_tma.cuh"

using namespace orbx;

static_assert(sizeof(orbx_keypoint) == 28 && sizeof(KeypointRec) == 28, "keypoint record must match cv::KeyPoint");

static thread_local std::string g_create_error;

// Identity of a host-buffer batch call for CUDA-graph replay: buffers (null = the handle's own staging), geometry, stream.
struct GraphKey {
    const void *in, *kp, *desc, *aux0, *aux1;
    cudaStream_t stream;
    int batch, width, height, stride, lap0, lap1, cap, chunk, fmt, gray_shift;
    size_t frame_stride;   // bytes between device-resident frames (0 for host-buffer calls)
    bool operator==(const GraphKey &o) const {
        return in == o.in && kp == o.kp && desc == o.desc && aux0 == o.aux0 && aux1 == o.aux1 && stream == o.stream && batch == o.batch && width == o.width &&
               height == o.height && stride == o.stride && lap0 == o.lap0 && lap1 == o.lap1 && cap == o.cap && chunk == o.chunk && fmt == o.fmt &&
               gray_shift == o.gray_shift && frame_stride == o.frame_stride;
    }
};
struct GraphEntry { GraphKey key; cudaGraphExec_t exec; long long launches; bool disabled; };

static bool graphs_enabled() {
    static const bool on = [] { const char *e = getenv("ORBX_GRAPHS"); return !(e && e[0] == '0'); }();
    return on;
}

struct orbx_handle {
    orbx_config cfg{};
    ExtractorParams P{};
    int device = 0;
    cudaStream_t stream = nullptr;      // stream all work of this handle is issued on
    cudaStream_t own_stream = nullptr;  // created by orbx_create; `stream` may be redirected by orbx_set_stream
    cudaStream_t side_stream = nullptr; // the Gaussian pass runs here, concurrently with FAST + quadtree
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // software pipeline of the host-buffer batch entry point (created on first use)
    static constexpr int kMaxChunks = 32;
    static constexpr int kComputeStreams = 8;
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr, cs[kComputeStreams] = {};
    int ncs = 4;                        // compute streams in use
    cudaEvent_t ev_start = nullptr, ev_end = nullptr, ev_in[kMaxChunks] = {}, ev_done[kMaxChunks] = {};
    std::vector<GraphEntry> graphs;   // captured call shapes of orbx_extract_batch (dropped on re-plan)
    mutable std::string err;
    long long launches = 0;

    // plan-dependent state
    Plan plan;
    bool planned = false;
    int batch_cap = 0;
    int kp_cap = 0;                 // keypoints per frame the internal result buffers hold
    std::vector<void *> dev_allocs; // everything freed on re-plan / destroy
    LevelDev h_levels[kMaxLevels];
    CellRect *d_cells = nullptr;
    BlurTile *d_tiles = nullptr;
    int ntiles = 0;
    int *d_counts = nullptr;        // [2][nlevels][batch]: cand_count then sel_count
    int *d_overflow = nullptr;
    int *d_slot = nullptr;          // [batch][total_out_cap]
    uint2 *d_items = nullptr;       // [batch][total_out_cap]  dense work list of the descriptor kernel, written by the slot kernel
    KeypointRec *d_kp = nullptr;    // [batch][kp_cap]
    uint8_t *d_desc = nullptr;      // [batch][kp_cap][32]
    int *d_n = nullptr, *d_mono = nullptr;
    FastTma ftma{};                 // tensor maps of the level planes (TMA-staged FAST kernel)
    Fast2Tma ftma2{};               // ... and for the pair-plane FAST kernel (orbx_fast2.cu), the one the pipeline prefers
    DescTma dtma{};                 // ... and of the un-blurred / blurred planes for the descriptor kernel
    BlurTma btma{};                 // ... and of the un-blurred planes for the Gaussian pass
    BlurTc btc{};                   // ... and, swizzled, for its tensor-core variant
    std::vector<ConeLaunch> cones;  // fused pyramid launches (empty or a failed map: the per-level resize kernels run)
    bool cones_ok = false;
    int sm_count = 148;
    // colour input (orbx_set_input_format): frames are uploaded to d_color and converted into the level-0 planes on the device
    int in_fmt = ORBX_FMT_GRAY8, gray_shift = ORBX_GRAY_Q15;
    uint8_t *d_color = nullptr, *h_color = nullptr;
    int color_pitch = 0, color_fmt = 0;
    size_t color_fstride = 0;
    uint8_t *l0_own = nullptr;      // arena copy of level 0 (host-input path)
    size_t l0_own_fstride = 0;
    int l0_own_pitch = 0;
    // pinned host staging
    uint8_t *h_in = nullptr; size_t h_in_bytes = 0;
    KeypointRec *h_kp = nullptr; uint8_t *h_desc = nullptr; int *h_n = nullptr, *h_mono = nullptr, *h_overflow = nullptr;
    int last_batch = 0;             // batch size of the most recent run (debug getters)
    // a host-buffer batch queued by orbx_extract_batch_submit and not yet collected
    struct Pending { bool active; int batch, cap; bool out_direct; orbx_keypoint *kp_out; uint8_t *desc_out; };
    Pending pending{false, 0, 0, false, nullptr, nullptr};
    // optional per-stage timing (orbx_set_profiling): events around resize / blur / fast / octree / finalize / describe
    bool profiling = false;
    cudaEvent_t ev[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool ev_valid = false;
};

#define CU_TRY(h, expr)                                                                                         \
    do {                                                                                                        \
        cudaError_t e__ = (expr);                                                                               \
        if (e__ != cudaSuccess) {                                                                               \
            (h)->err = std::string(#expr) + ": " + cudaGetErrorString(e__);                                     \
            return ORBX_E_CUDA;                                                                                 \
        }                                                                                                       \
    } while (0)

// true for cudaHostAlloc'd / cudaHostRegister'ed memory (DMA-able without staging)
static bool is_pinned_host(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

static int fail(const orbx_handle *h, int code, const char *msg) {
    if (h) h->err = msg; else g_create_error = msg;
    return code;
}

static void free_plan(orbx_handle *h) {
    for (auto &g : h->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    h->graphs.clear();
    for (void *p : h->dev_allocs) cudaFree(p);
    h->dev_allocs.clear();
    if (h->h_in) cudaFreeHost(h->h_in);
    if (h->h_kp) cudaFreeHost(h->h_kp);
    if (h->h_desc) cudaFreeHost(h->h_desc);
    if (h->h_n) cudaFreeHost(h->h_n);
    if (h->h_color) cudaFreeHost(h->h_color);
    h->h_color = nullptr; h->d_color = nullptr; h->color_fmt = 0;
    h->h_in = nullptr; h->h_kp = nullptr; h->h_desc = nullptr; h->h_n = nullptr; h->h_mono = nullptr; h->h_overflow = nullptr;
    h->planned = false;
}

template <typename T>
static cudaError_t dev_alloc(orbx_handle *h, T **out, size_t count) {
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(count * sizeof(T), 256));
    if (e == cudaSuccess) { h->dev_allocs.push_back(p); *out = (T *)p; }
    return e;
}

template <typename T>
static cudaError_t dev_upload(orbx_handle *h, T **out, const std::vector<T> &v) {
    cudaError_t e = dev_alloc(h, out, v.size());
    if (e != cudaSuccess || v.empty()) return e;
    return cudaMemcpy(*out, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
}

static_assert(sizeof(CUtensorMap) == 128, "FastTma stores tensor maps as 128-byte blobs");

// Gaussian pass: the tensor-core kernel when every plane has its swizzled map, else (or when it cannot be configured) the CUDA-core ones
static int launch_blur_any(orbx_handle *h, int f0, int batch, cudaStream_t stream) {
    int n = h->btc.ok ? launch_blur_tc(h->h_levels, h->btc, f0, batch, stream, h->sm_count) : 0;
    if (!n) n = launch_blur(h->h_levels, h->d_tiles, h->ntiles, f0, batch, stream, &h->btma);
    return n;
}

// (Re)encode the tensor map of one level plane; marks the level unusable when the plane misses the TMA alignment rules.
static void encode_fast_map(orbx_handle *h, int l) {
    const LevelDev &D = h->h_levels[l];
    FastTma &T = h->ftma;
    T.level_ok[l] = T.box_w[l] > 0 && tma_make_plane_map(reinterpret_cast<CUtensorMap *>(T.map[l]), D.img, D.w, D.h, h->batch_cap, (size_t)D.pitch,
                                                       D.img_fstride, T.box_w[l], T.box_h[l]);
    T.ok = !getenv("ORBX_NO_TMA");
    for (int k = 0; k < h->plan.nlevels; k++) T.ok = T.ok && (T.level_ok[k] || h->plan.lv[k].ncells == 0);
    Fast2Tma &T2 = h->ftma2;
    T2.level_ok[l] = T2.pitch > 0 && T2.box_w[l] > 0 &&
                     tma_make_plane_map(reinterpret_cast<CUtensorMap *>(T2.map[l]), D.img, D.w, D.h, h->batch_cap, (size_t)D.pitch, D.img_fstride,
                                        T2.box_w[l], T2.box_h);
    {
        static const int fast_v = [] { const char *e = getenv("ORBX_FAST_V"); return e ? atoi(e) : 1; }();
        T2.ok = !getenv("ORBX_NO_TMA") && fast_v == 2 && T2.pitch > 0;
    }
    for (int k = 0; k < h->plan.nlevels; k++) T2.ok = T2.ok && (T2.level_ok[k] || h->plan.lv[k].ncells == 0);
    // descriptor kernel: 64 x 31 box on the un-blurred plane, 64 x 39 box on the blurred plane
    DescTma &Q = h->dtma;
    Q.level_ok[l] = tma_make_plane_map(reinterpret_cast<CUtensorMap *>(Q.img[l]), D.img, D.w, D.h, h->batch_cap, (size_t)D.pitch, D.img_fstride, 64, 31) &&
                    tma_make_plane_map(reinterpret_cast<CUtensorMap *>(Q.blur[l]), D.blur, D.w, D.h, h->batch_cap, (size_t)D.blur_pitch, D.blur_fstride, 64, 39);
    Q.ok = !getenv("ORBX_NO_TMA");
    for (int k = 0; k < h->plan.nlevels; k++) Q.ok = Q.ok && Q.level_ok[k];
    // fused pyramid: the launch whose source is this level
    for (ConeLaunch &cl : h->cones)
        if (cl.src == l)
            cl.ok = tma_make_plane_map(reinterpret_cast<CUtensorMap *>(cl.map), D.img, D.w, D.h, h->batch_cap, (size_t)D.pitch, D.img_fstride, cl.box_w, cl.box_h);
    h->cones_ok = !h->cones.empty() && !getenv("ORBX_NO_TMA") && !getenv("ORBX_NO_CONE");
    for (const ConeLaunch &cl : h->cones) h->cones_ok = h->cones_ok && cl.ok;
    // Gaussian pass: 96 x 118 box on the un-blurred plane; planes under 16 x 16 keep the word-load kernel (multiple reflections)
    BlurTma &G = h->btma;
    G.level_ok[l] = D.w >= 16 && D.h >= 16 &&
                    tma_make_plane_map(reinterpret_cast<CUtensorMap *>(G.map[l]), D.img, D.w, D.h, h->batch_cap, (size_t)D.pitch, D.img_fstride, 96, 118);
    G.ok = !getenv("ORBX_NO_TMA") && !getenv("ORBX_BLUR_WORDS");
    for (int k = 0; k < h->plan.nlevels; k++) G.ok = G.ok && G.level_ok[k];
    // tensor-core Gaussian: 128 x 128 swizzled box; the blurred planes are the library's own (pitch a multiple of 32)
    BlurTc &C = h->btc;
    C.level_ok[l] = D.w >= 16 && D.h >= 16 && (D.blur_pitch & 31) == 0 &&
                    tma_make_plane_map_sw128(reinterpret_cast<CUtensorMap *>(C.map[l]), D.img, D.w, D.h, h->batch_cap, (size_t)D.pitch, D.img_fstride, 128) &&
                    tma_make_plane_map(reinterpret_cast<CUtensorMap *>(C.omap[l]), D.blur, D.w, D.h, h->batch_cap, (size_t)D.blur_pitch, D.blur_fstride,
                                       kBlurTcTileW, kBlurTcTileH);
    {
        static const int blur_tc = [] { const char *e = getenv("ORBX_BLUR_TC"); return e ? atoi(e) : 1; }();   // default on; 0 = k_blur_tma
        C.ok = blur_tc != 0 && !getenv("ORBX_NO_TMA") && C.ntiles > 0;
    }
    for (int k = 0; k < h->plan.nlevels; k++) C.ok = C.ok && C.level_ok[k];
}

// Box of a level = the largest ROI of its cells plus the 4-pixel-group / row-pair overhang the scoring items read.
static void build_fast_maps(orbx_handle *h) {
    FastTma &T = h->ftma;
    std::memset(&T, 0, sizeof(T));
    std::memset(&h->dtma, 0, sizeof(h->dtma));
    std::memset(&h->btma, 0, sizeof(h->btma));
    { const BlurTile *keep_t = h->btc.d_tiles; const int keep_n = h->btc.ntiles; std::memset(&h->btc, 0, sizeof(h->btc)); h->btc.d_tiles = keep_t; h->btc.ntiles = keep_n; }
    for (int l = 0; l < h->plan.nlevels; l++) {
        const LevelPlan &LP = h->plan.lv[l];
        int rw = 0, rh = 0;
        for (int c = LP.first_cell; c < LP.first_cell + LP.ncells; c++) {
            rw = std::max(rw, h->plan.cells[c].x1 - h->plan.cells[c].x0);
            rh = std::max(rh, h->plan.cells[c].y1 - h->plan.cells[c].y0);
        }
        if (LP.ncells == 0) continue;
        int bw = (15 + rw + 5 + 15) / 16 * 16;   // the box starts at x0 rounded down to 16 bytes (TMA rule)
        T.box_w[l] = bw; T.box_h[l] = rh + 1;
        T.max_iw = std::max(T.max_iw, rw - 6); T.max_ih = std::max(T.max_ih, rh - 6);
        if (bw > 96 || rh + 1 > 80) { T.box_w[l] = 0; }   // larger than the shared-memory stage: no TMA for this plan
    }
    // pair-plane kernel: per level the bytes a row of the widest cell needs behind its 16-byte aligned box start (up to 12 bytes of
    // lead + 4 (K + 1) bytes, K = byte words unpacked per row); one chunk height for the whole pyramid
    Fast2Tma &T2 = h->ftma2;
    std::memset(&T2, 0, sizeof(T2));
    int kmax = 0;
    for (int l = 0; l < h->plan.nlevels; l++) {
        const LevelPlan &LP = h->plan.lv[l];
        if (LP.ncells == 0) continue;
        int iw = 0, ih = 0;
        for (int c = LP.first_cell; c < LP.first_cell + LP.ncells; c++) {
            iw = std::max(iw, h->plan.cells[c].x1 - h->plan.cells[c].x0 - 6);
            ih = std::max(ih, h->plan.cells[c].y1 - h->plan.cells[c].y0 - 6);
        }
        const int np = (iw + 1) / 2, K = (2 * np + 8 + 3 + 7) / 8;      // 8-pixel groups unpacked per row
        T2.box_w[l] = (12 + 4 * (2 * K + 1) + 15) / 16 * 16;
        T2.max_np = std::max(T2.max_np, np); T2.max_iw = std::max(T2.max_iw, iw); T2.max_ih = std::max(T2.max_ih, ih);
        kmax = std::max(kmax, K);
    }
    if (kmax > 0) {
        static const int ch_env = [] { const char *e = getenv("ORBX_FAST_CH"); return e ? atoi(e) : 0; }();
        const int ch_want = ch_env >= 4 ? ch_env : 24;
        // balanced chunks: a cell of ih rows runs as nch = ceil(ih / ch) chunks of ceil(ih / nch) rows; the per-warp buffers are sized
        // for the tallest chunk that actually occurs
        T2.ch = std::min(ch_want, T2.max_ih);
        int crmax = 1;
        for (const CellRect &c : h->plan.cells) {
            const int ih = c.y1 - c.y0 - 6, nch = (ih + T2.ch - 1) / T2.ch;
            crmax = std::max(crmax, (ih + nch - 1) / nch);
        }
        T2.rows = crmax;
        T2.box_h = crmax + 6;
        T2.pitch = fast2_pick_pitch(8 * kmax + 1);
        int bw = 0;
        for (int l = 0; l < h->plan.nlevels; l++) bw = std::max(bw, T2.box_w[l]);
        T2.stage_bytes = (bw * T2.box_h + 8 + 127) / 128 * 128;
        if (bw > 256 || T2.box_h > 256) T2.pitch = 0;
    }
    for (int l = 0; l < h->plan.nlevels; l++) encode_fast_map(h, l);
}

// (Re)build the geometry plan and the workspace for w x h frames, `batch` frames per call.
static int ensure_plan(orbx_handle *h, int w, int ht, int batch) {
    if (w > h->cfg.max_width || ht > h->cfg.max_height) return fail(h, ORBX_E_CAPACITY, "frame larger than max_width x max_height");
    if (batch > h->cfg.max_batch) return fail(h, ORBX_E_CAPACITY, "batch larger than max_batch");
    if (h->planned && h->plan.width == w && h->plan.height == ht) return ORBX_OK;
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    free_plan(h);
    std::string err;
    if (!build_plan(h->P, w, ht, h->plan, err)) { h->err = err; return ORBX_E_INVALID; }
    const Plan &pl = h->plan;
    const int B = h->cfg.max_batch, nl = pl.nlevels;
    h->batch_cap = B;
    h->kp_cap = pl.total_out_cap;
    CU_TRY(h, dev_upload(h, &h->d_cells, pl.cells));
    CU_TRY(h, dev_alloc(h, &h->d_counts, (size_t)2 * nl * B));
    CU_TRY(h, dev_alloc(h, &h->d_slot, (size_t)B * pl.total_out_cap));
    CU_TRY(h, dev_alloc(h, &h->d_items, (size_t)B * pl.total_out_cap));
    CU_TRY(h, dev_alloc(h, &h->d_kp, (size_t)B * h->kp_cap));
    CU_TRY(h, dev_alloc(h, &h->d_desc, (size_t)B * h->kp_cap * 32));
    // counts, mono indices and the overflow flag share one block (same layout as the pinned h_n block): one D2H copy per call
    CU_TRY(h, dev_alloc(h, &h->d_n, (size_t)2 * B + 1));
    h->d_mono = h->d_n + B; h->d_overflow = h->d_n + 2 * B;
    std::vector<BlurTile> tiles;
    int out_base = 0;
    std::memset(h->h_levels, 0, sizeof(h->h_levels));
    for (int l = 0; l < nl; l++) {
        const LevelPlan &LP = pl.lv[l];
        LevelDev &D = h->h_levels[l];
        D.w = LP.w; D.h = LP.h; D.pitch = LP.pitch; D.blur_pitch = LP.pitch; D.padded = 1;
        D.img_fstride = LP.plane_bytes; D.blur_fstride = LP.plane_bytes;
        CU_TRY(h, dev_alloc(h, &D.img, LP.plane_bytes * B + 64));
        CU_TRY(h, dev_alloc(h, &D.blur, LP.plane_bytes * B + 64));
        if (l == 0) { h->l0_own = D.img; h->l0_own_fstride = LP.plane_bytes; h->l0_own_pitch = LP.pitch; }
        if (l > 0) {
            ResizeTap *xt = nullptr, *yt = nullptr;
            CU_TRY(h, dev_upload(h, &xt, LP.xtap));
            CU_TRY(h, dev_upload(h, &yt, LP.ytap));
            D.xtap = xt; D.ytap = yt;
            if (!LP.xpack.empty()) { uint32_t *xp = nullptr; CU_TRY(h, dev_upload(h, &xp, LP.xpack)); D.xpack = xp; }
        }
        uint32_t *t0 = nullptr, *t1 = nullptr, *t2 = nullptr, *t3 = nullptr;
        CU_TRY(h, dev_upload(h, &t0, LP.xbin)); CU_TRY(h, dev_upload(h, &t1, LP.ybin));
        CU_TRY(h, dev_upload(h, &t2, LP.xord)); CU_TRY(h, dev_upload(h, &t3, LP.yord));
        D.xbin = t0; D.ybin = t1; D.xord = t2; D.yord = t3;
        D.reg_w = LP.reg_w; D.reg_h = LP.reg_h; D.n_ini = LP.n_ini; D.depth0 = LP.depth0; D.nbins = LP.nbins;
        D.quota = LP.quota; D.out_cap = LP.out_cap;
        for (int r = 0; r < kMaxRoots; r++) { D.root_ulx[r] = LP.root_ulx[r]; D.root_brx[r] = LP.root_brx[r]; }
        D.ord_cell_area = LP.ord_cell_area; D.ord_ncols = LP.ord_ncols; D.wcell = LP.wcell; D.hcell = LP.hcell;
        D.cand_cap = LP.cand_cap;
        CU_TRY(h, dev_alloc(h, &D.cand, (size_t)B * LP.cand_cap));
        CU_TRY(h, dev_alloc(h, &D.sorted, (size_t)B * LP.cand_cap));
        CU_TRY(h, dev_alloc(h, &D.bin_cursor, (size_t)B * std::max(LP.nbins, 1)));
        D.cand_count = h->d_counts + (size_t)l * B;
        D.sel_count = h->d_counts + (size_t)(nl + l) * B;
        CU_TRY(h, dev_alloc(h, &D.sel, (size_t)B * LP.out_cap));
        D.out_base = out_base; out_base += LP.out_cap;
        D.scale = h->P.scale[l]; D.kp_size = LP.kp_size;
        for (int ty = 0; ty * kBlurTileH < LP.h; ty++)
            for (int tx = 0; tx * kBlurTileW < LP.w; tx++) tiles.push_back(BlurTile{(int16_t)l, (int16_t)tx, (int16_t)ty, 0});
    }
    h->ntiles = (int)tiles.size();
    CU_TRY(h, dev_upload(h, &h->d_tiles, tiles));
    {
        std::vector<BlurTile> tc_tiles;
        for (int l = 0; l < nl; l++)
            for (int ty = 0; ty * kBlurTcTileH < pl.lv[l].h; ty++)
                for (int tx = 0; tx * kBlurTcTileW < pl.lv[l].w; tx++) tc_tiles.push_back(BlurTile{(int16_t)l, (int16_t)tx, (int16_t)ty, 0});
        BlurTile *d_tc = nullptr;
        CU_TRY(h, dev_upload(h, &d_tc, tc_tiles));
        h->btc.d_tiles = d_tc; h->btc.ntiles = (int)tc_tiles.size();
    }
    h->cones.clear();
    for (const ConePlan &cp : pl.cones) {
        ConeLaunch cl;
        std::memset(&cl, 0, sizeof(cl));
        ConeLevel *d = nullptr;
        CU_TRY(h, dev_upload(h, &d, cp.lv));
        cl.d_tiles = d; cl.src = cp.src; cl.nl = cp.last - cp.src + 1; cl.ntiles = cp.ntiles; cl.box_w = cp.box_w; cl.box_h = cp.box_h;
        cl.pitch = cp.pitch; cl.buf0_bytes = cp.buf0_bytes; cl.buf1_bytes = cp.buf1_bytes; cl.ok = false;
        h->cones.push_back(cl);
    }
    build_fast_maps(h);
    // pinned staging
    h->h_in_bytes = (size_t)B * pl.lv[0].plane_bytes;
    CU_TRY(h, cudaMallocHost((void **)&h->h_in, h->h_in_bytes));
    CU_TRY(h, cudaMallocHost((void **)&h->h_kp, (size_t)B * h->kp_cap * sizeof(KeypointRec)));
    CU_TRY(h, cudaMallocHost((void **)&h->h_desc, (size_t)B * h->kp_cap * 32));
    CU_TRY(h, cudaMallocHost((void **)&h->h_n, (size_t)(2 * B + 1) * sizeof(int)));
    h->h_mono = h->h_n + B; h->h_overflow = h->h_n + 2 * B;
    CU_TRY(h, cudaStreamSynchronize(cudaStreamLegacy));   // table uploads above are blocking copies on the legacy stream
    h->planned = true;
    return ORBX_OK;
}

// Point level 0 at `img` (device memory) and refresh the device copy of the level table if it moved.
static int set_level0(orbx_handle *h, const uint8_t *img, int pitch, size_t fstride) {
    LevelDev &D = h->h_levels[0];
    if (D.img == img && D.pitch == pitch && D.img_fstride == fstride) return ORBX_OK;
    D.img = const_cast<uint8_t *>(img); D.pitch = pitch; D.img_fstride = fstride;
    D.padded = img == h->l0_own ? 1 : 0;
    encode_fast_map(h, 0);
    return ORBX_OK;
}

// Frames per range of the host-buffer software pipeline (>= batch disables pipelining; ORBX_CHUNK overrides).  Measured
// on 64 x 640x480 (ms per call): one range 1.19, 2 x 32 0.86, 4 x 16 0.90, 8 x 8 1.15, geometric 8/16/40 1.07 - small
// ranges leave the latency-bound kernels (pyramid chain, quadtree) exposed, so ranges stay at 32 frames.
static int pipeline_chunk(int batch) {
    static const int env = [] { const char *e = getenv("ORBX_CHUNK"); return e ? atoi(e) : 0; }();
    int c = env > 0 ? env : (batch >= 64 ? 32 : 16);
    if (batch < 2 * c) return batch;
    while ((batch + c - 1) / c > orbx_handle::kMaxChunks) c *= 2;
    return c;
}

static int ensure_pipeline(orbx_handle *h) {
    if (h->h2d_stream) return ORBX_OK;
    CU_TRY(h, cudaStreamCreateWithFlags(&h->h2d_stream, cudaStreamNonBlocking));
    CU_TRY(h, cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking));
    { const char *e = getenv("ORBX_STREAMS"); if (e && atoi(e) >= 1 && atoi(e) <= orbx_handle::kComputeStreams) h->ncs = atoi(e); }
    for (int i = 0; i < orbx_handle::kComputeStreams; i++) CU_TRY(h, cudaStreamCreateWithFlags(&h->cs[i], cudaStreamNonBlocking));
    CU_TRY(h, cudaEventCreateWithFlags(&h->ev_start, cudaEventDisableTiming));
    CU_TRY(h, cudaEventCreateWithFlags(&h->ev_end, cudaEventDisableTiming));
    for (int i = 0; i < orbx_handle::kMaxChunks; i++) {
        CU_TRY(h, cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
        CU_TRY(h, cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
    }
    return ORBX_OK;
}

// Device + pinned staging for colour frames of the current plan (3 or 4 bytes per pixel).
static int ensure_color(orbx_handle *h) {
    const int bpp = h->in_fmt >= ORBX_FMT_RGBA8 ? 4 : 3;
    if (h->d_color && h->color_fmt == bpp) return ORBX_OK;
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    for (auto &g : h->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    h->graphs.clear();
    if (h->h_color) { cudaFreeHost(h->h_color); h->h_color = nullptr; }
    if (h->d_color) {   // a 3 <-> 4 bytes-per-pixel switch: the old staging goes now, not at the next re-plan
        h->dev_allocs.erase(std::remove(h->dev_allocs.begin(), h->dev_allocs.end(), (void *)h->d_color), h->dev_allocs.end());
        cudaFree(h->d_color);
        h->d_color = nullptr;
    }
    h->color_pitch = (h->plan.width * bpp + 15) / 16 * 16;
    h->color_fstride = (size_t)h->color_pitch * h->plan.height;
    CU_TRY(h, dev_alloc(h, &h->d_color, h->color_fstride * h->batch_cap + 64));   // freed with the plan
    CU_TRY(h, cudaMallocHost((void **)&h->h_color, h->color_fstride * h->batch_cap));
    h->color_fmt = bpp;
    return ORBX_OK;
}

// CUDA-graph replay of a call shape: the first sighting of a key runs `enqueue` directly (which also performs every lazy
// one-time initialisation outside a capture), the second captures it into a graph, later ones replay the graph.
// enqueue(direct): direct = true when the work is issued for real, false while capturing.  before_launch runs before a replay.
template <typename Enqueue, typename Before>
static int run_graphed(orbx_handle *h, const GraphKey &key, Enqueue enqueue, Before before_launch) {
    GraphEntry *e = nullptr;
    for (auto &g : h->graphs) if (g.key == key) { e = &g; break; }
    if (!e) {
        if (h->graphs.size() >= 32) { if (h->graphs.front().exec) cudaGraphExecDestroy(h->graphs.front().exec); h->graphs.erase(h->graphs.begin()); }
        h->graphs.push_back(GraphEntry{key, nullptr, 0, h->stream == cudaStreamLegacy});   // the legacy stream cannot be captured
        return enqueue(true);
    }
    if (e->disabled) return enqueue(true);
    if (!e->exec) {
        // Any failure to capture or instantiate falls back to direct issue for this key (never an error for the caller).
        const long long l0 = h->launches;
        if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError();
            e->disabled = true;
            return enqueue(true);
        }
        const int rc = enqueue(false);
        cudaGraph_t graph = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
        e->launches = h->launches - l0;
        h->launches = l0;
        if (rc != ORBX_OK || ce != cudaSuccess || !graph || cudaGraphInstantiate(&e->exec, graph, 0) != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            e->exec = nullptr; e->disabled = true;
            return enqueue(true);
        }
        cudaGraphDestroy(graph);
    }
    before_launch();
    CU_TRY(h, cudaGraphLaunch(e->exec, h->stream));
    h->launches += e->launches;
    return ORBX_OK;
}

// Clears the per-frame counters of the whole workspace (must precede the kernels of every frame range of a call).
static int reset_counters(orbx_handle *h, cudaStream_t stream) {
    CU_TRY(h, cudaMemsetAsync(h->d_counts, 0, sizeof(int) * 2 * h->plan.nlevels * h->batch_cap, stream));
    CU_TRY(h, cudaMemsetAsync(h->d_overflow, 0, sizeof(int), stream));
    return ORBX_OK;
}

// The kernel pipeline for frames [f0, f0 + batch) of the workspace, whose level 0 is already in place.  Everything is
// issued on `stream`; with a `side` stream the Gaussian pass is forked onto it (ev_fork / ev_join order it).
struct ColorSrc { const uint8_t *ptr; int pitch; size_t fstride; };   // device memory holding colour frames; ptr == nullptr: gray input

static int run_pipeline(orbx_handle *h, int f0, int batch, int lap0, int lap1, KeypointRec *d_kp, uint8_t *d_desc, int cap,
                        int *d_n, int *d_mono, cudaStream_t stream, cudaStream_t side, cudaEvent_t ev_fork, cudaEvent_t ev_join,
                        ColorSrc color = ColorSrc{nullptr, 0, 0}) {
    const Plan &pl = h->plan;
    const int nl = pl.nlevels;
    const bool prof = h->profiling && stream == h->stream;
    const bool fork = !prof && side != nullptr;
    static const bool dbg_sync = getenv("ORBX_DEBUG_SYNC") != nullptr;   // localise a faulting kernel: sync after every stage
    // what-if timing only (results are wrong): bit 0 pyramid, 1 blur, 2 FAST, 3 quadtree + slots, 4 descriptors are not launched
    static const int skip = [] { const char *e = getenv("ORBX_SKIP_STAGES"); return e ? atoi(e) : 0; }();
#define STAGE_MARK(i)                                                                                             \
    do {                                                                                                          \
        if (prof) CU_TRY(h, cudaEventRecord(h->ev[i], stream));                                                   \
        if (dbg_sync) {                                                                                           \
            cudaError_t e__ = cudaStreamSynchronize(stream);                                                      \
            if (e__ != cudaSuccess) { h->err = std::string("stage ") + #i + ": " + cudaGetErrorString(e__); return ORBX_E_CUDA; } \
        }                                                                                                         \
    } while (0)
    if (color.ptr)   // cv::cvtColor of Tracking::GrabImageMonocular, on the device, into the level-0 planes
        h->launches += launch_gray(color.ptr, color.fstride, color.pitch, h->in_fmt, h->gray_shift, h->l0_own, h->l0_own_fstride, h->l0_own_pitch,
                                   pl.width, pl.height, f0, batch, stream);
    STAGE_MARK(0);
    if (!(skip & 1)) {
        if (h->cones_ok) for (const ConeLaunch &cl : h->cones) h->launches += launch_pyramid_cone(h->h_levels, cl, f0, batch, stream);
        else for (int l = 1; l < nl; l++) h->launches += launch_resize(h->h_levels, l, f0, batch, stream);
    }
    STAGE_MARK(1);
    // The blurred planes are only needed by the descriptor stage, and the quadtree kernel (one CTA per frame x level,
    // latency-bound) cannot fill the machine: outside profiling mode quadtree + slot assignment run on the side stream,
    // which has the highest priority so that its few CTAs are placed first, while the Gaussian pass fills the rest of
    // the machine from the main stream.  The fork comes after FAST because two machine-filling kernels gain nothing
    // from running side by side.
    if (!fork && !(skip & 2)) h->launches += launch_blur_any(h, f0, batch, stream);
    STAGE_MARK(2);
    if (!(skip & 4)) {
        int nl2 = 0;
        if (h->ftma2.ok) nl2 = launch_fast2(h->h_levels, h->d_cells, (int)pl.cells.size(), f0, batch, h->P.ini_th, h->P.min_th, h->d_overflow, stream, &h->ftma2, h->sm_count);
        if (!nl2) nl2 = launch_fast(h->h_levels, h->d_cells, (int)pl.cells.size(), f0, batch, h->P.ini_th, h->P.min_th, h->d_overflow, stream, &h->ftma, h->sm_count);
        h->launches += nl2;
    }
    STAGE_MARK(3);
    cudaStream_t qs = fork ? side : stream;
    if (fork) {
        CU_TRY(h, cudaEventRecord(ev_fork, stream));
        CU_TRY(h, cudaStreamWaitEvent(side, ev_fork, 0));
    }
    if (!(skip & 8)) h->launches += launch_octree(h->h_levels, nl, f0, batch, h->d_overflow, qs);
    STAGE_MARK(4);
    if (!(skip & 8)) h->launches += launch_finalize(h->h_levels, nl, f0, batch, pl.total_out_cap, lap0, lap1, d_kp, cap, h->d_slot, h->d_items, d_n, d_mono, h->d_overflow, qs);
    STAGE_MARK(5);
    if (fork) {
        CU_TRY(h, cudaEventRecord(ev_join, side));
        if (!(skip & 2)) h->launches += launch_blur_any(h, f0, batch, stream);
        CU_TRY(h, cudaStreamWaitEvent(stream, ev_join, 0));
    }
    if (!(skip & 16)) h->launches += launch_describe(h->h_levels, nl, f0, batch, pl.total_out_cap, h->d_slot, h->d_items, d_kp, d_desc, cap, stream, &h->dtma, h->sm_count);
    STAGE_MARK(6);
#undef STAGE_MARK
    if (stream == h->stream) h->ev_valid = prof;
    CU_TRY(h, cudaGetLastError());
    return ORBX_OK;
}

// scratch device buffer helper for the stand-alone stage entry points
struct Scratch {
    std::vector<void *> p;
    ~Scratch() { for (void *q : p) cudaFree(q); }
    template <typename T> T *get(size_t n) { void *q = nullptr; if (cudaMalloc(&q, std::max<size_t>(n * sizeof(T), 256)) != cudaSuccess) return nullptr; p.push_back(q); return (T *)q; }
};

extern "C" {

const char *orbx_version(void) { return "orbx 0.2 sm_100a"; }

void *orbx_host_alloc(size_t bytes, int write_combined) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void orbx_host_free(void *p) { if (p) cudaFreeHost(p); }

int orbx_create(const orbx_config *cfg, orbx_handle **out) {
    if (!cfg || !out) return fail(nullptr, ORBX_E_INVALID, "null argument");
    *out = nullptr;
    if (cfg->max_width < 1 || cfg->max_height < 1 || cfg->max_width > 4095 || cfg->max_height > 4095 || cfg->max_batch < 1 || cfg->max_batch > 4096)
        return fail(nullptr, ORBX_E_INVALID, "max_width/max_height must be 1..4095, max_batch 1..4096");
    orbx_handle *h = new (std::nothrow) orbx_handle();
    if (!h) return fail(nullptr, ORBX_E_INVALID, "out of host memory");
    h->cfg = *cfg;
    if (!init_params(h->P, cfg->nfeatures, cfg->scale_factor, cfg->nlevels, cfg->ini_th_fast, cfg->min_th_fast)) {
        delete h;
        return fail(nullptr, ORBX_E_INVALID, "unsupported extractor parameters (nlevels 1..16, 1 < scaleFactor < 2, 1 <= minTh <= iniTh <= 254)");
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev < 1) {
        g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (orbx has no CPU fallback)";
        delete h; return ORBX_E_CUDA;
    }
    if (cfg->device < 0 || cfg->device >= ndev) { delete h; return fail(nullptr, ORBX_E_INVALID, "device ordinal out of range"); }
    h->device = cfg->device;
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, h->device);
    if (h->sm_count < 1) h->sm_count = 148;
    if ((e = cudaSetDevice(h->device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        g_create_error = std::string("cuda init: ") + cudaGetErrorString(e);
        delete h; return ORBX_E_CUDA;
    }
    h->own_stream = h->stream;
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if ((e = cudaStreamCreateWithPriority(&h->side_stream, cudaStreamNonBlocking, prio_hi)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming)) != cudaSuccess) {
        g_create_error = std::string("cuda init: ") + cudaGetErrorString(e);
        orbx_destroy(h); return ORBX_E_CUDA;
    }
    upload_constants();
    if ((e = cudaGetLastError()) != cudaSuccess) {
        g_create_error = std::string("constant upload: ") + cudaGetErrorString(e);
        orbx_destroy(h); return ORBX_E_CUDA;
    }
    *out = h;
    return ORBX_OK;
}

void orbx_destroy(orbx_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    free_plan(h);
    for (int i = 0; i < 7; i++) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->side_stream) cudaStreamDestroy(h->side_stream);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->h2d_stream) cudaStreamDestroy(h->h2d_stream);
    if (h->d2h_stream) cudaStreamDestroy(h->d2h_stream);
    for (int i = 0; i < orbx_handle::kComputeStreams; i++) if (h->cs[i]) cudaStreamDestroy(h->cs[i]);
    if (h->ev_start) cudaEventDestroy(h->ev_start);
    if (h->ev_end) cudaEventDestroy(h->ev_end);
    for (int i = 0; i < orbx_handle::kMaxChunks; i++) {
        if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]);
        if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]);
    }
    delete h;
}

const char *orbx_last_error(const orbx_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int orbx_keypoint_capacity(const orbx_handle *h) {
    if (!h) return ORBX_E_INVALID;
    // sum over levels of max(quota + 3, 4 * roots); roots <= 8
    int cap = 0;
    for (int l = 0; l < h->P.nlevels; l++) cap += std::max(h->P.quota[l] + 3, 4 * kMaxRoots);
    return cap;
}

int orbx_get_tables(const orbx_handle *h, float *scale, float *inv_scale, float *sigma2, float *inv_sigma2, int *features_per_level) {
    if (!h) return ORBX_E_INVALID;
    for (int l = 0; l < h->P.nlevels; l++) {
        if (scale) scale[l] = h->P.scale[l];
        if (inv_scale) inv_scale[l] = h->P.inv_scale[l];
        if (sigma2) sigma2[l] = h->P.sigma2[l];
        if (inv_sigma2) inv_sigma2[l] = h->P.inv_sigma2[l];
        if (features_per_level) features_per_level[l] = h->P.quota[l];
    }
    return h->P.nlevels;
}

int orbx_get_level_sizes(const orbx_handle *h, int width, int height, int *widths, int *heights) {
    if (!h || width < 1 || height < 1) return ORBX_E_INVALID;
    for (int l = 0; l < h->P.nlevels; l++) {
        if (widths) widths[l] = (int)lrintf((float)width * h->P.inv_scale[l]);
        if (heights) heights[l] = (int)lrintf((float)height * h->P.inv_scale[l]);
    }
    return h->P.nlevels;
}

long long orbx_launch_count(const orbx_handle *h) { return h ? h->launches : 0; }

int orbx_set_stream(orbx_handle *h, void *cuda_stream) {
    if (!h) return ORBX_E_INVALID;
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
    if (s == h->stream) return ORBX_OK;
    CU_TRY(h, cudaSetDevice(h->device));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    h->stream = s;
    return ORBX_OK;
}

int orbx_set_profiling(orbx_handle *h, int enable) {
    if (!h) return ORBX_E_INVALID;
    CU_TRY(h, cudaSetDevice(h->device));
    if (enable && !h->ev[0]) for (int i = 0; i < 7; i++) CU_TRY(h, cudaEventCreate(&h->ev[i]));
    h->profiling = enable != 0;
    h->ev_valid = false;
    return ORBX_OK;
}

int orbx_get_stage_times(orbx_handle *h, float *ms6) {
    if (!h || !ms6) return ORBX_E_INVALID;
    if (!h->ev_valid) return fail(h, ORBX_E_INVALID, "no profiled batch: call orbx_set_profiling(h, 1) before extracting");
    CU_TRY(h, cudaSetDevice(h->device));
    CU_TRY(h, cudaEventSynchronize(h->ev[6]));
    for (int i = 0; i < 6; i++) CU_TRY(h, cudaEventElapsedTime(&ms6[i], h->ev[i], h->ev[i + 1]));
    return ORBX_OK;
}

int orbx_set_input_format(orbx_handle *h, int format, int gray_shift) {
    if (!h) return ORBX_E_INVALID;
    if (format < ORBX_FMT_GRAY8 || format > ORBX_FMT_BGRA8 || (gray_shift != ORBX_GRAY_Q15 && gray_shift != ORBX_GRAY_Q14))
        return fail(h, ORBX_E_INVALID, "format must be ORBX_FMT_*, gray_shift ORBX_GRAY_Q15 or ORBX_GRAY_Q14");
    h->in_fmt = format; h->gray_shift = gray_shift;
    return ORBX_OK;
}

int orbx_debug_gray(orbx_handle *h, const uint8_t *src, int w, int ht, int stride, int format, int gray_shift, uint8_t *dst, int dstride) {
    if (!h || !src || !dst || w < 1 || ht < 1 || format < ORBX_FMT_RGB8 || format > ORBX_FMT_BGRA8 || dstride < w ||
        (gray_shift != ORBX_GRAY_Q15 && gray_shift != ORBX_GRAY_Q14))
        return ORBX_E_INVALID;
    const int bpp = format >= ORBX_FMT_RGBA8 ? 4 : 3;
    if (stride < w * bpp) return ORBX_E_INVALID;
    CU_TRY(h, cudaSetDevice(h->device));
    Scratch S;
    const int sp = (w * bpp + 15) / 16 * 16, dp = (w + 31) / 32 * 32;
    uint8_t *d_src = S.get<uint8_t>((size_t)sp * ht + 64), *d_dst = S.get<uint8_t>((size_t)dp * ht + 64);
    if (!d_src || !d_dst) return fail(h, ORBX_E_CUDA, "cudaMalloc failed");
    CU_TRY(h, cudaMemcpy2D(d_src, sp, src, stride, (size_t)w * bpp, ht, cudaMemcpyHostToDevice));
    CU_TRY(h, cudaStreamSynchronize(cudaStreamLegacy));   // blocking pageable uploads may still be in DMA; h->stream is non-blocking
    h->launches += launch_gray(d_src, (size_t)sp * ht, sp, format, gray_shift, d_dst, (size_t)dp * ht, dp, w, ht, 0, 1, h->stream);
    CU_TRY(h, cudaGetLastError());
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    CU_TRY(h, cudaMemcpy2D(dst, dstride, d_dst, dp, w, ht, cudaMemcpyDeviceToHost));
    return ORBX_OK;
}

int orbx_sync(orbx_handle *h) {
    if (!h) return ORBX_E_INVALID;
    CU_TRY(h, cudaSetDevice(h->device));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    if (h->planned) {   // surface an internal overflow of the last device-resident batch
        int ov = 0;
        CU_TRY(h, cudaMemcpy(&ov, h->d_overflow, sizeof(int), cudaMemcpyDeviceToHost));
        if (ov) {
            char msg[96]; snprintf(msg, sizeof(msg), "internal buffer overflow (stage code %d)", ov);
            return fail(h, ORBX_E_OVERFLOW, msg);
        }
    }
    return ORBX_OK;
}

int orbx_extract_batch_device(orbx_handle *h, const uint8_t *d_frames, size_t frame_stride_bytes, int batch, int width,
                              int height, int stride, int lap0, int lap1, orbx_keypoint *d_kp_out, uint8_t *d_desc_out,
                              int cap, int *d_n_out, int *d_mono_out) {
    if (!h) return ORBX_E_INVALID;
    if (h->pending.active) return fail(h, ORBX_E_INVALID, "a submitted batch has not been collected");
    if (!d_frames || !d_kp_out || !d_desc_out || !d_n_out || !d_mono_out) return fail(h, ORBX_E_INVALID, "null argument");
    if (width < 1 || height < 1) return fail(h, ORBX_E_EMPTY, "empty image");
    const int bpp = h->in_fmt == ORBX_FMT_GRAY8 ? 1 : h->in_fmt >= ORBX_FMT_RGBA8 ? 4 : 3, rowb = width * bpp;
    if (batch < 1 || stride < rowb || frame_stride_bytes < (size_t)stride * (height - 1) + rowb) return fail(h, ORBX_E_INVALID, "bad batch / stride");
    CU_TRY(h, cudaSetDevice(h->device));
    int rc = ensure_plan(h, width, height, batch);
    if (rc) return rc;
    if (cap < h->plan.total_out_cap) return fail(h, ORBX_E_CAPACITY, "cap smaller than orbx_keypoint_capacity for this frame size");
    ColorSrc color{nullptr, 0, 0};
    if (bpp == 1) {
        if ((rc = set_level0(h, d_frames, stride, frame_stride_bytes))) return rc;
    } else {
        color = ColorSrc{d_frames, stride, frame_stride_bytes};      // converted straight out of the caller's memory
        if ((rc = set_level0(h, h->l0_own, h->l0_own_pitch, h->l0_own_fstride))) return rc;
    }
    h->last_batch = batch;
    // Large batches are issued as several frame ranges on concurrent streams: the latency-bound kernels of one range (pyramid
    // chain, quadtree, slot assignment) then overlap the machine-filling ones of another (ORBX_DEV_SPLIT overrides the count).
    static const int split_env = [] { const char *e = getenv("ORBX_DEV_SPLIT"); return e ? atoi(e) : 0; }();
    int parts = split_env > 0 ? split_env : 4;   // measured on 64 x 640x480: 1 -> 126.0 k, 2 -> 127.1 k, 3 -> 131.7 k, 4 -> 131.7 k frames/s
    parts = std::min(std::min(parts, orbx_handle::kComputeStreams), batch / 8);
    if (h->profiling) parts = 1;
    if (parts > 1 && (rc = ensure_pipeline(h))) return rc;
    KeypointRec *d_kp = reinterpret_cast<KeypointRec *>(d_kp_out);
    auto enqueue = [&](bool) -> int {
        int rc2;
        if ((rc2 = reset_counters(h, h->stream))) return rc2;
        if (parts <= 1)
            return run_pipeline(h, 0, batch, lap0, lap1, d_kp, d_desc_out, cap, d_n_out, d_mono_out, h->stream, h->side_stream, h->ev_fork,
                                h->ev_join, color);
        CU_TRY(h, cudaEventRecord(h->ev_start, h->stream));
        const int per = (batch + parts - 1) / parts;
        int k = 0;
        for (int g0 = 0; g0 < batch; g0 += per, k++) {
            const int n = std::min(per, batch - g0);
            CU_TRY(h, cudaStreamWaitEvent(h->cs[k], h->ev_start, 0));
            if ((rc2 = run_pipeline(h, g0, n, lap0, lap1, d_kp, d_desc_out, cap, d_n_out, d_mono_out, h->cs[k], nullptr, nullptr, nullptr, color)))
                return rc2;
            CU_TRY(h, cudaEventRecord(h->ev_done[k], h->cs[k]));
            CU_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_done[k], 0));
        }
        return ORBX_OK;
    };
    // A streaming caller cycles through a few device buffers: a (buffers, geometry) key seen before is replayed as a CUDA graph,
    // which takes the ~60 API calls of a call off the host (with several ranks per host the enqueue cost otherwise limits scaling).
    if (!graphs_enabled() || h->profiling) return enqueue(true);
    GraphKey key{d_frames, d_kp_out, d_desc_out, d_n_out, d_mono_out, h->stream, batch, width, height, stride, lap0, lap1, cap,
                 0x1000 + parts, h->in_fmt, h->gray_shift, frame_stride_bytes};
    return run_graphed(h, key, enqueue, [] {});
}

// Asynchronous half of the host-buffer batch call: everything up to (not including) the wait for the device.  The call
// returns as soon as the copies and kernels are queued; orbx_extract_batch_collect waits and finishes the host side.
// A caller that keeps two handles busy (submit A, submit B, collect A, submit A', collect B, ...) overlaps the upload
// of one batch with the kernels and the download of the other from a single host thread.
// `blocking` = the caller will wait for this very batch (orbx_extract_batch): the batch is then issued as several frame ranges so that its
// own upload, kernels and download overlap.  A caller that streams batches through _submit / _collect over several handles gets the
// overlap from the batches in flight, and one range per batch is cheaper (measured, four handles, 64 x 640x480: 32-frame ranges
// 164.3 k frames/s, one 64-frame range 167.4 k of a 178.7 k copy ceiling; ORBX_CHUNK overrides both).
static int submit_impl(orbx_handle *h, const uint8_t *const *frames, int batch, int width, int height, int stride, int lap0,
                       int lap1, orbx_keypoint *kp_out, uint8_t *desc_out, int cap, bool blocking) {
    if (!h) return ORBX_E_INVALID;
    if (h->pending.active) return fail(h, ORBX_E_INVALID, "a submitted batch has not been collected");
    if (!frames || !kp_out || !desc_out) return fail(h, ORBX_E_INVALID, "null argument");
    if (width < 1 || height < 1) return fail(h, ORBX_E_EMPTY, "empty image");
    const int bpp = h->in_fmt == ORBX_FMT_GRAY8 ? 1 : h->in_fmt >= ORBX_FMT_RGBA8 ? 4 : 3, rowb = width * bpp;
    if (batch < 1 || stride < rowb) return fail(h, ORBX_E_INVALID, "bad batch / stride");
    for (int i = 0; i < batch; i++) if (!frames[i]) return fail(h, ORBX_E_INVALID, "null frame pointer");
    CU_TRY(h, cudaSetDevice(h->device));
    int rc = ensure_plan(h, width, height, batch);
    if (rc) return rc;
    if (bpp > 1 && (rc = ensure_color(h))) return rc;
    const int kc = h->kp_cap;
    // upload target: the level-0 planes themselves for gray frames, the colour staging planes otherwise
    uint8_t *const up_dev = bpp == 1 ? h->l0_own : h->d_color;
    uint8_t *const up_host = bpp == 1 ? h->h_in : h->h_color;
    const int pitch0 = bpp == 1 ? h->l0_own_pitch : h->color_pitch;
    const size_t fstride0 = bpp == 1 ? h->l0_own_fstride : h->color_fstride;
    const ColorSrc color = bpp == 1 ? ColorSrc{nullptr, 0, 0} : ColorSrc{h->d_color, h->color_pitch, h->color_fstride};
    // Input: page-locked caller memory is DMA'd straight into the level-0 planes (one strided 2D copy per frame, or a
    // single one per frame range when the frames are evenly spaced); pageable memory is first repacked into the
    // handle's pinned staging.  Output: page-locked caller buffers of the internal record capacity receive the
    // result blocks directly.
    const bool in_pinned = is_pinned_host(frames[0]) && is_pinned_host(frames[batch - 1] + (size_t)stride * (height - 1));
    const int width_b = rowb;   // bytes per row to move
    const bool out_direct = cap >= kc && is_pinned_host(kp_out) && is_pinned_host(desc_out);
    bool even = in_pinned && fstride0 == (size_t)pitch0 * height;
    for (int i = 1; i < batch && even; i++) even = (frames[i] - frames[i - 1]) == (ptrdiff_t)stride * height;
    // CPU half of the input path (pageable memory only): repack frames [f0, f0 + n) into the pinned staging
    auto stage_in = [&](int f0, int n) {
        if (in_pinned) return;
        for (int i = f0; i < f0 + n; i++) {
            uint8_t *dst = up_host + (size_t)i * fstride0;
            const uint8_t *src = frames[i];
            if (stride == pitch0) std::memcpy(dst, src, (size_t)stride * (height - 1) + width_b);
            else for (int y = 0; y < height; y++) std::memcpy(dst + (size_t)y * pitch0, src + (size_t)y * stride, (size_t)width_b);
        }
    };
    auto copy_in = [&](int f0, int n, cudaStream_t st) -> int {
        if (!in_pinned) {
            CU_TRY(h, cudaMemcpyAsync(up_dev + (size_t)f0 * fstride0, up_host + (size_t)f0 * fstride0, (size_t)n * fstride0, cudaMemcpyHostToDevice, st));
        } else if (even) {
            CU_TRY(h, cudaMemcpy2DAsync(up_dev + (size_t)f0 * fstride0, pitch0, frames[f0], stride, width_b, (size_t)height * n, cudaMemcpyHostToDevice, st));
        } else {
            for (int i = f0; i < f0 + n; i++)
                CU_TRY(h, cudaMemcpy2DAsync(up_dev + (size_t)i * fstride0, pitch0, frames[i], stride, width_b, height, cudaMemcpyHostToDevice, st));
        }
        return ORBX_OK;
    };
    auto download = [&](int f0, int n, cudaStream_t st) -> int {
        if (out_direct) {
            CU_TRY(h, cudaMemcpy2DAsync(kp_out + (size_t)f0 * cap, (size_t)cap * sizeof(KeypointRec), h->d_kp + (size_t)f0 * kc, (size_t)kc * sizeof(KeypointRec),
                                        (size_t)kc * sizeof(KeypointRec), n, cudaMemcpyDeviceToHost, st));
            CU_TRY(h, cudaMemcpy2DAsync(desc_out + (size_t)f0 * cap * 32, (size_t)cap * 32, h->d_desc + (size_t)f0 * kc * 32, (size_t)kc * 32, (size_t)kc * 32, n, cudaMemcpyDeviceToHost, st));
        } else {
            CU_TRY(h, cudaMemcpyAsync(h->h_kp + (size_t)f0 * kc, h->d_kp + (size_t)f0 * kc, (size_t)n * kc * sizeof(KeypointRec), cudaMemcpyDeviceToHost, st));
            CU_TRY(h, cudaMemcpyAsync(h->h_desc + (size_t)f0 * kc * 32, h->d_desc + (size_t)f0 * kc * 32, (size_t)n * kc * 32, cudaMemcpyDeviceToHost, st));
        }
        return ORBX_OK;
    };
    h->last_batch = batch;
    if ((rc = set_level0(h, h->l0_own, h->l0_own_pitch, h->l0_own_fstride))) return rc;
    static const bool chunk_forced = getenv("ORBX_CHUNK") != nullptr;
    const int chunk = h->profiling || (!blocking && !chunk_forced) ? batch : pipeline_chunk(batch);
    if (chunk < batch && (rc = ensure_pipeline(h))) return rc;
    // Everything the device does for this call, issued relative to the handle's stream.  `cpu_stage` = do the pageable
    // repacking inline (false when it was done up front because the device work is replayed from a CUDA graph).
    auto enqueue = [&](bool cpu_stage) -> int {
        int rc2;
        if ((rc2 = reset_counters(h, h->stream))) return rc2;
        if (chunk >= batch) {
            // one frame range: copy in, compute, copy out, all on the handle's stream (+ its side stream)
            if (cpu_stage) stage_in(0, batch);
            if ((rc2 = copy_in(0, batch, h->stream))) return rc2;
            if ((rc2 = run_pipeline(h, 0, batch, lap0, lap1, h->d_kp, h->d_desc, kc, h->d_n, h->d_mono, h->stream, h->side_stream, h->ev_fork, h->ev_join, color))) return rc2;
            if ((rc2 = download(0, batch, h->stream))) return rc2;
        } else {
            // Software pipeline over frame ranges: the copy engines stream range k+1 in and range k-1 out while the SMs
            // work on range k.  Ranges alternate between two compute streams so that the latency-bound quadtree kernel
            // of one range overlaps the machine-filling kernels of the next.
            CU_TRY(h, cudaEventRecord(h->ev_start, h->stream));
            CU_TRY(h, cudaStreamWaitEvent(h->h2d_stream, h->ev_start, 0));
            // each uploaded range is computed as `sub` sub-ranges on different streams (same reason as in the device-resident path)
            static const int sub_env = [] { const char *e = getenv("ORBX_SUB"); return e ? atoi(e) : 0; }();
            const int nchunks = (batch + chunk - 1) / chunk;
            // measured (frames/s, 64 x 640x480, blocking call | two handles alternated with submit / collect): 1 sub-range 81 k | 134-139 k,
            // 2 sub-ranges 82-87 k | 137-143 k, 4 (one 64-frame range) 58 k | 129 k
            int sub = sub_env > 0 ? sub_env : 2;
            while (sub > 1 && (chunk / sub < 8 || nchunks * sub > orbx_handle::kMaxChunks)) sub--;
            const int ncs = std::min(h->ncs, nchunks * sub);
            for (int i = 0; i < ncs; i++) CU_TRY(h, cudaStreamWaitEvent(h->cs[i], h->ev_start, 0));
            int k = 0, r = 0;
            for (int f0 = 0; f0 < batch; f0 += chunk, k++) {
                const int n = std::min(chunk, batch - f0);
                if (cpu_stage) stage_in(f0, n);
                if ((rc2 = copy_in(f0, n, h->h2d_stream))) return rc2;
                CU_TRY(h, cudaEventRecord(h->ev_in[k], h->h2d_stream));
                const int per = (n + sub - 1) / sub;
                for (int g0 = f0; g0 < f0 + n; g0 += per, r++) {
                    const int m = std::min(per, f0 + n - g0);
                    cudaStream_t cs = h->cs[r % ncs];
                    CU_TRY(h, cudaStreamWaitEvent(cs, h->ev_in[k], 0));
                    if ((rc2 = run_pipeline(h, g0, m, lap0, lap1, h->d_kp, h->d_desc, kc, h->d_n, h->d_mono, cs, nullptr, nullptr, nullptr, color))) return rc2;
                    CU_TRY(h, cudaEventRecord(h->ev_done[r], cs));
                    CU_TRY(h, cudaStreamWaitEvent(h->d2h_stream, h->ev_done[r], 0));
                    if ((rc2 = download(g0, m, h->d2h_stream))) return rc2;
                }
            }
            CU_TRY(h, cudaEventRecord(h->ev_end, h->d2h_stream));
            CU_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_end, 0));
        }
        CU_TRY(h, cudaMemcpyAsync(h->h_n, h->d_n, sizeof(int) * (2 * h->batch_cap + 1), cudaMemcpyDeviceToHost, h->stream));
        return ORBX_OK;
    };
    // CUDA-graph replay: a call shape seen before (same buffers, same geometry) is captured once and replayed, which
    // removes the ~100 API calls of the pipelined flow from the critical path.  Pageable input is graphed only in the
    // single-range form (the repacking then happens up front); page-locked buffers are part of the key.
    const bool graphable = graphs_enabled() && !h->profiling && (in_pinned ? even : chunk >= batch);
    if (!graphable) {
        if ((rc = enqueue(true))) return rc;
    } else {
        GraphKey key{in_pinned ? frames[0] : nullptr, out_direct ? (const void *)kp_out : nullptr, out_direct ? (const void *)desc_out : nullptr,
                     nullptr, nullptr, h->stream, batch, width, height, stride, lap0, lap1, out_direct ? cap : 0, chunk,
                     h->in_fmt, h->gray_shift, 0};
        if ((rc = run_graphed(h, key, enqueue, [&] { stage_in(0, batch); }))) return rc;
    }
    h->pending = orbx_handle::Pending{true, batch, cap, out_direct, kp_out, desc_out};
    return ORBX_OK;
}

int orbx_extract_batch_submit(orbx_handle *h, const uint8_t *const *frames, int batch, int width, int height, int stride, int lap0,
                              int lap1, orbx_keypoint *kp_out, uint8_t *desc_out, int cap) {
    return submit_impl(h, frames, batch, width, height, stride, lap0, lap1, kp_out, desc_out, cap, false);
}

int orbx_extract_batch_collect(orbx_handle *h, int *n_out, int *mono_index_out) {
    if (!h) return ORBX_E_INVALID;
    if (!h->pending.active) return fail(h, ORBX_E_INVALID, "no submitted batch to collect");
    if (!n_out || !mono_index_out) return fail(h, ORBX_E_INVALID, "null argument");
    const orbx_handle::Pending p = h->pending;
    h->pending.active = false;
    CU_TRY(h, cudaSetDevice(h->device));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    if (*h->h_overflow) {
        char msg[96]; snprintf(msg, sizeof(msg), "internal buffer overflow (stage code %d)", *h->h_overflow);
        return fail(h, ORBX_E_OVERFLOW, msg);
    }
    const int kc = h->kp_cap;
    for (int i = 0; i < p.batch; i++) {
        const int n = h->h_n[i];
        if (n > p.cap) return fail(h, ORBX_E_CAPACITY, "kp_out/desc_out capacity smaller than the number of keypoints");
        n_out[i] = n; mono_index_out[i] = h->h_mono[i];
        if (!p.out_direct) {
            std::memcpy(p.kp_out + (size_t)i * p.cap, h->h_kp + (size_t)i * kc, (size_t)n * sizeof(KeypointRec));
            std::memcpy(p.desc_out + (size_t)i * p.cap * 32, h->h_desc + (size_t)i * kc * 32, (size_t)n * 32);
        }
    }
    return ORBX_OK;
}

int orbx_extract_batch(orbx_handle *h, const uint8_t *const *frames, int batch, int width, int height, int stride, int lap0,
                       int lap1, orbx_keypoint *kp_out, uint8_t *desc_out, int cap, int *n_out, int *mono_index_out) {
    if (!h) return ORBX_E_INVALID;
    if (!frames || !kp_out || !desc_out || !n_out || !mono_index_out) return fail(h, ORBX_E_INVALID, "null argument");
    if (width < 1 || height < 1) {
        for (int i = 0; i < batch; i++) { n_out[i] = 0; mono_index_out[i] = -1; }
        return fail(h, ORBX_E_EMPTY, "empty image");
    }
    const int rc = submit_impl(h, frames, batch, width, height, stride, lap0, lap1, kp_out, desc_out, cap, true);
    return rc ? rc : orbx_extract_batch_collect(h, n_out, mono_index_out);
}

int orbx_extract(orbx_handle *h, const uint8_t *gray, int width, int height, int stride, int lap0, int lap1,
                 orbx_keypoint *kp_out, uint8_t *desc_out, int cap, int *n_out, int *mono_index_out) {
    const uint8_t *frames[1] = {gray};
    if (h && (!gray || width < 1 || height < 1)) {
        if (n_out) *n_out = 0;
        if (mono_index_out) *mono_index_out = -1;
        return fail(h, ORBX_E_EMPTY, "empty image");
    }
    return orbx_extract_batch(h, frames, 1, width, height, stride, lap0, lap1, kp_out, desc_out, cap, n_out, mono_index_out);
}

// ---- frames as they arrive on the reference's wire: binary PNM (`encoding: "ppm"`) -------------------------------
int orbx_extract_pnm(orbx_handle *h, const uint8_t *data, size_t nbytes, int camera_rgb, int lap0, int lap1, orbx_keypoint *kp_out,
                     uint8_t *desc_out, int cap, int *n_out, int *mono_index_out, int *width_out, int *height_out) {
    if (!h) return ORBX_E_INVALID;
    if (n_out) *n_out = 0;
    if (mono_index_out) *mono_index_out = -1;
    int w = 0, ht = 0, ch = 0; size_t off = 0;
    const int prc = orbx_pnm_header(data, nbytes, &w, &ht, &ch, &off);
    if (prc == ORBX_E_EMPTY) return fail(h, ORBX_E_EMPTY, "PNM data does not decode (the reference logs 'Failed to decode frame image data' and skips the frame)");
    if (prc) return fail(h, ORBX_E_INVALID, data ? "PNM variant outside binary 8-bit P5 / P6 (the extractor needs CV_8U)" : "null argument");
    if (width_out) *width_out = w;
    if (height_out) *height_out = ht;
    // cv::imdecode stores a P6 frame as BGR; Tracking::GrabImageMonocular then applies RGB2GRAY (Camera.RGB = 1) or BGR2GRAY to
    // that memory.  In terms of the payload's own byte order (R, G, B) that is the BGR weighting for Camera.RGB = 1 and the RGB
    // weighting otherwise -- the conversion reads the payload where it lies, no channel swap is materialised.
    const int saved = h->in_fmt;
    h->in_fmt = ch == 1 ? ORBX_FMT_GRAY8 : camera_rgb ? ORBX_FMT_BGR8 : ORBX_FMT_RGB8;
    const int rc = orbx_extract(h, data + off, w, ht, w * ch, lap0, lap1, kp_out, desc_out, cap, n_out, mono_index_out);
    h->in_fmt = saved;
    return rc;
}

int orbx_wire_process_frame(orbx_handle *h, const uint8_t *payload, size_t nbytes, int camera_rgb, int lap0, int lap1,
                            orbx_keypoint *kp_out, uint8_t *desc_out, int cap, int *n_out, int *mono_index_out, int *width_out,
                            int *height_out, double *timestamp_out, int *camera_id_out) {
    if (!h) return ORBX_E_INVALID;
    if (n_out) *n_out = 0;
    if (mono_index_out) *mono_index_out = -1;
    orbx_wire_frame m;
    if (orbx_wire_parse_frame(payload, nbytes, &m) != ORBX_OK) return fail(h, ORBX_E_INVALID, "Failed to parse MessagePack payload");
    if (m.type_len != 5 || std::memcmp(m.type, "frame", 5) != 0) return fail(h, ORBX_E_INVALID, "not a frame message");
    // the receive loop's checks, in its order (orbslam3_mono_networked.cc:527-544); each one skips the message there
    if (!m.has_camera_id || !m.camera_id) return fail(h, ORBX_E_EMPTY, "Frame message missing camera identifier.");
    if (!m.image || !m.image_bytes) return fail(h, ORBX_E_EMPTY, "Frame message missing binary image data.");
    if (!m.has_timestamp) return fail(h, ORBX_E_EMPTY, "Frame message missing timestamp.");
    if (timestamp_out) *timestamp_out = m.timestamp;
    if (camera_id_out) *camera_id_out = m.camera_id;
    return orbx_extract_pnm(h, m.image, m.image_bytes, camera_rgb, lap0, lap1, kp_out, desc_out, cap, n_out, mono_index_out, width_out, height_out);
}

// ---- stage inspection ------------------------------------------------------------------------------------------
int orbx_debug_get_level(orbx_handle *h, int frame, int level, int blurred, uint8_t *out, int out_stride, int *width_out, int *height_out) {
    if (!h || !out) return ORBX_E_INVALID;
    if (!h->planned || frame < 0 || frame >= h->last_batch || level < 0 || level >= h->plan.nlevels) return fail(h, ORBX_E_INVALID, "no such frame / level");
    CU_TRY(h, cudaSetDevice(h->device));
    const LevelDev &D = h->h_levels[level];
    if (out_stride < D.w) return fail(h, ORBX_E_CAPACITY, "out_stride < level width");
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    const uint8_t *src = blurred ? D.blur + (size_t)frame * D.blur_fstride : D.img + (size_t)frame * D.img_fstride;
    CU_TRY(h, cudaMemcpy2D(out, out_stride, src, blurred ? D.blur_pitch : D.pitch, D.w, D.h, cudaMemcpyDeviceToHost));
    if (width_out) *width_out = D.w;
    if (height_out) *height_out = D.h;
    return ORBX_OK;
}

int orbx_debug_get_candidates(orbx_handle *h, int frame, int level, float *xyr_out, int cap, int *n_out) {
    if (!h || !xyr_out || !n_out) return ORBX_E_INVALID;
    if (!h->planned || frame < 0 || frame >= h->last_batch || level < 0 || level >= h->plan.nlevels) return fail(h, ORBX_E_INVALID, "no such frame / level");
    CU_TRY(h, cudaSetDevice(h->device));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    const LevelDev &D = h->h_levels[level];
    int n = 0;
    CU_TRY(h, cudaMemcpy(&n, D.cand_count + frame, sizeof(int), cudaMemcpyDeviceToHost));
    *n_out = n;
    if (n > cap) return fail(h, ORBX_E_CAPACITY, "candidate buffer too small");
    std::vector<uint32_t> tmp((size_t)std::max(n, 1));
    CU_TRY(h, cudaMemcpy(tmp.data(), D.cand + (size_t)frame * D.cand_cap, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; i++) {
        xyr_out[3 * i] = (float)((tmp[i] >> 8) & 0xFFF); xyr_out[3 * i + 1] = (float)(tmp[i] >> 20); xyr_out[3 * i + 2] = (float)(tmp[i] & 0xFF);
    }
    return ORBX_OK;
}

int orbx_debug_get_level_keypoints(orbx_handle *h, int frame, int level, float *xyra_out, int cap, int *n_out) {
    if (!h || !xyra_out || !n_out) return ORBX_E_INVALID;
    if (!h->planned || frame < 0 || frame >= h->last_batch || level < 0 || level >= h->plan.nlevels) return fail(h, ORBX_E_INVALID, "no such frame / level");
    CU_TRY(h, cudaSetDevice(h->device));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    const LevelDev &D = h->h_levels[level];
    int n = 0;
    CU_TRY(h, cudaMemcpy(&n, D.sel_count + frame, sizeof(int), cudaMemcpyDeviceToHost));
    *n_out = n;
    if (n > cap) return fail(h, ORBX_E_CAPACITY, "keypoint buffer too small");
    std::vector<uint32_t> tmp((size_t)std::max(n, 1));
    std::vector<int> slots((size_t)std::max(n, 1));
    CU_TRY(h, cudaMemcpy(tmp.data(), D.sel + (size_t)frame * D.out_cap, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
    CU_TRY(h, cudaMemcpy(slots.data(), h->d_slot + (size_t)frame * h->plan.total_out_cap + D.out_base, sizeof(int) * n, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; i++) {
        KeypointRec r;
        CU_TRY(h, cudaMemcpy(&r, h->d_kp + (size_t)frame * h->kp_cap + slots[i], sizeof(r), cudaMemcpyDeviceToHost));
        xyra_out[4 * i] = (float)((tmp[i] >> 8) & 0xFFF); xyra_out[4 * i + 1] = (float)(tmp[i] >> 20);
        xyra_out[4 * i + 2] = (float)(tmp[i] & 0xFF); xyra_out[4 * i + 3] = r.angle;
    }
    return ORBX_OK;
}

int orbx_debug_resize(orbx_handle *h, const uint8_t *src, int sw, int sh, int sstride, uint8_t *dst, int dw, int dh, int dstride) {
    if (!h || !src || !dst || sw < 1 || sh < 1 || dw < 1 || dh < 1 || sstride < sw || dstride < dw) return ORBX_E_INVALID;
    CU_TRY(h, cudaSetDevice(h->device));
    // a two-level table built ad hoc: level 0 = src, level 1 = dst
    Scratch S;
    const int sp = (sw + 31) / 32 * 32, dp = (dw + 31) / 32 * 32;
    uint8_t *d_src = S.get<uint8_t>((size_t)sp * sh + 64), *d_dst = S.get<uint8_t>((size_t)dp * dh + 64);
    std::vector<ResizeTap> xt, yt;
    build_resize_taps(dw, sw, true, xt);
    build_resize_taps(dh, sh, false, yt);
    ResizeTap *d_xt = S.get<ResizeTap>(dw), *d_yt = S.get<ResizeTap>(dh);
    LevelDev lv[kMaxLevels]; std::memset(lv, 0, sizeof(lv));
    if (!d_src || !d_dst || !d_xt || !d_yt) return fail(h, ORBX_E_CUDA, "cudaMalloc failed");
    lv[0].img = d_src; lv[0].w = sw; lv[0].h = sh; lv[0].pitch = sp; lv[0].img_fstride = (size_t)sp * sh;
    lv[1].img = d_dst; lv[1].w = dw; lv[1].h = dh; lv[1].pitch = dp; lv[1].img_fstride = (size_t)dp * dh; lv[1].xtap = d_xt; lv[1].ytap = d_yt;
    {   // compact taps when they fit (same rule as build_plan)
        bool ok = true;
        for (int d = 0; d < dw && ok; d++) {
            const ResizeTap &tp = xt[d];
            ok = (tp.c0 + tp.c1 == 2048) && tp.c1 >= 0 && (tp.ofs1 == tp.ofs + 1 || tp.c1 == 0);
            if (ok && (d & 3) == 0) { const int last = d + 3 < dw ? d + 3 : dw - 1; ok = xt[last].ofs - tp.ofs <= 6 && xt[last].ofs >= tp.ofs; }
        }
        for (int d = 0; d < dh && ok; d++) ok = yt[d].c0 >= 0 && yt[d].c1 >= 0;
        if (ok) {
            std::vector<uint32_t> xp((size_t)(dw + 3) / 4 * 4);
            for (size_t d = 0; d < xp.size(); d++) { const ResizeTap &tp = xt[d < (size_t)dw ? d : (size_t)dw - 1]; xp[d] = ((uint32_t)tp.ofs << 16) | (uint32_t)tp.c1; }
            uint32_t *d_xp = S.get<uint32_t>(xp.size());
            if (!d_xp) return fail(h, ORBX_E_CUDA, "cudaMalloc failed");
            CU_TRY(h, cudaMemcpy(d_xp, xp.data(), xp.size() * 4, cudaMemcpyHostToDevice));
            lv[1].xpack = d_xp;
        }
    }
    CU_TRY(h, cudaMemcpy2D(d_src, sp, src, sstride, sw, sh, cudaMemcpyHostToDevice));
    CU_TRY(h, cudaMemcpy(d_xt, xt.data(), sizeof(ResizeTap) * dw, cudaMemcpyHostToDevice));
    CU_TRY(h, cudaMemcpy(d_yt, yt.data(), sizeof(ResizeTap) * dh, cudaMemcpyHostToDevice));
    CU_TRY(h, cudaStreamSynchronize(cudaStreamLegacy));   // blocking pageable uploads may still be in DMA; h->stream is non-blocking
    h->launches += launch_resize(lv, 1, 0, 1, h->stream);
    CU_TRY(h, cudaGetLastError());
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    CU_TRY(h, cudaMemcpy2D(dst, dstride, d_dst, dp, dw, dh, cudaMemcpyDeviceToHost));
    return ORBX_OK;
}

int orbx_debug_blur(orbx_handle *h, const uint8_t *src, int w, int ht, int sstride, uint8_t *dst, int dstride) {
    if (!h || !src || !dst || w < 1 || ht < 1 || sstride < w || dstride < w) return ORBX_E_INVALID;
    CU_TRY(h, cudaSetDevice(h->device));
    Scratch S;
    const int p = (w + 31) / 32 * 32;
    uint8_t *d_src = S.get<uint8_t>((size_t)p * ht + 64), *d_dst = S.get<uint8_t>((size_t)p * ht + 64);
    std::vector<BlurTile> tiles;
    for (int ty = 0; ty * kBlurTileH < ht; ty++) for (int tx = 0; tx * kBlurTileW < w; tx++) tiles.push_back(BlurTile{0, (int16_t)tx, (int16_t)ty, 0});
    BlurTile *d_t = S.get<BlurTile>(tiles.size());
    LevelDev lvs[kMaxLevels]; std::memset(lvs, 0, sizeof(lvs));
    LevelDev &lv = lvs[0];
    if (!d_src || !d_dst || !d_t) return fail(h, ORBX_E_CUDA, "cudaMalloc failed");
    lv.img = d_src; lv.blur = d_dst; lv.w = w; lv.h = ht; lv.pitch = p; lv.blur_pitch = p; lv.img_fstride = lv.blur_fstride = (size_t)p * ht;
    CU_TRY(h, cudaMemcpy2D(d_src, p, src, sstride, w, ht, cudaMemcpyHostToDevice));
    CU_TRY(h, cudaMemcpy(d_t, tiles.data(), sizeof(BlurTile) * tiles.size(), cudaMemcpyHostToDevice));
    CU_TRY(h, cudaStreamSynchronize(cudaStreamLegacy));   // blocking pageable uploads may still be in DMA; h->stream is non-blocking
    h->launches += launch_blur(lvs, d_t, (int)tiles.size(), 0, 1, h->stream);
    CU_TRY(h, cudaGetLastError());
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    CU_TRY(h, cudaMemcpy2D(dst, dstride, d_dst, p, w, ht, cudaMemcpyDeviceToHost));
    return ORBX_OK;
}

int orbx_debug_describe(orbx_handle *h, const uint8_t *img, const uint8_t *blurred, int w, int ht, int stride, const float *xy,
                        int n, const float *angle_in, float *angle_out, uint8_t *desc_out) {
    if (!h || !xy || n < 0 || w < 1 || ht < 1 || stride < w) return ORBX_E_INVALID;
    if (!img && !angle_in) return fail(h, ORBX_E_INVALID, "need img or angle_in");
    for (int i = 0; i < n; i++) {
        const float x = xy[2 * i], y = xy[2 * i + 1];
        if (!(x >= kEdge && x <= w - 1 - kEdge && y >= kEdge && y <= ht - 1 - kEdge)) return fail(h, ORBX_E_INVALID, "keypoint closer than 19 px to the border");
    }
    CU_TRY(h, cudaSetDevice(h->device));
    Scratch S;
    const int p = (w + 31) / 32 * 32;
    uint8_t *d_img = img ? S.get<uint8_t>((size_t)p * ht + 64) : nullptr, *d_bl = blurred ? S.get<uint8_t>((size_t)p * ht + 64) : nullptr;
    float *d_xy = S.get<float>((size_t)2 * std::max(n, 1)), *d_ai = angle_in ? S.get<float>(std::max(n, 1)) : nullptr, *d_ao = S.get<float>(std::max(n, 1));
    uint8_t *d_desc = S.get<uint8_t>((size_t)32 * std::max(n, 1));
    if ((img && !d_img) || (blurred && !d_bl) || !d_xy || !d_ao || !d_desc) return fail(h, ORBX_E_CUDA, "cudaMalloc failed");
    if (img) CU_TRY(h, cudaMemcpy2D(d_img, p, img, stride, w, ht, cudaMemcpyHostToDevice));
    if (blurred) CU_TRY(h, cudaMemcpy2D(d_bl, p, blurred, stride, w, ht, cudaMemcpyHostToDevice));
    CU_TRY(h, cudaMemcpy(d_xy, xy, sizeof(float) * 2 * n, cudaMemcpyHostToDevice));
    if (angle_in) CU_TRY(h, cudaMemcpy(d_ai, angle_in, sizeof(float) * n, cudaMemcpyHostToDevice));
    CU_TRY(h, cudaStreamSynchronize(cudaStreamLegacy));   // blocking pageable uploads may still be in DMA; h->stream is non-blocking
    h->launches += launch_describe_points(d_img, d_bl, p, d_xy, n, d_ai, d_ao, d_desc, h->stream);
    CU_TRY(h, cudaGetLastError());
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    if (angle_out) CU_TRY(h, cudaMemcpy(angle_out, d_ao, sizeof(float) * n, cudaMemcpyDeviceToHost));
    if (desc_out && blurred) CU_TRY(h, cudaMemcpy(desc_out, d_desc, (size_t)32 * n, cudaMemcpyDeviceToHost));
    return ORBX_OK;
}

int orbx_debug_octree(orbx_handle *h, const float *keys, int n, int minX, int maxX, int minY, int maxY, int N, int *out_idx,
                      int cap, int *n_out) {
    if (!h || (!keys && n > 0) || !out_idx || !n_out || n < 0 || N < 0) return ORBX_E_INVALID;
    // The quadtree kernel needs the level LUTs: synthesise a one-level plan whose region is (maxX-minX) x (maxY-minY),
    // i.e. a level of size (reg_w + 32) x (reg_h + 32) with minBorder 16.
    if (minX != kMinBorder || minY != kMinBorder) return fail(h, ORBX_E_INVALID, "octree debug entry expects minX = minY = 16");
    CU_TRY(h, cudaSetDevice(h->device));
    ExtractorParams P1 = h->P;
    P1.nlevels = 1; P1.quota[0] = N;
    Plan pl; std::string err;
    if (!build_plan(P1, maxX + kMinBorder, maxY + kMinBorder, pl, err)) { h->err = err; return ORBX_E_INVALID; }
    const LevelPlan &LP = pl.lv[0];
    if (LP.ncols <= 0 || LP.nrows <= 0) { *n_out = 0; return ORBX_OK; }
    Scratch S;
    LevelDev Ds[kMaxLevels]; std::memset(Ds, 0, sizeof(Ds));
    LevelDev &D = Ds[0];
    uint32_t *t0 = S.get<uint32_t>(LP.xbin.size()), *t1 = S.get<uint32_t>(LP.ybin.size()), *t2 = S.get<uint32_t>(LP.xord.size()), *t3 = S.get<uint32_t>(LP.yord.size());
    const int ccap = std::max(n, 1);
    uint32_t *d_cand = S.get<uint32_t>(ccap), *d_sorted = S.get<uint32_t>(ccap), *d_cur = S.get<uint32_t>(std::max(LP.nbins, 1)), *d_sel = S.get<uint32_t>(LP.out_cap);
    int *d_cnt = S.get<int>(4);
    if (!t0 || !t1 || !t2 || !t3 || !d_cand || !d_sorted || !d_cur || !d_sel || !d_cnt) return fail(h, ORBX_E_CUDA, "cudaMalloc failed");
    CU_TRY(h, cudaMemcpy(t0, LP.xbin.data(), LP.xbin.size() * 4, cudaMemcpyHostToDevice));
    CU_TRY(h, cudaMemcpy(t1, LP.ybin.data(), LP.ybin.size() * 4, cudaMemcpyHostToDevice));
    CU_TRY(h, cudaMemcpy(t2, LP.xord.data(), LP.xord.size() * 4, cudaMemcpyHostToDevice));
    CU_TRY(h, cudaMemcpy(t3, LP.yord.data(), LP.yord.size() * 4, cudaMemcpyHostToDevice));
    // pack keys; remember the original index of every (x, y) to map the winners back
    std::vector<uint32_t> packed((size_t)ccap);
    std::vector<int> index_of((size_t)LP.reg_w * LP.reg_h, -1);
    for (int i = 0; i < n; i++) {
        const int x = (int)keys[3 * i], y = (int)keys[3 * i + 1], r = (int)keys[3 * i + 2];
        if (x < 3 || y < 3 || x >= LP.reg_w - 3 || y >= LP.reg_h - 3 || r < 0 || r > 255 || (float)x != keys[3 * i] || (float)y != keys[3 * i + 1])
            return fail(h, ORBX_E_INVALID, "octree debug entry: keys must be integer FAST candidates inside the tested region");
        packed[i] = ((uint32_t)y << 20) | ((uint32_t)x << 8) | (uint32_t)r;
        if (index_of[(size_t)y * LP.reg_w + x] >= 0) return fail(h, ORBX_E_INVALID, "duplicate key position");
        index_of[(size_t)y * LP.reg_w + x] = i;
    }
    CU_TRY(h, cudaMemcpy(d_cand, packed.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice));
    int counts[4] = {n, 0, 0, 0};
    CU_TRY(h, cudaMemcpy(d_cnt, counts, sizeof(counts), cudaMemcpyHostToDevice));
    D.xbin = t0; D.ybin = t1; D.xord = t2; D.yord = t3;
    D.w = LP.w; D.h = LP.h;
    D.reg_w = LP.reg_w; D.reg_h = LP.reg_h; D.n_ini = LP.n_ini; D.depth0 = LP.depth0; D.nbins = LP.nbins; D.quota = N; D.out_cap = LP.out_cap;
    for (int r = 0; r < kMaxRoots; r++) { D.root_ulx[r] = LP.root_ulx[r]; D.root_brx[r] = LP.root_brx[r]; }
    D.ord_cell_area = LP.ord_cell_area; D.ord_ncols = LP.ord_ncols; D.wcell = LP.wcell; D.hcell = LP.hcell;
    D.cand = d_cand; D.sorted = d_sorted; D.bin_cursor = d_cur; D.cand_cap = ccap; D.cand_count = d_cnt; D.sel = d_sel; D.sel_count = d_cnt + 1;
    CU_TRY(h, cudaStreamSynchronize(cudaStreamLegacy));   // blocking pageable uploads may still be in DMA; h->stream is non-blocking
    h->launches += launch_octree(Ds, 1, 0, 1, d_cnt + 2, h->stream);
    CU_TRY(h, cudaGetLastError());
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    CU_TRY(h, cudaMemcpy(counts, d_cnt, sizeof(counts), cudaMemcpyDeviceToHost));
    if (counts[2]) return fail(h, ORBX_E_OVERFLOW, "quadtree node buffer overflow");
    const int m = counts[1];
    *n_out = m;
    if (m > cap) return fail(h, ORBX_E_CAPACITY, "out_idx too small");
    std::vector<uint32_t> sel((size_t)std::max(m, 1));
    CU_TRY(h, cudaMemcpy(sel.data(), d_sel, sizeof(uint32_t) * m, cudaMemcpyDeviceToHost));
    for (int i = 0; i < m; i++) {
        const int x = (int)((sel[i] >> 8) & 0xFFF) - kMinBorder, y = (int)(sel[i] >> 20) - kMinBorder;
        out_idx[i] = index_of[(size_t)y * LP.reg_w + x];
    }
    return ORBX_OK;
}

// ---- matching entry points on the extractor handle (kernels in orbx_match.cu) --------------------------------------
int orbx_distance_batch(orbx_handle *h, const uint8_t *a, const uint8_t *b, int n, int32_t *dist_out) {
    if (!h) return ORBX_E_INVALID;
    if (!a || !b || !dist_out || n < 0) return fail(h, ORBX_E_INVALID, "null argument");
    return match_distance_batch(h->device, h->stream, a, b, n, dist_out, h->err, h->launches);
}

int orbx_match_windowed(orbx_handle *h, const uint8_t *q_desc, const float *q_uvr, const int32_t *q_levels, int nq,
                        const orbx_keypoint *t_kp, const uint8_t *t_desc, int nt, const float *bounds4, int32_t *best_idx,
                        int32_t *best_dist, int32_t *second_idx, int32_t *second_dist) {
    if (!h) return ORBX_E_INVALID;
    if (!q_desc || !q_uvr || !q_levels || !bounds4 || !best_idx || !best_dist || !second_idx || !second_dist || nq < 0 || nt < 0 ||
        (nt > 0 && (!t_kp || !t_desc)))
        return fail(h, ORBX_E_INVALID, "null argument");
    if (!(bounds4[2] > bounds4[0]) || !(bounds4[3] > bounds4[1])) return fail(h, ORBX_E_INVALID, "empty image bounds");
    return match_windowed(h->device, h->stream, q_desc, q_uvr, q_levels, nq, t_kp, t_desc, nt, bounds4, best_idx, best_dist,
                          second_idx, second_dist, h->err, h->launches);
}

// ---- Frame post-extraction steps (kernels in orbx_frame.cu) -----------------------------------------------------------------
static_assert(sizeof(orbx_camera) == sizeof(CameraDev), "orbx_camera layout");
static CameraDev to_dev(const orbx_camera *c) { CameraDev d; std::memcpy(&d, c, sizeof(d)); return d; }
static bool camera_ok(const orbx_camera *c) { return c && c->fx != 0.f && c->fy != 0.f; }

int orbx_undistort_points(orbx_handle *h, const float *xy, int n, const orbx_camera *cam, float *xy_out) {
    if (!h) return ORBX_E_INVALID;
    if (!xy || !xy_out || n < 0 || !camera_ok(cam)) return fail(h, ORBX_E_INVALID, "null argument / bad camera");
    if (n == 0) return ORBX_OK;
    CU_TRY(h, cudaSetDevice(h->device));
    Scratch S;
    float *d_in = S.get<float>((size_t)2 * n), *d_out = S.get<float>((size_t)2 * n);
    if (!d_in || !d_out) return fail(h, ORBX_E_CUDA, "cudaMalloc failed");
    CU_TRY(h, cudaMemcpyAsync(d_in, xy, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, h->stream));
    h->launches += launch_undistort_xy(d_in, n, to_dev(cam), d_out, h->stream);
    CU_TRY(h, cudaGetLastError());
    CU_TRY(h, cudaMemcpyAsync(xy_out, d_out, sizeof(float) * 2 * n, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    return ORBX_OK;
}

int orbx_image_bounds(orbx_handle *h, const orbx_camera *cam, int width, int height, float *bounds4_out) {
    if (!h) return ORBX_E_INVALID;
    if (!bounds4_out || width < 1 || height < 1 || !camera_ok(cam)) return fail(h, ORBX_E_INVALID, "null argument / bad camera");
    if (cam->k1 == 0.f) {   // Frame::ComputeImageBounds: mDistCoef.at<float>(0) == 0.0
        bounds4_out[0] = 0.f; bounds4_out[1] = 0.f; bounds4_out[2] = (float)width; bounds4_out[3] = (float)height;
        return ORBX_OK;
    }
    const float corners[8] = {0.f, 0.f, (float)width, 0.f, 0.f, (float)height, (float)width, (float)height};
    float un[8];
    const int rc = orbx_undistort_points(h, corners, 4, cam, un);
    if (rc) return rc;
    bounds4_out[0] = std::min(un[0], un[4]); bounds4_out[2] = std::max(un[2], un[6]);     // mnMinX, mnMaxX
    bounds4_out[1] = std::min(un[1], un[3]); bounds4_out[3] = std::max(un[5], un[7]);     // mnMinY, mnMaxY
    return ORBX_OK;
}

int orbx_frame_grid_batch_device(orbx_handle *h, const orbx_keypoint *d_kp, const int *d_n, int batch, int cap, const orbx_camera *cam,
                                 const float *bounds4, orbx_keypoint *d_kp_un, int32_t *d_cell_start, int32_t *d_cell_items) {
    if (!h) return ORBX_E_INVALID;
    if (!d_kp || !d_n || !bounds4 || !d_kp_un || !d_cell_start || !d_cell_items || batch < 1 || cap < 1 || !camera_ok(cam))
        return fail(h, ORBX_E_INVALID, "null argument / bad camera");
    if (!(bounds4[2] > bounds4[0]) || !(bounds4[3] > bounds4[1])) return fail(h, ORBX_E_INVALID, "empty image bounds");
    CU_TRY(h, cudaSetDevice(h->device));
    h->launches += launch_frame_grid(reinterpret_cast<const KeypointRec *>(d_kp), d_n, 0, batch, cap, to_dev(cam), bounds4,
                                     reinterpret_cast<KeypointRec *>(d_kp_un), d_cell_start, d_cell_items, h->stream);
    CU_TRY(h, cudaGetLastError());
    return ORBX_OK;
}

int orbx_frame_grid(orbx_handle *h, const orbx_keypoint *kp, int n, const orbx_camera *cam, const float *bounds4, orbx_keypoint *kp_un_out,
                    int32_t *cell_start_out, int32_t *cell_items_out) {
    if (!h) return ORBX_E_INVALID;
    if ((!kp && n > 0) || n < 0 || !bounds4 || (!kp_un_out && n > 0) || !cell_start_out || (!cell_items_out && n > 0) || !camera_ok(cam))
        return fail(h, ORBX_E_INVALID, "null argument / bad camera");
    if (!(bounds4[2] > bounds4[0]) || !(bounds4[3] > bounds4[1])) return fail(h, ORBX_E_INVALID, "empty image bounds");
    CU_TRY(h, cudaSetDevice(h->device));
    Scratch S;
    const int cap = std::max(n, 1);
    KeypointRec *d_kp = S.get<KeypointRec>(cap), *d_un = S.get<KeypointRec>(cap);
    int32_t *d_start = S.get<int32_t>(ORBX_GRID_CELLS + 1), *d_items = S.get<int32_t>(cap);
    if (!d_kp || !d_un || !d_start || !d_items) return fail(h, ORBX_E_CUDA, "cudaMalloc failed");
    if (n > 0) CU_TRY(h, cudaMemcpyAsync(d_kp, kp, sizeof(KeypointRec) * n, cudaMemcpyHostToDevice, h->stream));
    h->launches += launch_frame_grid(d_kp, nullptr, n, 1, cap, to_dev(cam), bounds4, d_un, d_start, d_items, h->stream);
    CU_TRY(h, cudaGetLastError());
    if (n > 0) CU_TRY(h, cudaMemcpyAsync(kp_un_out, d_un, sizeof(KeypointRec) * n, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaMemcpyAsync(cell_start_out, d_start, sizeof(int32_t) * (ORBX_GRID_CELLS + 1), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    const int inside = cell_start_out[ORBX_GRID_CELLS];
    if (inside > 0) CU_TRY(h, cudaMemcpy(cell_items_out, d_items, sizeof(int32_t) * inside, cudaMemcpyDeviceToHost));
    return ORBX_OK;
}

int orbx_match_windowed_grid_device(orbx_handle *h, const uint8_t *d_q_desc, const float *d_q_uvr, const int32_t *d_q_levels, int nq,
                                    const orbx_keypoint *d_t_kp_un, const uint8_t *d_t_desc, const int32_t *d_cell_start,
                                    const int32_t *d_cell_items, const float *bounds4, int32_t *d_best_idx, int32_t *d_best_dist,
                                    int32_t *d_second_idx, int32_t *d_second_dist) {
    if (!h) return ORBX_E_INVALID;
    if (nq < 0 || !bounds4 || (nq > 0 && (!d_q_desc || !d_q_uvr || !d_q_levels || !d_t_kp_un || !d_t_desc || !d_cell_start || !d_cell_items ||
                                           !d_best_idx || !d_best_dist || !d_second_idx || !d_second_dist)))
        return fail(h, ORBX_E_INVALID, "null argument");
    if (((uintptr_t)d_q_desc | (uintptr_t)d_t_desc) & 15) return fail(h, ORBX_E_INVALID, "descriptor arrays must be 16-byte aligned");
    if (!(bounds4[2] > bounds4[0]) || !(bounds4[3] > bounds4[1])) return fail(h, ORBX_E_INVALID, "empty image bounds");
    CU_TRY(h, cudaSetDevice(h->device));
    h->launches += match_windowed_grid_device(h->stream, d_q_desc, d_q_uvr, d_q_levels, nq, reinterpret_cast<const KeypointRec *>(d_t_kp_un), d_t_desc,
                                              d_cell_start, d_cell_items, bounds4, d_best_idx, d_best_dist, d_second_idx, d_second_dist);
    CU_TRY(h, cudaGetLastError());
    return ORBX_OK;
}

int orbx_match_windowed_grid_batch_device(orbx_handle *h, int npairs, const int32_t *pair_query_frame, const int32_t *pair_train_frame, int batch,
                                          int cap, const uint8_t *d_q_desc, const float *d_q_uvr, const int32_t *d_q_levels, const int32_t *d_n,
                                          const orbx_keypoint *d_t_kp_un, const uint8_t *d_t_desc, const int32_t *d_cell_start,
                                          const int32_t *d_cell_items, const float *bounds4, int32_t *d_best_idx, int32_t *d_best_dist,
                                          int32_t *d_second_idx, int32_t *d_second_dist) {
    if (!h) return ORBX_E_INVALID;
    if (npairs < 0 || batch < 1 || cap < 1 || !bounds4 ||
        (npairs > 0 && (!pair_query_frame || !pair_train_frame || !d_q_desc || !d_q_uvr || !d_q_levels || !d_n || !d_t_kp_un || !d_t_desc ||
                        !d_cell_start || !d_cell_items || !d_best_idx || !d_best_dist || !d_second_idx || !d_second_dist)))
        return fail(h, ORBX_E_INVALID, "null argument");
    for (int i = 0; i < npairs; i++)
        if (pair_query_frame[i] < 0 || pair_query_frame[i] >= batch || pair_train_frame[i] < 0 || pair_train_frame[i] >= batch)
            return fail(h, ORBX_E_INVALID, "frame index of a pair outside the batch");
    if (((uintptr_t)d_q_desc | (uintptr_t)d_t_desc) & 15) return fail(h, ORBX_E_INVALID, "descriptor arrays must be 16-byte aligned");
    if (!(bounds4[2] > bounds4[0]) || !(bounds4[3] > bounds4[1])) return fail(h, ORBX_E_INVALID, "empty image bounds");
    if (npairs == 0) return ORBX_OK;
    CU_TRY(h, cudaSetDevice(h->device));
    h->launches += match_windowed_grid_batch_device(h->stream, npairs, pair_query_frame, pair_train_frame, cap, d_q_desc, d_q_uvr, d_q_levels, d_n,
                                                    reinterpret_cast<const KeypointRec *>(d_t_kp_un), d_t_desc, d_cell_start, d_cell_items, bounds4,
                                                    d_best_idx, d_best_dist, d_second_idx, d_second_dist);
    CU_TRY(h, cudaGetLastError());
    return ORBX_OK;
}

// ---- plan inspection without a GPU (used by the CPU-only tests) --------------------------------------------------
// Fills level sizes, cell counts, quotas and candidate capacities of the plan for (params, w, h); returns nlevels or < 0.
int orbx_plan_probe(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th, int width, int height,
                    int *widths, int *heights, int *ncells, int *quota, int *n_ini, int *depth0, long long *algorithmic_bytes) {
    ExtractorParams P;
    if (!init_params(P, nfeatures, scale_factor, nlevels, ini_th, min_th)) return ORBX_E_INVALID;
    Plan pl; std::string err;
    if (!build_plan(P, width, height, pl, err)) { g_create_error = err; return ORBX_E_INVALID; }
    for (int l = 0; l < nlevels; l++) {
        if (widths) widths[l] = pl.lv[l].w;
        if (heights) heights[l] = pl.lv[l].h;
        if (ncells) ncells[l] = pl.lv[l].ncells;
        if (quota) quota[l] = pl.lv[l].quota;
        if (n_ini) n_ini[l] = pl.lv[l].n_ini;
        if (depth0) depth0[l] = pl.lv[l].depth0;
    }
    if (algorithmic_bytes) *algorithmic_bytes = (long long)pl.algorithmic_bytes(nfeatures);
    return nlevels;
}

}  // extern "C"

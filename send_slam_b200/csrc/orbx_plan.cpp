// Host geometry plan (see orbx_plan.h).  Pure C++, no CUDA: unit-testable without a GPU through the
// orbx_plan_* debug exports at the bottom of orbx_api.cu.
#include "orbx_plan.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>

namespace orbx {

static inline int cv_round(float v) { return (int)lrintf(v); }
static inline int cv_round(double v) { return (int)lrint(v); }

// ORBextractor::ORBextractor (UPSTREAM src/ORBextractor.cc ctor; SURVEY.md C.1 "Constructor")
bool init_params(ExtractorParams &p, int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th) {
    if (nlevels < 1 || nlevels > kMaxLevels || nfeatures < 0 || nfeatures > (1 << 20)) return false;
    if (!(scale_factor > 1.0f) || !(scale_factor < 2.0f)) return false;   // == 2 would take cv::resize's INTER_AREA path
    if (ini_th < 1 || ini_th > 254 || min_th < 1 || min_th > ini_th) return false;
    p.nfeatures = nfeatures; p.scale_factor_f = scale_factor; p.nlevels = nlevels; p.ini_th = ini_th; p.min_th = min_th;
    const double sf = (double)scale_factor;   // float ctor argument stored in a double member
    p.scale[0] = 1.0f; p.sigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) {
        p.scale[i] = (float)((double)p.scale[i - 1] * sf);
        p.sigma2[i] = p.scale[i] * p.scale[i];
    }
    for (int i = 0; i < nlevels; i++) { p.inv_scale[i] = 1.0f / p.scale[i]; p.inv_sigma2[i] = 1.0f / p.sigma2[i]; }
    const float factor = (float)(1.0 / sf);
    float want = (float)nfeatures * (1.0f - factor) / (1.0f - (float)std::pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; l++) { p.quota[l] = cv_round(want); sum += p.quota[l]; want *= factor; }
    p.quota[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;
    // umax: quarter-circle half-widths of the 31-px patch
    std::memset(p.umax, 0, sizeof(p.umax));
    const int vmax = (int)std::floor(kHalfPatch * std::sqrt(2.0) / 2 + 1), vmin = (int)std::ceil(kHalfPatch * std::sqrt(2.0) / 2);
    for (int v = 0; v <= vmax; v++) p.umax[v] = cv_round(std::sqrt((double)(kHalfPatch * kHalfPatch - v * v)));
    for (int v = kHalfPatch, v0 = 0; v >= vmin; --v) {
        while (p.umax[v0] == p.umax[v0 + 1]) ++v0;
        p.umax[v] = v0; ++v0;
    }
    return true;
}

// cv::resize INTER_LINEAR coefficient tables, 8UC1.  is_x: the x axis clamps the fraction at the
// borders, the y axis clips the row index on access instead.
void build_resize_taps(int dst, int src, bool is_x, std::vector<ResizeTap> &taps) {
    taps.resize(dst);
    const double scale = 1.0 / ((double)dst / (double)src);
    for (int d = 0; d < dst; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)std::floor(f);
        f -= (float)s;
        if (is_x) {
            if (s < 0) { f = 0.f; s = 0; }
            if (s >= src - 1) { f = 0.f; s = src - 1; }
        }
        auto sat16 = [](int v) { return (int16_t)(v < -32768 ? -32768 : v > 32767 ? 32767 : v); };
        ResizeTap t;
        t.c0 = sat16(cv_round((1.f - f) * 2048.f));
        t.c1 = sat16(cv_round(f * 2048.f));
        int s0 = s < 0 ? 0 : (s > src - 1 ? src - 1 : s);
        int s1 = s + 1 < 0 ? 0 : (s + 1 > src - 1 ? src - 1 : s + 1);
        t.ofs = (int16_t)s0; t.ofs1 = (int16_t)s1;
        taps[d] = t;
    }
}

static inline uint32_t spread_bits(uint32_t v) {   // bit i -> bit 2i
    uint32_t r = 0;
    for (int i = 0; i < 16; i++) r |= ((v >> i) & 1u) << (2 * i);
    return r;
}

static inline int floor4(int v) { return v & ~3; }
static inline int ceil4(int v) { return (v + 3) & ~3; }

// One attempt with a given tile grid; false when a region does not fit the kernel's limits.
static bool cone_try(const Plan &plan, int src, int last, int gx, int gy, ConePlan &c) {
    const int NL = last - src + 1;
    c = ConePlan();
    c.src = src; c.last = last; c.gx = gx; c.gy = gy; c.ntiles = gx * gy;
    c.lv.assign((size_t)c.ntiles * NL, ConeLevel{0, 0, 0, 0, 0, 0, 0, 0});
    int pitch = 0, b0 = 0, b1 = 0;
    // what the region `r` of level l reads from level l - 1: columns [nx0, nx1), rows [ny0, ny1)
    auto needs = [&](int l, const ConeLevel &r, int &nx0, int &nx1, int &ny0, int &ny1) {
        const LevelPlan &N = plan.lv[l];
        const int a = r.rx0, b = r.rx0 + r.rw;
        nx0 = N.xtap[std::min(a, N.w - 1)].ofs; nx1 = N.xtap[std::min(b - 1, N.w - 1)].ofs + 2;
        const int r0 = r.ry0, r1 = std::min(r.ry0 + ceil4(r.rh) - 1, N.h - 1);
        ny0 = N.ytap[r0].ofs; ny1 = N.ytap[r1].ofs1 + 1;
        for (int y = r0; y <= r1; y++) { ny0 = std::min<int>(ny0, N.ytap[y].ofs); ny1 = std::max<int>(ny1, N.ytap[y].ofs1 + 1); }
    };
    for (int j = 0; j < gy; j++)
        for (int i = 0; i < gx; i++) {
            ConeLevel *R = &c.lv[(size_t)(j * gx + i) * NL] - src;      // R[l] = entry of level l
            for (int l = last; l > src; l--) {
                const int W = plan.lv[l].w, H = plan.lv[l].h;
                int ox0 = i == 0 ? 0 : floor4((int)((long long)i * W / gx)), ox1 = i == gx - 1 ? W : floor4((int)((long long)(i + 1) * W / gx));
                int oy0 = (int)((long long)j * H / gy), oy1 = j == gy - 1 ? H : (int)((long long)(j + 1) * H / gy);
                if (ox1 <= ox0 || oy1 <= oy0) { ox1 = ox0; oy1 = oy0; }
                int x0 = ox0, x1 = ox1, y0 = oy0, y1 = oy1;
                if (l < last && R[l + 1].rw > 0 && R[l + 1].rh > 0) {
                    int nx0, nx1, ny0, ny1;
                    needs(l + 1, R[l + 1], nx0, nx1, ny0, ny1);
                    if (x1 > x0 && y1 > y0) { x0 = std::min(x0, nx0); x1 = std::max(x1, nx1); y0 = std::min(y0, ny0); y1 = std::max(y1, ny1); }
                    else { x0 = nx0; x1 = nx1; y0 = ny0; y1 = ny1; }
                }
                const int rx0 = floor4(x0), rx1 = std::min(ceil4(x1), ceil4(W));
                R[l] = ConeLevel{(int16_t)rx0, (int16_t)y0, (int16_t)std::max(rx1 - rx0, 0), (int16_t)std::max(std::min(y1, H) - y0, 0),
                                 (int16_t)ox0, (int16_t)ox1, (int16_t)oy0, (int16_t)oy1};
                pitch = std::max(pitch, R[l].rw + 16);
            }
            if (R[src + 1].rw > 0 && R[src + 1].rh > 0) {               // the TMA box of the source level
                int nx0, nx1, ny0, ny1;
                needs(src + 1, R[src + 1], nx0, nx1, ny0, ny1);
                const int bx0 = nx0 & ~15;
                R[src] = ConeLevel{(int16_t)bx0, (int16_t)ny0, (int16_t)(nx1 - bx0), (int16_t)(ny1 - ny0), 0, 0, 0, 0};
                c.box_w = std::max(c.box_w, (nx1 + 10 - bx0 + 15) / 16 * 16);   // the 3-word read window of an item reaches 11 bytes past its first word
                c.box_h = std::max(c.box_h, ny1 - ny0);
            }
        }
    c.pitch = (pitch + 3) & ~3;
    for (int t = 0; t < c.ntiles; t++)
        for (int k = 1; k < NL; k++) {
            const ConeLevel &r = c.lv[(size_t)t * NL + k];
            const int bytes = c.pitch * ceil4(r.rh) + 16;
            if (k & 1) b1 = std::max(b1, bytes); else b0 = std::max(b0, bytes);
        }
    c.buf0_bytes = (std::max(b0, c.box_w * c.box_h + 16) + 127) / 128 * 128;
    c.buf1_bytes = (b1 + 127) / 128 * 128;
    if (c.box_w < 16 || c.box_h < 1 || c.box_w > 256 || c.box_h > 256) return false;
    if (c.buf0_bytes + c.buf1_bytes > 56 * 1024) return false;
    // self-check: the own parts partition every level and lie inside their regions
    for (int k = 1; k < NL; k++) {
        long long area = 0;
        for (int t = 0; t < c.ntiles; t++) {
            const ConeLevel &r = c.lv[(size_t)t * NL + k];
            area += (long long)(r.ox1 - r.ox0) * (r.oy1 - r.oy0);
            if (r.ox1 > r.ox0 && (r.ox0 < r.rx0 || r.ox1 > r.rx0 + r.rw || r.oy0 < r.ry0 || r.oy1 > r.ry0 + r.rh || (r.ox0 & 3))) return false;
            if ((r.rx0 & 3) || (r.rw & 3) || r.rw > 240 || ceil4(r.rh) > 272) return false;   // the kernel's tap staging arrays
        }
        if (area != (long long)plan.lv[src + k].w * plan.lv[src + k].h) return false;
    }
    c.ok = true;
    return true;
}

// Tiles of about 128 x 120 pixels of the source level (a ~22 KB box + a ~14 KB first region per CTA); finer grids until everything fits.
bool build_cone_plan(const Plan &plan, int src, int last, ConePlan &cone) {
    cone = ConePlan();
    if (src < 0 || last <= src || last >= plan.nlevels) return false;
    for (int l = src + 1; l <= last; l++) if (plan.lv[l].xpack.empty() || plan.lv[l].w < 8 || plan.lv[l].h < 8) return false;
    int gx = std::max(1, (plan.lv[src].w + 64) / 128), gy = std::max(1, (plan.lv[src].h + 60) / 120);
    for (int attempt = 0; attempt < 6; attempt++) {
        if (cone_try(plan, src, last, gx, gy, cone)) return true;
        gx = gx * 5 / 4 + 1; gy = gy * 5 / 4 + 1;
    }
    cone = ConePlan();
    return false;
}

bool build_plan(const ExtractorParams &p, int width, int height, Plan &plan, std::string &err) {
    if (width < 1 || height < 1 || width > 4095 || height > 4095) { err = "frame size out of range (1..4095)"; return false; }
    plan = Plan();
    plan.width = width; plan.height = height; plan.nlevels = p.nlevels;
    plan.lv.resize(p.nlevels);
    for (int l = 0; l < p.nlevels; l++) {
        LevelPlan &L = plan.lv[l];
        // ComputePyramid: Size(cvRound((float)cols*scale), cvRound((float)rows*scale)), scale = mvInvScaleFactor[l]
        L.w = cv_round((float)width * p.inv_scale[l]);
        L.h = cv_round((float)height * p.inv_scale[l]);
        if (L.w < 1 || L.h < 1) { err = "pyramid level collapses to zero size"; return false; }
        L.pitch = (L.w + kPitchAlign - 1) / kPitchAlign * kPitchAlign;
        L.plane_bytes = (size_t)L.pitch * L.h;
        if (l > 0) {
            build_resize_taps(L.w, plan.lv[l - 1].w, true, L.xtap);
            build_resize_taps(L.h, plan.lv[l - 1].h, false, L.ytap);
            // compact form: needs c0 + c1 == 2048, second tap = first + 1 (or weight 0) and <= 6 source bytes between
            // the first taps of 4 neighbouring outputs (always true for scale factors in (1, 2))
            bool ok = true;
            const int sw = plan.lv[l - 1].w;
            for (int d = 0; d < L.w && ok; d++) {
                const ResizeTap &t = L.xtap[d];
                ok = (t.c0 + t.c1 == 2048) && t.c1 >= 0 && (t.ofs1 == t.ofs + 1 || t.c1 == 0) && t.ofs >= 0 && t.ofs < sw;
                if (ok && (d & 3) == 0) {
                    const int last = d + 3 < L.w ? d + 3 : L.w - 1;
                    ok = L.xtap[last].ofs - t.ofs <= 6 && L.xtap[last].ofs >= t.ofs;
                }
            }
            for (int d = 0; d < L.h && ok; d++) ok = L.ytap[d].c0 >= 0 && L.ytap[d].c1 >= 0;
            if (ok) {
                L.xpack.resize((size_t)(L.w + 3) / 4 * 4);
                for (size_t d = 0; d < L.xpack.size(); d++) {
                    const ResizeTap &t = L.xtap[d < (size_t)L.w ? d : (size_t)L.w - 1];
                    L.xpack[d] = ((uint32_t)t.ofs << 16) | (uint32_t)t.c1;
                }
            }
        }
        L.quota = p.quota[l];
        L.kp_size = (float)(int)((float)kPatchSize * p.scale[l]);

        // ---- ComputeKeyPointsOctTree cell grid ----
        const int maxBX = L.w - kEdge + 3, maxBY = L.h - kEdge + 3;
        L.reg_w = maxBX - kMinBorder; L.reg_h = maxBY - kMinBorder;
        L.first_cell = (int)plan.cells.size();
        L.cand_cap = 0;
        if (L.reg_w > 0 && L.reg_h > 0) {
            const float W = 35.f;
            const float fw = (float)L.reg_w, fh = (float)L.reg_h;
            L.ncols = (int)(fw / W); L.nrows = (int)(fh / W);
            if (L.ncols > 0 && L.nrows > 0) {
                L.wcell = (int)std::ceil(fw / (float)L.ncols);
                L.hcell = (int)std::ceil(fh / (float)L.nrows);
                if (L.wcell + 6 > 80 || L.hcell + 6 > 80) { err = "FAST cell larger than the kernel tile"; return false; }
                for (int i = 0; i < L.nrows; i++) {
                    const int iniY = kMinBorder + i * L.hcell;
                    int maxY = iniY + L.hcell + 6;
                    if (iniY >= maxBY - 3) continue;
                    if (maxY > maxBY) maxY = maxBY;
                    for (int j = 0; j < L.ncols; j++) {
                        const int iniX = kMinBorder + j * L.wcell;
                        int maxX = iniX + L.wcell + 6;
                        if (iniX >= maxBX - 6) continue;
                        if (maxX > maxBX) maxX = maxBX;
                        const int a = maxX - iniX - 6, b = maxY - iniY - 6;   // tested pixels
                        if (a <= 0 || b <= 0) continue;                          // cv::FAST finds nothing in such a ROI
                        CellRect c{};
                        c.level = (int16_t)l; c.x0 = (int16_t)iniX; c.y0 = (int16_t)iniY; c.x1 = (int16_t)maxX; c.y1 = (int16_t)maxY;
                        plan.cells.push_back(c);
                        L.cand_cap += ((a + 1) / 2) * ((b + 1) / 2);
                    }
                }
            }
        }
        L.ncells = (int)plan.cells.size() - L.first_cell;
        L.cand_cap = (L.cand_cap + 63) / 64 * 64 + 64;

        // ---- DistributeOctTree set-up ----
        L.n_ini = 0; L.depth0 = 0; L.nbins = 0;
        if (L.reg_w > 0 && L.reg_h > 0) {
            L.n_ini = (int)std::round((float)L.reg_w / (float)L.reg_h);
            if (L.n_ini < 1 || L.n_ini > kMaxRoots) { err = "unsupported aspect ratio (quadtree roots outside 1..16: a level wider than 16.5 x its height, or flatter than 0.5)"; return false; }
            const float hX = (float)L.reg_w / (float)L.n_ini;
            for (int i = 0; i < L.n_ini; i++) {
                L.root_ulx[i] = (int)(hX * (float)i);
                L.root_brx[i] = (int)(hX * (float)(i + 1));
            }
            int d0 = 6;
            while (d0 > 1 && (L.n_ini << (2 * d0)) > kMaxBins) d0--;
            L.depth0 = d0;
            L.nbins = L.n_ini << (2 * d0);
            L.xbin.assign(L.reg_w, 0); L.ybin.assign(L.reg_h, 0);
            L.xord.assign(L.reg_w, 0); L.yord.assign(L.reg_h, 0);
            for (int x = 0; x < L.reg_w; x++) {
                int r = (int)((float)x / hX);
                if (r >= L.n_ini) r = L.n_ini - 1;   // cannot happen for key coordinates (x <= reg_w - 4)
                int ul = L.root_ulx[r], br = L.root_brx[r];
                uint32_t path = 0;
                for (int d = 0; d < d0; d++) {
                    const int half = (int)std::ceil((float)(br - ul) / 2);
                    const int mx = ul + half;
                    if (x < mx) { path = path << 1; br = mx; } else { path = (path << 1) | 1u; ul = mx; }
                }
                L.xbin[x] = ((uint32_t)r << (2 * d0)) | spread_bits(path);
            }
            for (int y = 0; y < L.reg_h; y++) {
                int ul = 0, br = L.reg_h;
                uint32_t path = 0;
                for (int d = 0; d < d0; d++) {
                    const int half = (int)std::ceil((float)(br - ul) / 2);
                    const int my = ul + half;
                    if (y < my) { path = path << 1; br = my; } else { path = (path << 1) | 1u; ul = my; }
                }
                L.ybin[y] = spread_bits(path) << 1;
            }
            if (L.ncols > 0 && L.nrows > 0) {
                const uint32_t area = (uint32_t)L.wcell * (uint32_t)L.hcell;
                L.ord_cell_area = area; L.ord_ncols = (uint32_t)L.ncols;
                if ((uint64_t)area * (uint64_t)L.ncols * (uint64_t)L.nrows >= (1u << 24)) { err = "frame too large for the 24-bit order key"; return false; }
                for (int x = 0; x < L.reg_w; x++) {
                    int xi = x - 3; if (xi < 0) xi = 0;
                    int j = xi / L.wcell; if (j > L.ncols - 1) j = L.ncols - 1;
                    L.xord[x] = (uint32_t)j * area + (uint32_t)(xi - j * L.wcell);
                }
                for (int y = 0; y < L.reg_h; y++) {
                    int yi = y - 3; if (yi < 0) yi = 0;
                    int i = yi / L.hcell; if (i > L.nrows - 1) i = L.nrows - 1;
                    L.yord[y] = (uint32_t)i * (uint32_t)L.ncols * area + (uint32_t)(yi - i * L.hcell) * (uint32_t)L.wcell;
                }
            }
        }
        const int minout = 4 * (L.n_ini > 0 ? L.n_ini : 1);
        L.out_cap = (L.quota + 3 > minout ? L.quota + 3 : minout);
        plan.total_out_cap += L.out_cap;
    }
    // fused pyramid launches of up to four levels each (optional: an empty list means the per-level kernels run)
    for (int src = 0; src + 1 < p.nlevels; src += 4) {
        ConePlan c;
        if (!build_cone_plan(plan, src, std::min(src + 4, p.nlevels - 1), c)) { plan.cones.clear(); break; }
        plan.cones.push_back(std::move(c));
    }
    return true;
}

size_t Plan::algorithmic_bytes(int nkeypoints) const {
    size_t S = 0;
    for (const auto &L : lv) S += (size_t)L.w * L.h;
    const size_t P0 = (size_t)lv.front().w * lv.front().h, PL = (size_t)lv.back().w * lv.back().h;
    return 5 * S - P0 - PL + (size_t)1321 * (size_t)nkeypoints;
}

}  // namespace orbx

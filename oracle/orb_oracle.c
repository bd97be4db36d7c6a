/* ORACLE -- test infrastructure, NOT product code.
 *
 * Plain-C CPU restatement of the hot path named by BASELINE.json: ORB-SLAM3 ORBextractor::operator() and the
 * Hamming matchers, with the OpenCV primitives it delegates to written out as closed-form integer / fp32
 * arithmetic.  No OpenCV, no CUDA.  Compile with -ffp-contract=off.
 *
 * Where the algorithm lives: NOT under /root/reference.  The reference builds it from an un-vendored,
 * un-pinned clone (docker_container_setup.sh:42 `git clone https://github.com/devansh0703/ORB_SLAM3.git`,
 * default-branch HEAD) -- sources listed at slam_backends/orb_slam_3/CMakeLists.txt:52-53
 * (src/ORBextractor.cc, src/ORBmatcher.cc), reached from orbslam3_mono_networked.cc:511 (System ctor, ORB
 * parameters from the YAML literal at :193-206) and :594 (TrackMonocular).  The arithmetic inside cv::resize /
 * cv::FAST / cv::GaussianBlur / cv::fastAtan2 is OpenCV (Ubuntu 22.04 libopencv-dev 4.5.4 in the reference
 * image, docker_container_setup.sh:8-10).
 *
 * PIN STATUS: the reference holds no golden vector, known-answer test or fixture for this path
 * (send_slam/test/send_slam_test.exs:5-7 is its only test).  What pins this file instead:
 *   - resize / FAST+NMS / GaussianBlur / fastAtan2 / BFMatcher: bit-compared with the real OpenCV code via cv2
 *     4.13 (oracle/orb_cv2.py; tests/test_oracle_golden.py; fixtures under tests/golden/ made by tests/golden/make_golden.py).
 *   - cvtColor *2GRAY and cv::undistortPoints (the SURVEY.md 8f rows): bit-compared with cv2 4.13 in
 *     tests/test_oracle_golden.py, SHA pins of the verified outputs committed there.
 *   - steered BRIEF: bit-compared with cv2.ORB_create().compute() on the same keypoints/angles.
 *   - binary-PNM decode (the wire's frame encoding, orbslam3_mono_networked.cc:546): accept / reject verdict, geometry and
 *     Mat bytes compared with cv2.imdecode case by case (tests/golden/pnm_cases.npz, made by tests/golden/make_golden_pnm.py).
 *   - cell grid, DistributeOctTree, output ordering: "parity unpinned" -- restated from the published
 *     ORB-SLAM3 v1.0 algorithm; cross-checked only against the independent Python
 *     restatement in oracle/orb_cv2.py.  Known non-determinism in the reference itself: the final octree
 *     phase sorts (size, node pointer) pairs; equal sizes are ordered by heap address there and by creation
 *     sequence here.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef __AVX2__
#include <immintrin.h>
#endif

typedef unsigned char u8;

#define PATCH_SIZE 31
#define HALF_PATCH 15
#define EDGE_TH 19
#define MAX_LEVELS 16

static const int8_t k_pattern[1024] = {
#include "orb_pattern.inc"
};

/* FAST-9/16 ring, k = 0..15 */
static const int ring_dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int ring_dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

static inline int cv_round_f(float x) { return (int)lrintf(x); }
static inline int cv_round_d(double x) { return (int)lrint(x); }

/* ------------------------------------------------------------------------------------------------ */
/* handle                                                                                           */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} oracle_kp; /* layout of cv::KeyPoint */

typedef struct {
    int nfeatures, nlevels, ini_th, min_th;
    double scale_factor;
    float scale[MAX_LEVELS], inv_scale[MAX_LEVELS], sigma2[MAX_LEVELS], inv_sigma2[MAX_LEVELS];
    int quota[MAX_LEVELS];
    int umax[HALF_PATCH + 2];
    /* per-call workspace (kept for stage inspection) */
    int w[MAX_LEVELS], h[MAX_LEVELS];
    u8 *lvl[MAX_LEVELS];
    u8 *blr[MAX_LEVELS];
    float *cand[MAX_LEVELS];   /* x,y,resp triples relative to (16,16) */
    int ncand[MAX_LEVELS], cand_cap[MAX_LEVELS];
    int *sel[MAX_LEVELS];      /* indices into cand, list order */
    int nsel[MAX_LEVELS];
    float *ang[MAX_LEVELS];
    u8 *ldesc[MAX_LEVELS];
} oracle_t;

void *orb_oracle_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th) {
    if (nlevels < 1 || nlevels > MAX_LEVELS || nfeatures < 0 || !(scale_factor > 1.0f) || !(scale_factor < 2.0f))
        return NULL;
    oracle_t *o = (oracle_t *)calloc(1, sizeof(oracle_t));
    o->nfeatures = nfeatures; o->nlevels = nlevels; o->ini_th = ini_th; o->min_th = min_th;
    o->scale_factor = (double)scale_factor;
    o->scale[0] = 1.0f; o->sigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) {
        o->scale[i] = (float)((double)o->scale[i - 1] * o->scale_factor);
        o->sigma2[i] = o->scale[i] * o->scale[i];
    }
    for (int i = 0; i < nlevels; i++) { o->inv_scale[i] = 1.0f / o->scale[i]; o->inv_sigma2[i] = 1.0f / o->sigma2[i]; }
    float factor = (float)(1.0 / o->scale_factor);
    float nd = (float)nfeatures * (1.0f - factor) / (1.0f - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; l++) { o->quota[l] = cv_round_f(nd); sum += o->quota[l]; nd *= factor; }
    o->quota[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;
    int vmax = (int)floor(HALF_PATCH * sqrt(2.0) / 2 + 1), vmin = (int)ceil(HALF_PATCH * sqrt(2.0) / 2);
    for (int v = 0; v <= vmax; v++) o->umax[v] = cv_round_d(sqrt((double)(HALF_PATCH * HALF_PATCH - v * v)));
    for (int v = HALF_PATCH, v0 = 0; v >= vmin; --v) {
        while (o->umax[v0] == o->umax[v0 + 1]) ++v0;
        o->umax[v] = v0; ++v0;
    }
    return o;
}

static void free_ws(oracle_t *o) {
    for (int l = 0; l < MAX_LEVELS; l++) {
        free(o->lvl[l]); free(o->blr[l]); free(o->cand[l]); free(o->sel[l]); free(o->ang[l]); free(o->ldesc[l]);
        o->lvl[l] = o->blr[l] = o->ldesc[l] = NULL; o->cand[l] = o->ang[l] = NULL; o->sel[l] = NULL;
        o->cand_cap[l] = 0;
    }
}
void orb_oracle_destroy(void *h) { if (h) { free_ws((oracle_t *)h); free(h); } }

int orb_oracle_tables(void *h, float *scale, float *inv_scale, float *sigma2, float *inv_sigma2, int *quota, int *umax) {
    oracle_t *o = (oracle_t *)h;
    for (int l = 0; l < o->nlevels; l++) {
        if (scale) scale[l] = o->scale[l];
        if (inv_scale) inv_scale[l] = o->inv_scale[l];
        if (sigma2) sigma2[l] = o->sigma2[l];
        if (inv_sigma2) inv_sigma2[l] = o->inv_sigma2[l];
        if (quota) quota[l] = o->quota[l];
    }
    if (umax) for (int v = 0; v <= HALF_PATCH; v++) umax[v] = o->umax[v];
    return o->nlevels;
}

void orb_oracle_level_size(void *h, int w, int ht, int l, int *wl, int *hl) {
    oracle_t *o = (oracle_t *)h;
    *wl = cv_round_f((float)w * o->inv_scale[l]);
    *hl = cv_round_f((float)ht * o->inv_scale[l]);
}

/* ------------------------------------------------------------------------------------------------ */
/* cv::resize INTER_LINEAR 8UC1                                                     */
/* ------------------------------------------------------------------------------------------------ */
static inline short sat_s16(int v) { return (short)(v < -32768 ? -32768 : v > 32767 ? 32767 : v); }

/* per-axis offset + 11-bit coefficient pair; clamp_hi mirrors the x-axis handling (fx=0 at the far edge) */
static void resize_axis(int dst, int src, int is_x, int *ofs, short *c0, short *c1) {
    double scale = 1.0 / ((double)dst / (double)src);
    for (int d = 0; d < dst; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        if (is_x) {
            if (s < 0) { f = 0; s = 0; }
            if (s >= src - 1) { f = 0; s = src - 1; }
        }
        ofs[d] = s;
        c0[d] = sat_s16(cv_round_f((1.f - f) * 2048.f));
        c1[d] = sat_s16(cv_round_f(f * 2048.f));
    }
}

/* cv::cvtColor(src, COLOR_RGB2GRAY / BGR2GRAY / RGBA2GRAY / BGRA2GRAY), 8U: what UPSTREAM Tracking::GrabImageMonocular applies to a
 * colour frame before the Frame constructor (reached from slam_backends/orb_slam_3/orbslam3_mono_networked.cc:594; RGB vs BGR
 * = Camera.RGB, `rgb: 1` in send_slam/lib/send_slam/slam_handler.ex:222).  format: 1 RGB, 2 BGR, 3 RGBA, 4 BGRA.
 * shift 15: (R*9798 + G*19235 + B*3735 + 2^14) >> 15 -- bit-identical to cv2 4.13 (tests/test_oracle_golden.py);
 * shift 14: (R*4899 + G*9617 + B*1868 + 2^13) >> 14 -- the fixed-point form of older OpenCV builds. */
int orb_oracle_gray(const u8 *src, int w, int h, int stride, int format, int shift, u8 *dst, int dstride) {
    if (!src || !dst || w < 1 || h < 1 || format < 1 || format > 4 || (shift != 15 && shift != 14)) return -1;
    const int ch = format >= 3 ? 4 : 3, rgb = (format == 1 || format == 3);
    const unsigned r = shift == 15 ? 9798u : 4899u, g = shift == 15 ? 19235u : 9617u, b = shift == 15 ? 3735u : 1868u;
    const unsigned c0 = rgb ? r : b, c2 = rgb ? b : r, half = 1u << (shift - 1);
    for (int y = 0; y < h; y++) {
        const u8 *p = src + (size_t)y * stride;
        u8 *q = dst + (size_t)y * dstride;
        for (int x = 0; x < w; x++, p += ch) q[x] = (u8)((p[0] * c0 + p[1] * g + p[2] * c2 + half) >> shift);
    }
    return 0;
}

/* cv::imdecode(buf, IMREAD_UNCHANGED) for the only encoding the reference puts on the wire: binary PNM
 * (send_slam/lib/send_slam/slam_handler.ex:275-277 `Evision.imencode(".ppm", mat)`, field `encoding: "ppm"` :147; decoded at
 * slam_backends/orb_slam_3/orbslam3_mono_networked.cc:546).  Restates OpenCV's PxM reader (modules/imgcodecs/src/grfmt_pxm.cpp,
 * not in /root/reference; behaviour pinned against cv2 4.13 in tests/test_oracle_golden.py): 'P' + type digit, then width,
 * height, maxval as decimal numbers, each preceded by any run of white space and '#' comments (to end of line) and each
 * followed by exactly ONE consumed byte; binary samples are taken as they are (no rescaling for maxval < 255); a payload
 * shorter than width*height*channels is a decode failure.  P5 -> 1 channel; P6 -> 3 channels, which imdecode stores as BGR
 * (file order is RGB).  Returns 0 and (*w, *h, *ch, *offset = first payload byte), or -1 where cv::imdecode returns an empty
 * Mat, or -2 for what it decodes but the extractor cannot take (ASCII / bitmap variants and 16-bit samples: CV_16U fails
 * UPSTREAM's assert(image.type() == CV_8UC1)). */
static int pnm_number(const u8 *d, size_t n, size_t *pos, long long *out) {
    size_t p = *pos;
    int c;
#define PNM_GET() do { if (p >= n) return -1; c = d[p++]; } while (0)
#define PNM_SPACE(c) ((c) == ' ' || ((c) >= 9 && (c) <= 13))
    PNM_GET();
    while (c < '0' || c > '9') {
        if (c == '#') {
            do PNM_GET(); while (c != '\n' && c != '\r');
            PNM_GET();
        } else if (PNM_SPACE(c)) {
            while (PNM_SPACE(c)) PNM_GET();
        } else return -1;
    }
    long long v = 0;
    do {
        v = v * 10 + (c - '0');
        if (v > 2147483647LL) return -1;
        PNM_GET();                       /* the byte after the last digit is consumed, whatever it is */
    } while (c >= '0' && c <= '9');
#undef PNM_GET
#undef PNM_SPACE
    *pos = p; *out = v;
    return 0;
}

int orb_oracle_pnm_header(const u8 *data, size_t nbytes, int *w, int *h, int *ch, size_t *offset) {
    if (!data || nbytes < 2 || data[0] != 'P' || data[1] < '1' || data[1] > '6') return -1;
    const int type = data[1] - '0';
    size_t pos = 2;
    long long W, H, M = 1;
    if (pnm_number(data, nbytes, &pos, &W) || pnm_number(data, nbytes, &pos, &H)) return -1;
    if (type != 1 && type != 4 && pnm_number(data, nbytes, &pos, &M)) return -1;
    if (W <= 0 || H <= 0 || M <= 0 || M > 65535) return -1;
    if (type != 5 && type != 6) return -2;
    if (M > 255) return -2;
    const int c = type == 6 ? 3 : 1;
    if ((unsigned long long)W * (unsigned long long)H * c > nbytes - pos) return -1;
    *w = (int)W; *h = (int)H; *ch = c; *offset = pos;
    return 0;
}

/* The Mat imdecode returns: gray for P5, BGR for P6 (dst: h rows of w*ch bytes). */
int orb_oracle_pnm_decode(const u8 *data, size_t nbytes, u8 *dst) {
    int w, h, ch; size_t off;
    const int rc = orb_oracle_pnm_header(data, nbytes, &w, &h, &ch, &off);
    if (rc) return rc;
    const u8 *p = data + off;
    const size_t npx = (size_t)w * h;
    if (ch == 1) memcpy(dst, p, npx);
    else for (size_t i = 0; i < npx; i++) { dst[3 * i] = p[3 * i + 2]; dst[3 * i + 1] = p[3 * i + 1]; dst[3 * i + 2] = p[3 * i]; }
    return 0;
}

int orb_oracle_resize(const u8 *src, int sw, int sh, int sstride, u8 *dst, int dw, int dh, int dstride) {
    int *xo = (int *)malloc(sizeof(int) * dw), *yo = (int *)malloc(sizeof(int) * dh);
    short *xa = (short *)malloc(sizeof(short) * dw * 2), *ya = (short *)malloc(sizeof(short) * dh * 2);
    int *r0 = (int *)malloc(sizeof(int) * dw), *r1 = (int *)malloc(sizeof(int) * dw);
    resize_axis(dw, sw, 1, xo, xa, xa + dw);
    resize_axis(dh, sh, 0, yo, ya, ya + dh);
    for (int dy = 0; dy < dh; dy++) {
        int sy0 = yo[dy], sy1 = yo[dy] + 1;
        sy0 = sy0 < 0 ? 0 : (sy0 > sh - 1 ? sh - 1 : sy0);
        sy1 = sy1 < 0 ? 0 : (sy1 > sh - 1 ? sh - 1 : sy1);
        const u8 *S0 = src + (size_t)sy0 * sstride, *S1 = src + (size_t)sy1 * sstride;
        for (int dx = 0; dx < dw; dx++) {
            int sx = xo[dx], sx1 = sx + 1 < sw ? sx + 1 : sw - 1;
            r0[dx] = S0[sx] * xa[dx] + S0[sx1] * xa[dw + dx];
            r1[dx] = S1[sx] * xa[dx] + S1[sx1] * xa[dw + dx];
        }
        int b0 = ya[dy], b1 = ya[dh + dy];
        u8 *D = dst + (size_t)dy * dstride;
        for (int dx = 0; dx < dw; dx++) {
            int v = (((b0 * (r0[dx] >> 4)) >> 16) + ((b1 * (r1[dx] >> 4)) >> 16) + 2) >> 2;
            D[dx] = (u8)(v < 0 ? 0 : v > 255 ? 255 : v);
        }
    }
    free(xo); free(yo); free(xa); free(ya); free(r0); free(r1);
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* cv::GaussianBlur 7x7 sigma=2 REFLECT_101, 8U fixed point                         */
/* ------------------------------------------------------------------------------------------------ */
static inline int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) { if (i < 0) i = -i; else i = 2 * (n - 1) - i; }
    return i;
}

int orb_oracle_blur7(const u8 *src, int w, int h, int sstride, u8 *dst, int dstride) {
    static const int k[7] = {18, 34, 48, 56, 48, 34, 18};
    int *tmp = (int *)malloc(sizeof(int) * (size_t)w * h);
    for (int y = 0; y < h; y++) {
        const u8 *S = src + (size_t)y * sstride;
        int *T = tmp + (size_t)y * w;
        for (int x = 0; x < w; x++) {
            int acc = 0;
            if (x >= 3 && x < w - 3) for (int i = 0; i < 7; i++) acc += k[i] * S[x + i - 3];
            else for (int i = 0; i < 7; i++) acc += k[i] * S[reflect101(x + i - 3, w)];
            T[x] = acc;
        }
    }
    for (int y = 0; y < h; y++) {
        const int *R[7];
        for (int j = 0; j < 7; j++) R[j] = tmp + (size_t)reflect101(y + j - 3, h) * w;
        u8 *D = dst + (size_t)y * dstride;
        for (int x = 0; x < w; x++) {
            int acc = 32768;
            for (int j = 0; j < 7; j++) acc += k[j] * R[j][x];
            D[x] = (u8)(acc >> 16);
        }
    }
    free(tmp);
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* cv::FAST(roi, t, nonmax=true), TYPE_9_16                                         */
/* ------------------------------------------------------------------------------------------------ */
/* max(t, m) - 1 with m = best arc-of-9 contrast; >= t  <=>  corner at threshold t */
static int fast_score(const u8 *p, const int *off, int t) {
    int d[25], v = p[0];
    for (int k = 0; k < 16; k++) d[k] = v - p[off[k]];
    for (int k = 16; k < 25; k++) d[k] = d[k - 16];
    int a0 = t;
    for (int k = 0; k < 16; k += 2) {
        int a = d[k + 1] < d[k + 2] ? d[k + 1] : d[k + 2];
        if (d[k + 3] < a) a = d[k + 3];
        if (a <= a0) continue;
        for (int q = 4; q <= 8; q++) if (d[k + q] < a) a = d[k + q];
        int e = a < d[k] ? a : d[k];       if (e > a0) a0 = e;
        e = a < d[k + 9] ? a : d[k + 9];   if (e > a0) a0 = e;
    }
    int b0 = -a0;
    for (int k = 0; k < 16; k += 2) {
        int b = d[k + 1] > d[k + 2] ? d[k + 1] : d[k + 2];
        if (d[k + 3] > b) b = d[k + 3];
        if (d[k + 4] > b) b = d[k + 4];
        if (d[k + 5] > b) b = d[k + 5];
        if (b >= b0) continue;
        for (int q = 6; q <= 8; q++) if (d[k + q] > b) b = d[k + q];
        int e = b > d[k] ? b : d[k];       if (e < b0) b0 = e;
        e = b > d[k + 9] ? b : d[k + 9];   if (e < b0) b0 = e;
    }
    return -b0 - 1;
}

#ifdef __AVX2__
/* 32 pixels at once: m = max(v - min_arcs(max_arc r), max_arcs(min_arc r) - v, 0), the same value fast_score() derives
 * from d = v - r (min_k (v - r_k) = v - max_k r_k).  Arc extrema of the 16 arcs of 9: pairs Q[j] = (r[2j+1], r[2j+2]),
 * quads Q2[i] = Q[i] u Q[i+1]; the two arcs starting at 2i and 2i+1 share r[2i+1..2i+8] = Q2[i] u Q2[i+2]. */
static inline __m256i fast_m32(const u8 *p, const int *off) {
    __m256i r[16], qx[8], qn[8], q2x[8], q2n[8];
    for (int k = 0; k < 16; k++) r[k] = _mm256_loadu_si256((const __m256i *)(p + off[k]));
    const __m256i v = _mm256_loadu_si256((const __m256i *)p);
    for (int j = 0; j < 8; j++) {
        qx[j] = _mm256_max_epu8(r[2 * j + 1], r[(2 * j + 2) & 15]);
        qn[j] = _mm256_min_epu8(r[2 * j + 1], r[(2 * j + 2) & 15]);
    }
    for (int i = 0; i < 8; i++) {
        q2x[i] = _mm256_max_epu8(qx[i], qx[(i + 1) & 7]);
        q2n[i] = _mm256_min_epu8(qn[i], qn[(i + 1) & 7]);
    }
    __m256i min_arc_max = _mm256_set1_epi8((char)0xFF), max_arc_min = _mm256_setzero_si256();
    for (int i = 0; i < 8; i++) {
        const __m256i a = r[2 * i], b = r[(2 * i + 9) & 15];
        const __m256i fx = _mm256_max_epu8(_mm256_max_epu8(q2x[i], q2x[(i + 2) & 7]), _mm256_min_epu8(a, b));
        const __m256i fn = _mm256_min_epu8(_mm256_min_epu8(q2n[i], q2n[(i + 2) & 7]), _mm256_max_epu8(a, b));
        min_arc_max = _mm256_min_epu8(min_arc_max, fx);
        max_arc_min = _mm256_max_epu8(max_arc_min, fn);
    }
    return _mm256_max_epu8(_mm256_subs_epu8(v, min_arc_max), _mm256_subs_epu8(max_arc_min, v));
}
#endif

/* Runs FAST+NMS on the ROI [x0,x1)x[y0,y1) of img (img_w = full row width, bounds the vector loads); appends
 * (x+ox, y+oy, score) (ROI-relative coords plus offsets) row-major.  score buffer sc is (x1-x0)*(y1-y0) bytes of scratch.
 * Returns number appended.  score = cv's cornerScore = m - 1 for m > t. */
static int fast_roi(const u8 *img, int stride, int img_w, int x0, int y0, int x1, int y1, int t, u8 *sc,
                    float ox, float oy, float **out, int *n, int *cap) {
    int cw = x1 - x0, chh = y1 - y0, added = 0;
    if (cw < 7 || chh < 7) return 0;
    int off[16];
    for (int k = 0; k < 16; k++) off[k] = ring_dy[k] * stride + ring_dx[k];
    memset(sc, 0, (size_t)cw * chh);
    for (int y = 3; y < chh - 3; y++) {
        const u8 *row = img + (size_t)(y0 + y) * stride + x0;
        int x = 3;
#ifdef __AVX2__
        for (; x < cw - 3 && x0 + x + 3 + 32 <= img_w; x += 32) {
            u8 m[32];
            _mm256_storeu_si256((__m256i *)m, fast_m32(row + x, off));
            const int cnt = cw - 3 - x < 32 ? cw - 3 - x : 32;
            for (int i = 0; i < cnt; i++) if (m[i] > t) sc[y * cw + x + i] = (u8)(m[i] - 1);
        }
#endif
        for (; x < cw - 3; x++) {
            const u8 *p = row + x;
            int v = p[0], lo = v - t, hi = v + t;
            /* exact necessary condition: every opposite pair must hold one darker / one brighter pixel */
            int dk = 1, br = 1;
            for (int k = 0; k < 8 && (dk | br); k++) {
                int a = p[off[k]], b = p[off[k + 8]];
                dk &= (a < lo) | (b < lo);
                br &= (a > hi) | (b > hi);
            }
            if (!(dk | br)) continue;
            int s = fast_score(p, off, t);
            if (s >= t) sc[y * cw + x] = (u8)s;
        }
    }
    for (int y = 3; y < chh - 3; y++)
        for (int x = 3; x < cw - 3; x++) {
            int s = sc[y * cw + x];
            if (!s) continue;
            const u8 *q = sc + y * cw + x;
            if (s > q[-1] && s > q[1] && s > q[-cw - 1] && s > q[-cw] && s > q[-cw + 1] &&
                s > q[cw - 1] && s > q[cw] && s > q[cw + 1]) {
                if (*n >= *cap) { *cap = *cap ? *cap * 2 : 4096; *out = (float *)realloc(*out, sizeof(float) * 3 * (size_t)*cap); }
                float *o = *out + 3 * (size_t)*n;
                o[0] = (float)x + ox; o[1] = (float)y + oy; o[2] = (float)s;
                (*n)++; added++;
            }
        }
    return added;
}

/* Whole-image cv::FAST for pinning against cv2: out = (x,y,score) triples, returns count (<= cap) */
int orb_oracle_fast(const u8 *img, int w, int h, int stride, int t, float *out, int cap) {
    u8 *sc = (u8 *)malloc((size_t)w * h);
    float *buf = NULL; int n = 0, c = 0;
    fast_roi(img, stride, w, 0, 0, w, h, t, sc, 0.f, 0.f, &buf, &n, &c);
    int m = n < cap ? n : cap;
    if (m > 0) memcpy(out, buf, sizeof(float) * 3 * (size_t)m);
    free(buf); free(sc);
    return n;
}

/* ComputeKeyPointsOctTree cell loop.  Appends to *out candidates relative to (16,16). */
static void fast_cells(const u8 *img, int w, int h, int stride, int ini_th, int min_th,
                       float **out, int *n, int *cap) {
    const float W = 35;
    const int minBX = EDGE_TH - 3, minBY = minBX, maxBX = w - EDGE_TH + 3, maxBY = h - EDGE_TH + 3;
    const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
    const int nCols = (int)(width / W), nRows = (int)(height / W);
    *n = 0;
    if (nCols <= 0 || nRows <= 0) return;
    const int wCell = (int)ceilf(width / (float)nCols), hCell = (int)ceilf(height / (float)nRows);
    u8 *sc = (u8 *)malloc((size_t)(wCell + 6) * (hCell + 6));
    for (int i = 0; i < nRows; i++) {
        int iniY = minBY + i * hCell, maxY = iniY + hCell + 6;
        if (iniY >= maxBY - 3) continue;
        if (maxY > maxBY) maxY = maxBY;
        for (int j = 0; j < nCols; j++) {
            int iniX = minBX + j * wCell, maxX = iniX + wCell + 6;
            if (iniX >= maxBX - 6) continue;
            if (maxX > maxBX) maxX = maxBX;
            int got = fast_roi(img, stride, w, iniX, iniY, maxX, maxY, ini_th, sc, (float)(j * wCell), (float)(i * hCell), out, n, cap);
            if (!got) fast_roi(img, stride, w, iniX, iniY, maxX, maxY, min_th, sc, (float)(j * wCell), (float)(i * hCell), out, n, cap);
        }
    }
    free(sc);
}

int orb_oracle_fast_cells(const u8 *img, int w, int h, int stride, int ini_th, int min_th, float *out, int cap) {
    float *buf = NULL; int n = 0, c = 0;
    fast_cells(img, w, h, stride, ini_th, min_th, &buf, &n, &c);
    int m = n < cap ? n : cap;
    if (m > 0) memcpy(out, buf, sizeof(float) * 3 * (size_t)m);
    free(buf);
    return n;
}

/* ------------------------------------------------------------------------------------------------ */
/* DistributeOctTree -- std::list semantics on index-linked nodes                   */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
    int ulx, uly, brx, bry;
    int begin, end;    /* range in the key permutation */
    int nomore, prev, next, seq;
} onode;

typedef struct { int size, seq, node; } ocand;

static int ocand_cmp(const void *a, const void *b) {
    const ocand *x = (const ocand *)a, *y = (const ocand *)b;
    if (x->size != y->size) return x->size < y->size ? -1 : 1;
    return x->seq < y->seq ? -1 : (x->seq > y->seq);
}

typedef struct {
    onode *nd; int nn, ncap;
    int head, tail, count;
    int *perm, *tmp, *buf;
    const float *keys;
} otree;

static int ot_new(otree *t) {
    if (t->nn >= t->ncap) { t->ncap = t->ncap ? t->ncap * 2 : 1024; t->nd = (onode *)realloc(t->nd, sizeof(onode) * t->ncap); }
    t->nd[t->nn].seq = t->nn; t->nd[t->nn].prev = t->nd[t->nn].next = -1;
    return t->nn++;
}
static void ot_push_back(otree *t, int i) {
    t->nd[i].prev = t->tail; t->nd[i].next = -1;
    if (t->tail >= 0) t->nd[t->tail].next = i; else t->head = i;
    t->tail = i; t->count++;
}
static void ot_push_front(otree *t, int i) {
    t->nd[i].next = t->head; t->nd[i].prev = -1;
    if (t->head >= 0) t->nd[t->head].prev = i; else t->tail = i;
    t->head = i; t->count++;
}
static int ot_erase(otree *t, int i) { /* returns next */
    int p = t->nd[i].prev, n = t->nd[i].next;
    if (p >= 0) t->nd[p].next = n; else t->head = n;
    if (n >= 0) t->nd[n].prev = p; else t->tail = p;
    t->count--;
    return n;
}
/* DivideNode: stable 4-way partition of the parent's key range; children ids returned in c[4] (n1..n4) */
static void ot_divide(otree *t, int pi, int c[4]) {
    onode P = t->nd[pi];
    int halfX = (int)ceilf((float)(P.brx - P.ulx) / 2), halfY = (int)ceilf((float)(P.bry - P.uly) / 2);
    int mx = P.ulx + halfX, my = P.uly + halfY;
    int cnt[4] = {0, 0, 0, 0};
    for (int k = P.begin; k < P.end; k++) {
        const float *kp = t->keys + 3 * (size_t)t->perm[k];
        int q = (kp[0] < (float)mx) ? ((kp[1] < (float)my) ? 0 : 2) : ((kp[1] < (float)my) ? 1 : 3);
        t->tmp[k] = q; cnt[q]++;
    }
    int st[4], pos[4];
    st[0] = P.begin; st[1] = st[0] + cnt[0]; st[2] = st[1] + cnt[1]; st[3] = st[2] + cnt[2];
    memcpy(pos, st, sizeof(pos));
    for (int k = P.begin; k < P.end; k++) t->buf[pos[t->tmp[k]]++] = t->perm[k];
    memcpy(t->perm + P.begin, t->buf + P.begin, sizeof(int) * (size_t)(P.end - P.begin));
    const int bx[4][4] = {{P.ulx, P.uly, mx, my}, {mx, P.uly, P.brx, my}, {P.ulx, my, mx, P.bry}, {mx, my, P.brx, P.bry}};
    for (int q = 0; q < 4; q++) {
        int i = ot_new(t);
        onode *N = &t->nd[i];
        N->ulx = bx[q][0]; N->uly = bx[q][1]; N->brx = bx[q][2]; N->bry = bx[q][3];
        N->begin = st[q]; N->end = st[q] + cnt[q]; N->nomore = (cnt[q] == 1);
        c[q] = i;
    }
}

/* keys: n x (x,y,response) floats; out_idx: indices into keys in final list order; returns count */
int orb_oracle_octree(const float *keys, int n, int minX, int maxX, int minY, int maxY, int N, int *out_idx, int cap) {
    otree t; memset(&t, 0, sizeof(t));
    t.head = t.tail = -1; t.keys = keys;
    t.perm = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    t.tmp = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    t.buf = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    const int nIni = (int)roundf((float)(maxX - minX) / (float)(maxY - minY));
    if (nIni < 1) { free(t.perm); free(t.tmp); free(t.buf); return -1; }
    const float hX = (float)(maxX - minX) / (float)nIni;
    /* initial nodes: bucket keys by (int)(x / hX), stable */
    int *cnt = (int *)calloc((size_t)nIni + 1, sizeof(int));
    for (int k = 0; k < n; k++) { int b = (int)(keys[3 * (size_t)k] / hX); t.tmp[k] = b; cnt[b + 1]++; }
    for (int i = 0; i < nIni; i++) cnt[i + 1] += cnt[i];
    int *pos = (int *)malloc(sizeof(int) * (size_t)nIni);
    for (int i = 0; i < nIni; i++) pos[i] = cnt[i];
    for (int k = 0; k < n; k++) t.perm[pos[t.tmp[k]]++] = k;
    for (int i = 0; i < nIni; i++) {
        int id = ot_new(&t);
        onode *Nn = &t.nd[id];
        Nn->ulx = (int)(hX * (float)i); Nn->uly = 0;
        Nn->brx = (int)(hX * (float)(i + 1)); Nn->bry = maxY - minY;
        Nn->begin = cnt[i]; Nn->end = cnt[i + 1]; Nn->nomore = 0;
        ot_push_back(&t, id);
    }
    free(cnt); free(pos);
    for (int it = t.head; it >= 0;) {
        int sz = t.nd[it].end - t.nd[it].begin;
        if (sz == 1) { t.nd[it].nomore = 1; it = t.nd[it].next; }
        else if (sz == 0) it = ot_erase(&t, it);
        else it = t.nd[it].next;
    }
    ocand *cand = NULL, *prev = NULL; int ncand = 0, ccap = 0, nprev = 0, pcap = 0;
#define PUSH_CAND(id) do { if (ncand >= ccap) { ccap = ccap ? ccap * 2 : 256; cand = (ocand *)realloc(cand, sizeof(ocand) * ccap); } \
        cand[ncand].size = t.nd[id].end - t.nd[id].begin; cand[ncand].seq = t.nd[id].seq; cand[ncand].node = id; ncand++; } while (0)
    int finish = 0;
    while (!finish) {
        int prevSize = t.count, nToExpand = 0;
        ncand = 0;
        for (int it = t.head; it >= 0;) {
            if (t.nd[it].nomore) { it = t.nd[it].next; continue; }
            int c[4];
            ot_divide(&t, it, c);
            for (int q = 0; q < 4; q++) {
                int sz = t.nd[c[q]].end - t.nd[c[q]].begin;
                if (sz > 0) {
                    ot_push_front(&t, c[q]);
                    if (sz > 1) { nToExpand++; PUSH_CAND(c[q]); }
                }
            }
            it = ot_erase(&t, it);
        }
        if (t.count >= N || t.count == prevSize) finish = 1;
        else if (t.count + nToExpand * 3 > N) {
            while (!finish) {
                prevSize = t.count;
                if (ncand > pcap) { pcap = ncand; prev = (ocand *)realloc(prev, sizeof(ocand) * pcap); }
                memcpy(prev, cand, sizeof(ocand) * (size_t)ncand); nprev = ncand; ncand = 0;
                qsort(prev, (size_t)nprev, sizeof(ocand), ocand_cmp);
                for (int j = nprev - 1; j >= 0; j--) {
                    int c[4];
                    ot_divide(&t, prev[j].node, c);
                    for (int q = 0; q < 4; q++) {
                        int sz = t.nd[c[q]].end - t.nd[c[q]].begin;
                        if (sz > 0) {
                            ot_push_front(&t, c[q]);
                            if (sz > 1) PUSH_CAND(c[q]);
                        }
                    }
                    ot_erase(&t, prev[j].node);
                    if (t.count >= N) break;
                }
                if (t.count >= N || t.count == prevSize) finish = 1;
            }
        }
    }
    int m = 0;
    for (int it = t.head; it >= 0; it = t.nd[it].next) {
        int best = t.perm[t.nd[it].begin];
        float br = keys[3 * (size_t)best + 2];
        for (int k = t.nd[it].begin + 1; k < t.nd[it].end; k++) {
            int id = t.perm[k];
            if (keys[3 * (size_t)id + 2] > br) { best = id; br = keys[3 * (size_t)id + 2]; }
        }
        if (m < cap) out_idx[m] = best;
        m++;
    }
    free(cand); free(prev); free(t.nd); free(t.perm); free(t.tmp); free(t.buf);
    return m;
}

/* ------------------------------------------------------------------------------------------------ */
/* IC_Angle + cv::fastAtan2                                                         */
/* ------------------------------------------------------------------------------------------------ */
float orb_oracle_fast_atan2(float y, float x) {
    const float s = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s;
    const float p5 = 0.1555786518463281f * s, p7 = -0.04432655554792128f * s;
    float ax = fabsf(x), ay = fabsf(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

void orb_oracle_ic_moments(const u8 *img, int stride, int cx, int cy, const int *umax, int *m01_out, int *m10_out) {
    const u8 *c = img + (size_t)cy * stride + cx;
    int m01 = 0, m10 = 0;
    for (int u = -HALF_PATCH; u <= HALF_PATCH; u++) m10 += u * c[u];
    for (int v = 1; v <= HALF_PATCH; v++) {
        int vs = 0, d = umax[v];
        for (int u = -d; u <= d; u++) {
            int p = c[u + v * stride], m = c[u - v * stride];
            vs += p - m; m10 += u * (p + m);
        }
        m01 += v * vs;
    }
    *m01_out = m01; *m10_out = m10;
}

float orb_oracle_ic_angle(void *h, const u8 *img, int stride, float x, float y) {
    oracle_t *o = (oracle_t *)h;
    int m01, m10;
    orb_oracle_ic_moments(img, stride, cv_round_f(x), cv_round_f(y), o->umax, &m01, &m10);
    return orb_oracle_fast_atan2((float)m01, (float)m10);
}

/* ------------------------------------------------------------------------------------------------ */
/* computeOrbDescriptor                                                             */
/* ------------------------------------------------------------------------------------------------ */
/* Trig rule.  The reference calls cosf/sinf (std::cos/std::sin on a float).  glibc's cosf is NOT correctly rounded
 * (about 1.3 % of arguments are 1 ulp off here) and x86-64 glibc selects an FMA or SSE2 variant per CPU, so the
 * reference's own bits depend on the host.  Canonical oracle rule (mode 1, default): the correctly rounded fp32 of the
 * fp64 cos/sin -- what cv2.ORB-free Python (oracle/orb_cv2.py) and the CUDA kernel compute.  Mode 0 = this host's
 * libm cosf/sinf, kept to measure how many descriptor bits the choice moves (tests/test_oracle_vs_cv2.py). */
static int g_trig_mode = 1;
void orb_oracle_set_trig_mode(int mode) { g_trig_mode = mode; }

void orb_oracle_brief(const u8 *blurred, int stride, float x, float y, float angle_deg, u8 *desc) {
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    float ang = angle_deg * factorPI;
    float a, b;
    if (g_trig_mode == 0) { a = cosf(ang); b = sinf(ang); }
    else { a = (float)cos((double)ang); b = (float)sin((double)ang); }
    const u8 *c = blurred + (size_t)cv_round_f(y) * stride + cv_round_f(x);
    for (int i = 0; i < 32; i++) {
        int val = 0;
        for (int j = 0; j < 8; j++) {
            const int8_t *p = k_pattern + (i * 8 + j) * 4;
            float x0 = p[0], y0 = p[1], x1 = p[2], y1 = p[3];
            int t0 = c[cv_round_f(x0 * b + y0 * a) * stride + cv_round_f(x0 * a - y0 * b)];
            int t1 = c[cv_round_f(x1 * b + y1 * a) * stride + cv_round_f(x1 * a - y1 * b)];
            val |= (t0 < t1) << j;
        }
        desc[i] = (u8)val;
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* ORBextractor::operator()                                                         */
/* ------------------------------------------------------------------------------------------------ */
int orb_oracle_extract(void *h, const u8 *img, int w, int ht, int stride, int lap0, int lap1,
                       oracle_kp *kp_out, u8 *desc_out, int cap, int *n_out, int *mono_index_out) {
    oracle_t *o = (oracle_t *)h;
    if (!img || w <= 0 || ht <= 0) { *n_out = 0; *mono_index_out = -1; return -1; }
    /* ComputePyramid */
    for (int l = 0; l < o->nlevels; l++) {
        int wl, hl; orb_oracle_level_size(o, w, ht, l, &wl, &hl);
        o->w[l] = wl; o->h[l] = hl;
        if (wl < 1 || hl < 1) { *n_out = 0; *mono_index_out = -1; return -2; }
        o->lvl[l] = (u8 *)realloc(o->lvl[l], (size_t)wl * hl);
        if (l == 0) for (int y = 0; y < ht; y++) memcpy(o->lvl[0] + (size_t)y * w, img + (size_t)y * stride, (size_t)w);
        else orb_oracle_resize(o->lvl[l - 1], o->w[l - 1], o->h[l - 1], o->w[l - 1], o->lvl[l], wl, hl, wl);
    }
    /* ComputeKeyPointsOctTree */
    int total = 0;
    for (int l = 0; l < o->nlevels; l++) {
        int wl = o->w[l], hl = o->h[l];
        fast_cells(o->lvl[l], wl, hl, wl, o->ini_th, o->min_th, &o->cand[l], &o->ncand[l], &o->cand_cap[l]);
        int n = o->ncand[l];
        o->sel[l] = (int *)realloc(o->sel[l], sizeof(int) * (size_t)(n > 0 ? n : 1));
        int minB = EDGE_TH - 3;
        int m = 0;
        if (wl - 2 * minB > 0 && hl - 2 * minB > 0)
            m = orb_oracle_octree(o->cand[l], n, minB, wl - EDGE_TH + 3, minB, hl - EDGE_TH + 3, o->quota[l], o->sel[l], n);
        if (m < 0) m = 0;
        o->nsel[l] = m;
        o->ang[l] = (float *)realloc(o->ang[l], sizeof(float) * (size_t)(m > 0 ? m : 1));
        for (int i = 0; i < m; i++) {
            const float *c = o->cand[l] + 3 * (size_t)o->sel[l][i];
            o->ang[l][i] = orb_oracle_ic_angle(o, o->lvl[l], wl, c[0] + (float)minB, c[1] + (float)minB);
        }
        total += m;
    }
    *n_out = total;
    int mono = 0, stereo = total - 1;
    for (int l = 0; l < o->nlevels; l++) {
        int m = o->nsel[l], wl = o->w[l], hl = o->h[l];
        if (m == 0) continue;
        o->blr[l] = (u8 *)realloc(o->blr[l], (size_t)wl * hl);
        orb_oracle_blur7(o->lvl[l], wl, hl, wl, o->blr[l], wl);
        o->ldesc[l] = (u8 *)realloc(o->ldesc[l], (size_t)m * 32);
        float scale = o->scale[l];
        int patch = (int)((float)PATCH_SIZE * scale);
        for (int i = 0; i < m; i++) {
            const float *c = o->cand[l] + 3 * (size_t)o->sel[l][i];
            float x = c[0] + (float)(EDGE_TH - 3), y = c[1] + (float)(EDGE_TH - 3);
            u8 *d = o->ldesc[l] + (size_t)i * 32;
            orb_oracle_brief(o->blr[l], wl, x, y, o->ang[l][i], d);
            if (l != 0) { x *= scale; y *= scale; }
            int slot;
            if (x >= (float)lap0 && x <= (float)lap1) slot = stereo--; else slot = mono++;
            if (slot < cap) {
                oracle_kp *k = kp_out + slot;
                k->x = x; k->y = y; k->size = (float)patch; k->angle = o->ang[l][i]; k->response = c[2];
                k->octave = l; k->class_id = -1;
                memcpy(desc_out + (size_t)slot * 32, d, 32);
            }
        }
    }
    *mono_index_out = mono;
    return 0;
}

/* stage accessors (valid after orb_oracle_extract on the same handle) */
int orb_oracle_stage_level(void *h, int l, int *w, int *ht, const u8 **pix, const u8 **blurred) {
    oracle_t *o = (oracle_t *)h;
    *w = o->w[l]; *ht = o->h[l]; *pix = o->lvl[l]; *blurred = o->nsel[l] ? o->blr[l] : NULL;
    return 0;
}
int orb_oracle_stage_keys(void *h, int l, int *ncand, const float **cand, int *nsel, const int **sel, const float **ang, const u8 **desc) {
    oracle_t *o = (oracle_t *)h;
    *ncand = o->ncand[l]; *cand = o->cand[l]; *nsel = o->nsel[l]; *sel = o->sel[l]; *ang = o->ang[l];
    *desc = o->nsel[l] ? o->ldesc[l] : NULL;
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* multi-threaded batch (CPU baseline: one extractor instance per thread)           */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
    int nfeatures, nlevels, ini_th, min_th; float scale;
    const u8 *frames; int w, h, stride; size_t frame_bytes; int nframes;
    int lap0, lap1;
    oracle_kp *kps; u8 *desc; int cap; int *n_out; int *mono_out;
    volatile int *next;
} batch_job;

static void *batch_worker(void *arg) {
    batch_job *j = (batch_job *)arg;
    void *h = orb_oracle_create(j->nfeatures, j->scale, j->nlevels, j->ini_th, j->min_th);
    for (;;) {
        int f = __sync_fetch_and_add(j->next, 1);
        if (f >= j->nframes) break;
        orb_oracle_extract(h, j->frames + (size_t)f * j->frame_bytes, j->w, j->h, j->stride, j->lap0, j->lap1,
                           j->kps + (size_t)f * j->cap, j->desc + (size_t)f * j->cap * 32, j->cap,
                           j->n_out + f, j->mono_out + f);
    }
    orb_oracle_destroy(h);
    return NULL;
}

int orb_oracle_extract_batch(int nfeatures, float scale, int nlevels, int ini_th, int min_th,
                             const u8 *frames, int w, int h, int stride, size_t frame_bytes, int nframes,
                             int lap0, int lap1, oracle_kp *kps, u8 *desc, int cap, int *n_out, int *mono_out,
                             int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 1024) nthreads = 1024;
    volatile int next = 0;
    batch_job job = {nfeatures, nlevels, ini_th, min_th, scale, frames, w, h, stride, frame_bytes, nframes,
                     lap0, lap1, kps, desc, cap, n_out, mono_out, &next};
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    for (int i = 0; i < nthreads; i++) pthread_create(&th[i], NULL, batch_worker, &job);
    for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
    free(th);
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* ORBmatcher::DescriptorDistance + brute-force kNN k=2 + windowed search           */
/* ------------------------------------------------------------------------------------------------ */
int orb_oracle_distance(const u8 *a, const u8 *b) {
    const uint32_t *pa = (const uint32_t *)a, *pb = (const uint32_t *)b;
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        uint32_t v = pa[i] ^ pb[i];
        v = v - ((v >> 1) & 0x55555555u);
        v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u);
        dist += (int)((((v + (v >> 4)) & 0x0F0F0F0Fu) * 0x01010101u) >> 24);
    }
    return dist;
}

static inline int hamming256(const uint64_t *a, const uint64_t *b) {
    return __builtin_popcountll(a[0] ^ b[0]) + __builtin_popcountll(a[1] ^ b[1]) +
           __builtin_popcountll(a[2] ^ b[2]) + __builtin_popcountll(a[3] ^ b[3]);
}

typedef struct {
    const u8 *q; int nq; const u8 *db; long ndb; int32_t *idx; int32_t *dist; int q0, q1;
} knn_job;

#define QB 8
static void *knn_worker(void *arg) {
    knn_job *j = (knn_job *)arg;
    for (int qb = j->q0; qb < j->q1; qb += QB) {
        int nb = j->q1 - qb < QB ? j->q1 - qb : QB;
        int d1[QB], d2[QB]; int32_t i1[QB], i2[QB];
        uint64_t qq[QB][4];
        for (int k = 0; k < nb; k++) { d1[k] = d2[k] = 1 << 30; i1[k] = i2[k] = -1; memcpy(qq[k], j->q + (size_t)(qb + k) * 32, 32); }
        for (long r = 0; r < j->ndb; r++) {
            uint64_t row[4]; memcpy(row, j->db + (size_t)r * 32, 32);
            for (int k = 0; k < nb; k++) {
                int d = hamming256(qq[k], row);
                if (d < d2[k]) {               /* strict: ties keep the lower train index (BFMatcher) */
                    if (d < d1[k]) { d2[k] = d1[k]; i2[k] = i1[k]; d1[k] = d; i1[k] = (int32_t)r; }
                    else { d2[k] = d; i2[k] = (int32_t)r; }
                }
            }
        }
        for (int k = 0; k < nb; k++) {
            j->idx[2 * (qb + k)] = i1[k]; j->idx[2 * (qb + k) + 1] = i2[k];
            j->dist[2 * (qb + k)] = i1[k] < 0 ? -1 : d1[k]; j->dist[2 * (qb + k) + 1] = i2[k] < 0 ? -1 : d2[k];
        }
    }
    return NULL;
}

/* cv::BFMatcher(NORM_HAMMING).knnMatch(q, db, k=2): idx/dist are nq x 2 (missing -> -1) */
int orb_oracle_knn2(const u8 *q, int nq, const u8 *db, long ndb, int32_t *idx, int32_t *dist, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > nq) nthreads = nq > 0 ? nq : 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    knn_job *jobs = (knn_job *)malloc(sizeof(knn_job) * nthreads);
    int per = ((nq + nthreads - 1) / nthreads + QB - 1) / QB * QB;
    int used = 0;
    for (int i = 0; i < nthreads; i++) {
        int q0 = i * per, q1 = q0 + per < nq ? q0 + per : nq;
        if (q0 >= nq) break;
        jobs[i] = (knn_job){q, nq, db, ndb, idx, dist, q0, q1};
        pthread_create(&th[i], NULL, knn_worker, &jobs[i]); used++;
    }
    for (int i = 0; i < used; i++) pthread_join(th[i], NULL);
    free(th); free(jobs);
    return 0;
}

/* ---- Frame post-extraction steps (UPSTREAM ORB-SLAM3 src/Frame.cc, not under /root/reference) ------------------
 * cv::undistortPoints(src, dst, K, D, noArray(), K) as Frame::UndistortKeyPoints calls it: cvUndistortPointsInternal with
 * criteria = (MAX_ITER, 5), no tilt, R = I, P = K; all arithmetic in double, results stored as float.  cam = fx fy cx cy k1 k2 p1
 * p2 k3 (the calibration values of slam_backends/orb_slam_3/orbslam3_mono_networked.cc:173-176).  Pinned against
 * cv2.undistortPoints in tests/test_oracle_golden.py (agreement to float rounding; OpenCV's own build may contract FMAs). */
static void undistort_point(float u_in, float v_in, const float *cam, float *xo, float *yo) {
    const double fx = cam[0], fy = cam[1], cx = cam[2], cy = cam[3];
    const double k0 = cam[4], k1 = cam[5], k2 = cam[6], k3 = cam[7], k4 = cam[8];
    const double ifx = 1.0 / fx, ify = 1.0 / fy;
    const double u = u_in, v = v_in;
    double x = (u - cx) * ifx, y = (v - cy) * ify;
    const double x0 = x, y0 = y;
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
    *xo = (float)(xx * ww); *yo = (float)(yy * ww);
}

int orb_oracle_undistort_points(const float *xy, int n, const float *cam, float *out) {
    for (int i = 0; i < n; i++) undistort_point(xy[2 * i], xy[2 * i + 1], cam, &out[2 * i], &out[2 * i + 1]);
    return 0;
}

/* Frame::ComputeImageBounds: mnMinX, mnMinY, mnMaxX, mnMaxY */
int orb_oracle_image_bounds(const float *cam, int w, int h, float *b) {
    if (cam[4] == 0.f) { b[0] = 0.f; b[1] = 0.f; b[2] = (float)w; b[3] = (float)h; return 0; }
    const float c[8] = {0.f, 0.f, (float)w, 0.f, 0.f, (float)h, (float)w, (float)h};
    float un[8];
    orb_oracle_undistort_points(c, 4, cam, un);
    b[0] = un[0] < un[4] ? un[0] : un[4]; b[2] = un[2] > un[6] ? un[2] : un[6];
    b[1] = un[1] < un[3] ? un[1] : un[3]; b[3] = un[5] > un[7] ? un[5] : un[7];
    return 0;
}

/* Frame::UndistortKeyPoints + Frame::AssignFeaturesToGrid (PosInGrid, 64 x 48): kp_un = keypoints with undistorted pt,
 * mGrid as CSR: cell = posX * 48 + posY, cell_start[3073], items in push_back order. */
int orb_oracle_frame_grid(const oracle_kp *kp, int n, const float *cam, const float *bounds, oracle_kp *kp_un,
                          int32_t *cell_start, int32_t *cell_items) {
    const float minX = bounds[0], minY = bounds[1], maxX = bounds[2], maxY = bounds[3];
    const float invW = 64.f / (maxX - minX), invH = 48.f / (maxY - minY);
    int *cell_of = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    memset(cell_start, 0, sizeof(int32_t) * (64 * 48 + 1));
    for (int i = 0; i < n; i++) {
        kp_un[i] = kp[i];
        if (cam[4] != 0.f) undistort_point(kp[i].x, kp[i].y, cam, &kp_un[i].x, &kp_un[i].y);
        const int px = (int)roundf((kp_un[i].x - minX) * invW), py = (int)roundf((kp_un[i].y - minY) * invH);
        if (px < 0 || px >= 64 || py < 0 || py >= 48) { cell_of[i] = -1; continue; }
        cell_of[i] = px * 48 + py; cell_start[cell_of[i] + 1]++;
    }
    for (int c = 0; c < 64 * 48; c++) cell_start[c + 1] += cell_start[c];
    int *fill = (int *)malloc(sizeof(int) * 64 * 48);
    for (int c = 0; c < 64 * 48; c++) fill[c] = cell_start[c];
    for (int i = 0; i < n; i++) if (cell_of[i] >= 0) cell_items[fill[cell_of[i]]++] = i;
    free(cell_of); free(fill);
    return 0;
}

/* Frame::PosInGrid + Frame::GetFeaturesInArea + best / second-best DescriptorDistance.
 * train keypoints: oracle_kp records (x, y, octave used).  query q: desc + (u, v, r, minLevel, maxLevel).
 * bounds = {minX, minY, maxX, maxY} of the (undistorted) image.  Candidate visiting order = grid cell
 * (ix outer, iy inner), then train index; strict '<' so the first visited wins ties.
 * out: best_idx, best_dist, second_idx, second_dist per query (-1 / 256 when missing). */
#define GRID_COLS 64
#define GRID_ROWS 48
int orb_oracle_match_windowed(const u8 *qdesc, const float *quvr, const int32_t *qlevels, int nq,
                              const oracle_kp *tkp, const u8 *tdesc, int nt, const float *bounds,
                              int32_t *best_idx, int32_t *best_dist, int32_t *second_idx, int32_t *second_dist) {
    const float minX = bounds[0], minY = bounds[1], maxX = bounds[2], maxY = bounds[3];
    const float invW = (float)GRID_COLS / (maxX - minX), invH = (float)GRID_ROWS / (maxY - minY);
    int *cell_cnt = (int *)calloc(GRID_COLS * GRID_ROWS + 1, sizeof(int));
    int *cell_of = (int *)malloc(sizeof(int) * (size_t)(nt > 0 ? nt : 1));
    for (int i = 0; i < nt; i++) {
        int px = (int)roundf((tkp[i].x - minX) * invW), py = (int)roundf((tkp[i].y - minY) * invH);
        if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) { cell_of[i] = -1; continue; }
        cell_of[i] = px * GRID_ROWS + py; cell_cnt[cell_of[i] + 1]++;
    }
    for (int c = 0; c < GRID_COLS * GRID_ROWS; c++) cell_cnt[c + 1] += cell_cnt[c];
    int *fill = (int *)malloc(sizeof(int) * GRID_COLS * GRID_ROWS);
    memcpy(fill, cell_cnt, sizeof(int) * GRID_COLS * GRID_ROWS);
    int *items = (int *)malloc(sizeof(int) * (size_t)(nt > 0 ? nt : 1));
    for (int i = 0; i < nt; i++) if (cell_of[i] >= 0) items[fill[cell_of[i]]++] = i;
    for (int qi = 0; qi < nq; qi++) {
        float x = quvr[3 * qi], y = quvr[3 * qi + 1], r = quvr[3 * qi + 2];
        int minLevel = qlevels[2 * qi], maxLevel = qlevels[2 * qi + 1];
        int b1 = 256, b2 = 256, i1 = -1, i2 = -1;
        int cx0 = (int)floorf((x - minX - r) * invW); if (cx0 < 0) cx0 = 0;
        int cx1 = (int)ceilf((x - minX + r) * invW); if (cx1 > GRID_COLS - 1) cx1 = GRID_COLS - 1;
        int cy0 = (int)floorf((y - minY - r) * invH); if (cy0 < 0) cy0 = 0;
        int cy1 = (int)ceilf((y - minY + r) * invH); if (cy1 > GRID_ROWS - 1) cy1 = GRID_ROWS - 1;
        if (cx0 < GRID_COLS && cx1 >= 0 && cy0 < GRID_ROWS && cy1 >= 0) {
            int check = (minLevel > 0) || (maxLevel >= 0);
            for (int ix = cx0; ix <= cx1; ix++)
                for (int iy = cy0; iy <= cy1; iy++)
                    for (int k = cell_cnt[ix * GRID_ROWS + iy]; k < cell_cnt[ix * GRID_ROWS + iy + 1]; k++) {
                        int t = items[k];
                        if (check) {
                            if (tkp[t].octave < minLevel) continue;
                            if (maxLevel >= 0 && tkp[t].octave > maxLevel) continue;
                        }
                        if (!(fabsf(tkp[t].x - x) < r && fabsf(tkp[t].y - y) < r)) continue;
                        int d = orb_oracle_distance(qdesc + (size_t)qi * 32, tdesc + (size_t)t * 32);
                        if (d < b1) { b2 = b1; i2 = i1; b1 = d; i1 = t; }
                        else if (d < b2) { b2 = d; i2 = t; }
                    }
        }
        best_idx[qi] = i1; best_dist[qi] = b1; second_idx[qi] = i2; second_dist[qi] = b2;
    }
    free(cell_cnt); free(cell_of); free(fill); free(items);
    return 0;
}

/* ---- DBoW2 vocabulary-tree descent ----------------------------------------------------------------------------------
 * TemplatedVocabulary<ORB>::transform(feature, word_id, weight, nid, levelsup) of the DBoW2 library the reference build links (third
 * party, not under /root/reference; restated from the published source -- "parity unpinned").  Tree as arrays, node 0 = root, parents
 * before children; children of a node in id order; word ids = leaves in id order; L = depth of the deepest node. */
int orb_oracle_bow_transform(const int32_t *parent, const u8 *ndesc, const float *weight, int n_nodes, const u8 *feat, int n, int levelsup,
                             int32_t *word_id, float *word_weight, int32_t *node_id) {
    int *start = (int *)calloc((size_t)n_nodes + 1, sizeof(int)), *child = (int *)malloc(sizeof(int) * (size_t)n_nodes);
    int *word = (int *)malloc(sizeof(int) * (size_t)n_nodes), *depth = (int *)calloc((size_t)n_nodes, sizeof(int));
    int L = 0, nwords = 0;
    for (int i = 1; i < n_nodes; i++) start[parent[i] + 1]++;
    for (int i = 0; i < n_nodes; i++) start[i + 1] += start[i];
    int *fill = (int *)malloc(sizeof(int) * (size_t)n_nodes);
    memcpy(fill, start, sizeof(int) * (size_t)n_nodes);
    for (int i = 1; i < n_nodes; i++) { child[fill[parent[i]]++] = i; depth[i] = depth[parent[i]] + 1; if (depth[i] > L) L = depth[i]; }
    for (int i = 0; i < n_nodes; i++) word[i] = (start[i + 1] == start[i]) ? nwords++ : -1;
    const int nid_level = L - levelsup;
    for (int f = 0; f < n; f++) {
        int final_id = 0, current_level = 0, nid = 0;
        while (start[final_id + 1] > start[final_id]) {
            ++current_level;
            const int lo = start[final_id], hi = start[final_id + 1];
            int best = child[lo], best_d = orb_oracle_distance(feat + (size_t)f * 32, ndesc + (size_t)child[lo] * 32);
            for (int c = lo + 1; c < hi; c++) {
                const int d = orb_oracle_distance(feat + (size_t)f * 32, ndesc + (size_t)child[c] * 32);
                if (d < best_d) { best_d = d; best = child[c]; }
            }
            final_id = best;
            if (current_level == nid_level) nid = final_id;
        }
        word_id[f] = word[final_id]; word_weight[f] = weight[final_id]; node_id[f] = nid;
    }
    free(start); free(child); free(word); free(depth); free(fill);
    return L;
}


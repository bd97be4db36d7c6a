"""ORACLE (test infrastructure, not product code) -- cv2-backed restatement of ORB-SLAM3's ORBextractor.

PARITY PIN STATUS
  * The OpenCV primitives the reference's hot path delegates to (cv::resize INTER_LINEAR, cv::FAST with
    NMS, cv::GaussianBlur 7x7 sigma 2, cv::fastAtan2, BFMatcher NORM_HAMMING) are called here as the REAL
    OpenCV code through cv2 (4.13 in this image).  They pin oracle/orb_oracle.c (closed-form C restatement).
  * The ORB-SLAM3 control logic around them (cell grid, DistributeOctTree, IC_Angle, steered BRIEF, output
    ordering) is NOT present under /root/reference (cloned un-pinned at docker build time,
    docker_container_setup.sh:42; listed for build at slam_backends/orb_slam_3/CMakeLists.txt:52-53) and the
    reference holds no test or golden vector for it (send_slam/test/send_slam_test.exs:5-7 only).
    => for that logic: "parity unpinned".  It is restated from the published ORB-SLAM3 v1.0 algorithm as
    summarised in SURVEY.md Appendix C; call sites anchoring it: orbslam3_mono_networked.cc:511 (System ctor
    builds the extractors from the YAML at :193-206) and :594 (TrackMonocular -> Frame -> operator()).

Only tests/, tests/golden/make_golden.py and bench.py's cpu_baseline leg may import this module.
cv2 is needed only here (fixture generation / oracle cross-check), never by the product path.
"""
from __future__ import annotations

import math
import os

import numpy as np

try:  # cv2 is optional at import time so that `-m "not gpu"` collection never hard-fails without it
    import cv2
except Exception:  # pragma: no cover
    cv2 = None

PATCH_SIZE = 31
HALF_PATCH_SIZE = 15
EDGE_THRESHOLD = 19
F32 = np.float32


def load_pattern() -> np.ndarray:
    """256 x 4 int32 (x0,y0,x1,y1) from oracle/orb_pattern.inc."""
    here = os.path.dirname(os.path.abspath(__file__))
    txt = open(os.path.join(here, "orb_pattern.inc")).read()
    txt = txt[txt.index("*/") + 2:]
    vals = [int(v) for v in txt.replace("\n", " ").split(",") if v.strip()]
    assert len(vals) == 1024
    return np.array(vals, dtype=np.int32).reshape(256, 4)


def cv_round(x) -> int:
    """cvRound = lrint (round-half-even) of a float/double."""
    return int(np.rint(np.float64(x)))


class Params:
    """Constructor tables of ORBextractor (SURVEY.md C.1 'Constructor')."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.nfeatures, self.nlevels, self.ini_th, self.min_th = nfeatures, nlevels, ini_th, min_th
        sf = float(F32(scale_factor))  # float arg stored in a double member
        self.scale_factor = sf
        s = [F32(1.0)]
        for i in range(1, nlevels):
            s.append(F32(float(s[i - 1]) * sf))
        self.scale = np.array(s, dtype=F32)
        self.sigma2 = (self.scale * self.scale).astype(F32)
        self.inv_scale = (F32(1.0) / self.scale).astype(F32)
        self.inv_sigma2 = (F32(1.0) / self.sigma2).astype(F32)
        factor = F32(1.0 / sf)
        denom = F32(1) - F32(math.pow(float(factor), float(nlevels)))
        nd = F32(F32(nfeatures) * (F32(1) - factor)) / denom
        nd = F32(nd)
        q, tot = [], 0
        for _ in range(nlevels - 1):
            q.append(cv_round(nd))
            tot += q[-1]
            nd = F32(nd * factor)
        q.append(max(nfeatures - tot, 0))
        self.quota = q
        # umax
        umax = [0] * (HALF_PATCH_SIZE + 2)
        vmax = int(math.floor(HALF_PATCH_SIZE * math.sqrt(2.0) / 2 + 1))
        vmin = int(math.ceil(HALF_PATCH_SIZE * math.sqrt(2.0) / 2))
        hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE
        for v in range(vmax + 1):
            umax[v] = cv_round(math.sqrt(hp2 - v * v))
        v0 = 0
        v = HALF_PATCH_SIZE
        while v >= vmin:
            while umax[v0] == umax[v0 + 1]:
                v0 += 1
            umax[v] = v0
            v0 += 1
            v -= 1
        self.umax = umax[:HALF_PATCH_SIZE + 1]

    def level_size(self, w, h, l):
        inv = self.inv_scale[l]
        return cv_round(F32(F32(w) * inv)), cv_round(F32(F32(h) * inv))


def pyramid(img: np.ndarray, p: Params):
    """ComputePyramid: iterative INTER_LINEAR from the previous level's interior (borders never read, App. A.1)."""
    h, w = img.shape
    levels = [np.ascontiguousarray(img)]
    for l in range(1, p.nlevels):
        wl, hl = p.level_size(w, h, l)
        levels.append(cv2.resize(levels[l - 1], (wl, hl), interpolation=cv2.INTER_LINEAR))
    return levels


def cell_grid(wl, hl):
    """Cell geometry of ComputeKeyPointsOctTree for a level of size wl x hl. Returns list of
    (i, j, iniX, iniY, maxX, maxY, offx, offy) in the reference's visiting order."""
    W = F32(35)
    minBX = minBY = EDGE_THRESHOLD - 3
    maxBX, maxBY = wl - EDGE_THRESHOLD + 3, hl - EDGE_THRESHOLD + 3
    width, height = F32(maxBX - minBX), F32(maxBY - minBY)
    ncols, nrows = int(width / W), int(height / W)
    if ncols <= 0 or nrows <= 0:
        return [], (minBX, maxBX, minBY, maxBY)
    wcell = int(math.ceil(F32(width / F32(ncols))))
    hcell = int(math.ceil(F32(height / F32(nrows))))
    cells = []
    for i in range(nrows):
        iniY = minBY + i * hcell
        maxY = iniY + hcell + 6
        if iniY >= maxBY - 3:
            continue
        maxY = min(maxY, maxBY)
        for j in range(ncols):
            iniX = minBX + j * wcell
            maxX = iniX + wcell + 6
            if iniX >= maxBX - 6:
                continue
            maxX = min(maxX, maxBX)
            cells.append((i, j, iniX, iniY, maxX, maxY, j * wcell, i * hcell))
    return cells, (minBX, maxBX, minBY, maxBY)


def fast_cells(level: np.ndarray, p: Params):
    """Per-cell cv::FAST(ini) with per-cell fallback to cv::FAST(min). Returns float32 [n,3] (x,y,response)
    relative to (minBorderX,minBorderY), in the reference's push order."""
    hl, wl = level.shape
    cells, _ = cell_grid(wl, hl)
    det_ini = cv2.FastFeatureDetector_create(p.ini_th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    det_min = cv2.FastFeatureDetector_create(p.min_th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    out = []
    for (_i, _j, x0, y0, x1, y1, ox, oy) in cells:
        roi = level[y0:y1, x0:x1]
        kps = det_ini.detect(roi)
        if len(kps) == 0:
            kps = det_min.detect(roi)
        for k in kps:
            out.append((k.pt[0] + ox, k.pt[1] + oy, k.response))
    return np.array(out, dtype=F32).reshape(-1, 3)


class _Node:
    __slots__ = ("ulx", "uly", "brx", "bry", "keys", "nomore", "seq", "alive")


def distribute_octree(keys: np.ndarray, minX, maxX, minY, maxY, N):
    """DistributeOctTree + DivideNode, literal list semantics.
    keys: float32 [n,3]. Returns indices into keys, in final list order.
    Tie rule for the (size, pointer) sort: ascending (size, creation sequence) -- documented deviation,
    the reference orders equal sizes by heap address."""
    n = keys.shape[0]
    nIni = int(np.round(F32(maxX - minX) / F32(maxY - minY)))  # C round(): half away from zero
    r = float(F32(maxX - minX) / F32(maxY - minY))
    nIni = int(math.floor(r + 0.5))
    hX = F32(F32(maxX - minX) / F32(nIni))
    seq = [0]

    def mk(ulx, uly, brx, bry, idx):
        nd = _Node()
        nd.ulx, nd.uly, nd.brx, nd.bry = ulx, uly, brx, bry
        nd.keys = idx
        nd.nomore = len(idx) == 1
        nd.seq = seq[0]
        seq[0] += 1
        nd.alive = True
        return nd

    xs, ys = keys[:, 0], keys[:, 1]
    lst = []  # python list used as std::list; front = index 0
    ini_idx = (xs / hX).astype(np.int32) if n else np.zeros(0, np.int32)
    for i in range(nIni):
        ulx = int(F32(hX * F32(i)))
        urx = int(F32(hX * F32(i + 1)))
        nd = mk(ulx, 0, urx, maxY - minY, np.nonzero(ini_idx == i)[0])
        lst.append(nd)
    lst = [nd for nd in lst if len(nd.keys) > 0]

    def divide(nd):
        halfX = int(math.ceil(F32(nd.brx - nd.ulx) / F32(2)))
        halfY = int(math.ceil(F32(nd.bry - nd.uly) / F32(2)))
        mx, my = nd.ulx + halfX, nd.uly + halfY
        k = nd.keys
        left = xs[k] < F32(mx)
        top = ys[k] < F32(my)
        return [
            mk(nd.ulx, nd.uly, mx, my, k[left & top]),
            mk(mx, nd.uly, nd.brx, my, k[~left & top]),
            mk(nd.ulx, my, mx, nd.bry, k[left & ~top]),
            mk(mx, my, nd.brx, nd.bry, k[~left & ~top]),
        ]

    finish = False
    while not finish:
        prev_size = len(lst)
        to_expand = 0
        cand = []
        new_front = []  # children pushed to the front (kept reversed at the end)
        kept = []
        for nd in lst:
            if nd.nomore:
                kept.append(nd)
                continue
            for ch in divide(nd):
                if len(ch.keys) > 0:
                    new_front.append(ch)
                    if len(ch.keys) > 1:
                        to_expand += 1
                        cand.append(ch)
        lst = new_front[::-1] + kept
        if len(lst) >= N or len(lst) == prev_size:
            finish = True
        elif len(lst) + to_expand * 3 > N:
            while not finish:
                prev_size = len(lst)
                prev = sorted(cand, key=lambda c: (len(c.keys), c.seq))
                cand = []
                for nd in reversed(prev):
                    for ch in divide(nd):
                        if len(ch.keys) > 0:
                            lst.insert(0, ch)
                            if len(ch.keys) > 1:
                                cand.append(ch)
                    lst.remove(nd)
                    if len(lst) >= N:
                        break
                if len(lst) >= N or len(lst) == prev_size:
                    finish = True
    out = []
    for nd in lst:
        k = nd.keys
        best = k[0]
        br = keys[best, 2]
        for q in k[1:]:
            if keys[q, 2] > br:
                best, br = q, keys[q, 2]
        out.append(int(best))
    return np.array(out, dtype=np.int64)


def ic_angles(level: np.ndarray, pts: np.ndarray, umax):
    """IC_Angle + cv2.fastAtan2 for integer-valued level coordinates pts [n,2]."""
    out = np.zeros(len(pts), dtype=F32)
    L = level.astype(np.int64)
    for n, (x, y) in enumerate(pts):
        cx, cy = cv_round(x), cv_round(y)
        m01 = m10 = 0
        for u in range(-HALF_PATCH_SIZE, HALF_PATCH_SIZE + 1):
            m10 += u * L[cy, cx + u]
        for v in range(1, HALF_PATCH_SIZE + 1):
            d = umax[v]
            us = np.arange(-d, d + 1)
            plus = L[cy + v, cx - d:cx + d + 1]
            minus = L[cy - v, cx - d:cx + d + 1]
            m10 += int((us * (plus + minus)).sum())
            m01 += v * int((plus - minus).sum())
        out[n] = cv2.fastAtan2(float(F32(m01)), float(F32(m10)))
    return out


def blur(level: np.ndarray):
    return cv2.GaussianBlur(np.ascontiguousarray(level).copy(), (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101)


def steered_brief(blurred: np.ndarray, pts: np.ndarray, angles: np.ndarray, pattern: np.ndarray):
    """computeOrbDescriptor: fp32 products, no FMA, cvRound half-even; cosf/sinf modelled as
    the correctly rounded fp32 of the fp64 result."""
    n = len(pts)
    desc = np.zeros((n, 32), dtype=np.uint8)
    factor = F32(np.pi / 180.0)  # (float)(CV_PI/180.f): double/float -> double -> float
    px0, py0, px1, py1 = (pattern[:, i].astype(F32) for i in range(4))
    for k in range(n):
        ang = F32(angles[k] * factor)
        a, b = F32(math.cos(float(ang))), F32(math.sin(float(ang)))
        cx, cy = cv_round(pts[k, 0]), cv_round(pts[k, 1])

        def samp(px, py):
            ix = np.rint((px * a).astype(F32) - (py * b).astype(F32)).astype(np.int64)
            iy = np.rint((px * b).astype(F32) + (py * a).astype(F32)).astype(np.int64)
            return blurred[cy + iy, cx + ix]

        bits = (samp(px0, py0) < samp(px1, py1)).astype(np.uint8)
        desc[k] = np.packbits(bits, bitorder="little")
    return desc


def extract(img: np.ndarray, p: Params, lap=(0, 1000), pattern=None, keep_stages=False):
    """ORBextractor::operator() (mono). Returns dict(kps [N,7] float32-compatible struct fields, desc [N,32],
    mono_index) (+ per-stage intermediates if keep_stages)."""
    if pattern is None:
        pattern = load_pattern()
    levels = pyramid(img, p)
    per_level = []
    stages = {"levels": levels, "cand": [], "sel": [], "angles": [], "blur": [], "desc": []}
    for l, lv in enumerate(levels):
        hl, wl = lv.shape
        cand = fast_cells(lv, p)
        _, (minBX, maxBX, minBY, maxBY) = cell_grid(wl, hl)
        sel = distribute_octree(cand, minBX, maxBX, minBY, maxBY, p.quota[l])
        kp = cand[sel].copy() if len(sel) else np.zeros((0, 3), F32)
        kp[:, 0] += F32(minBX)
        kp[:, 1] += F32(minBY)
        ang = ic_angles(lv, kp[:, :2], p.umax)
        per_level.append((kp, ang))
        if keep_stages:
            stages["cand"].append(cand)
            stages["sel"].append(sel)
            stages["angles"].append(ang)
    ntot = sum(len(k) for k, _ in per_level)
    kps = np.zeros((ntot, 7), dtype=np.float64)  # x,y,size,angle,response,octave,class_id
    desc = np.zeros((ntot, 32), dtype=np.uint8)
    mono, stereo = 0, ntot - 1
    for l, (kp, ang) in enumerate(per_level):
        if len(kp) == 0:
            if keep_stages:
                stages["blur"].append(None)
                stages["desc"].append(np.zeros((0, 32), np.uint8))
            continue
        bl = blur(levels[l])
        d = steered_brief(bl, kp[:, :2], ang, pattern)
        if keep_stages:
            stages["blur"].append(bl)
            stages["desc"].append(d)
        sc = p.scale[l]
        size = F32(int(F32(F32(PATCH_SIZE) * sc)))
        for i in range(len(kp)):
            x, y = kp[i, 0], kp[i, 1]
            if l != 0:
                x, y = F32(x * sc), F32(y * sc)
            row = (x, y, size, ang[i], kp[i, 2], l, -1)
            if x >= F32(lap[0]) and x <= F32(lap[1]):
                kps[stereo] = row
                desc[stereo] = d[i]
                stereo -= 1
            else:
                kps[mono] = row
                desc[mono] = d[i]
                mono += 1
    out = {"kps": kps, "desc": desc, "mono_index": mono}
    if keep_stages:
        out["stages"] = stages
    return out


def knn2_bf(q: np.ndarray, db: np.ndarray):
    """cv::BFMatcher(NORM_HAMMING).knnMatch(k=2): returns idx [nq,2] int32, dist [nq,2] int32."""
    m = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, db, k=2)
    idx = np.full((len(q), 2), -1, np.int32)
    dist = np.full((len(q), 2), -1, np.int32)
    for i, row in enumerate(m):
        for j, mm in enumerate(row):
            idx[i, j] = mm.trainIdx
            dist[i, j] = int(mm.distance)
    return idx, dist

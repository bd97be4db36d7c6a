#!/usr/bin/env python
"""Index-level emulation of k_fast_pairs (send_slam_b200/csrc/orbx_fast2.cu) in numpy, checked against the CPU checker's per-cell FAST.

Not a product path: a development aid for a container without a GPU.  It follows the kernel's own arithmetic -- TMA box with zero
fill, pair plane U[row][m], 8-pixel unpack items, chunked rows with the rolling score tile, two-rows-per-item scoring, 2 x 2 pair NMS
items, front / back candidate list -- so that an indexing mistake shows up here instead of costing a GPU round trip.
Usage: python tools/emulate_fast2.py [ch]
"""
import sys, os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_lib as ol
from send_slam_b200 import synth

RING = [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3), (0, -3), (-1, -3), (-2, -2), (-3, -1), (-3, 0), (-3, 1), (-2, 2), (-1, 3)]


def cells_of(w, h):
    """cell ROIs of one level"""
    minb, maxbx, maxby = 16, w - 16, h - 16
    width, height = maxbx - minb, maxby - minb
    ncols, nrows = int(width / 35), int(height / 35)
    if ncols < 1 or nrows < 1:
        return []
    wc, hc = -(-width // ncols), -(-height // nrows)
    out = []
    for i in range(nrows):
        iy = minb + i * hc
        my = iy + hc + 6
        if iy >= maxby - 3:
            continue
        my = min(my, maxby)
        for j in range(ncols):
            ix = minb + j * wc
            mx = ix + wc + 6
            if ix >= maxbx - 6:
                continue
            mx = min(mx, maxbx)
            out.append((ix, iy, mx, my))
    return out


def lanes(x):      # u16x2 word -> (lo, hi) int arrays
    return x & 0xFFFF, x >> 16


def pack(lo, hi):
    return (lo & 0xFFFF) | ((hi & 0xFFFF) << 16)


def vop(f, *a):    # lane-wise op on packed words
    los, his = zip(*[lanes(x) for x in a])
    return pack(f(*los), f(*his))


def vmin(*a): return vop(lambda *x: np.minimum.reduce(x), *a)
def vmax(*a): return vop(lambda *x: np.maximum.reduce(x), *a)
def relu_sub(a, b): return vop(lambda x, y: np.maximum(x.astype(np.int64) - y, 0).astype(np.uint32), a, b)


def score_pair(r, v):
    qn = [vmin(r[2 * j + 1], r[(2 * j + 2) & 15]) for j in range(8)]
    qx = [vmax(r[2 * j + 1], r[(2 * j + 2) & 15]) for j in range(8)]
    q2x = [vmax(qx[i], qx[(i + 1) & 7]) for i in range(8)]
    q2n = [vmin(qn[i], qn[(i + 1) & 7]) for i in range(8)]
    fx, fn = [], []
    for i in range(8):
        a, b = r[2 * i], r[(2 * i + 9) & 15]
        fx.append(vmax(q2x[i], q2x[(i + 2) & 7], vmin(a, b)))
        fn.append(vmin(q2n[i], q2n[(i + 2) & 7], vmax(a, b)))
    mam = vmin(*fx)
    mim = vmax(*fn)
    return vmax(relu_sub(v, mam), relu_sub(mim, v))


def byte_perm(a, b, sel):
    src = [(a >> (8 * i)) & 0xFF for i in range(4)] + [(b >> (8 * i)) & 0xFF for i in range(4)]
    out = np.zeros_like(a)
    for i in range(4):
        out |= src[(sel >> (4 * i)) & 7] << (8 * i)
    return out


def run_level(img, ini_th, min_th, CH, P, box_w, maxnp):
    h, w = img.shape
    cells = cells_of(w, h)
    SP = (maxnp + 3 + 1) & ~1
    box_h = CH + 6
    out = []
    for (x0, y0, x1, y1) in cells:
        iw, ih = x1 - x0 - 6, y1 - y0 - 6
        npair = (iw + 1) >> 1
        off = x0 & 15; abw = off >> 2; sh = off & 3
        K8 = (2 * npair + 8 + sh + 7) >> 3
        nch = (ih + CH - 1) // CH; cr = (ih + nch - 1) // nch
        assert 8 * K8 <= P and 4 * (abw + 2 * K8 + 1) <= box_w, (K8, P, abw, box_w)
        rows_alloc = CH + 4
        sc = np.full((rows_alloc + 2) * SP, 0xDEADBEEF, dtype=np.uint32)   # poison
        t = np.arange(cr + 4)
        sc[t * SP] = 0; sc[t * SP + npair + 1] = 0; sc[t * SP + npair + 2] = 0
        sc[SP + np.arange(npair + 3)] = 0
        lst_hi, lst_lo = [], []
        prev_rows = 0
        for c in range(nch):
            a = c * cr; b = min(ih, a + cr); nrows = b - a; last = c == nch - 1
            # TMA: box at (x0 & ~15, y0 + a), box_w x box_h, zero fill outside the plane
            bx, by = x0 & ~15, y0 + a
            roi = np.zeros((box_h, box_w), dtype=np.uint8)
            ys = np.arange(by, by + box_h); xs = np.arange(bx, bx + box_w)
            vy = ys < h; vx = xs < w
            roi[np.ix_(vy, vx)] = img[np.ix_(ys[vy], xs[vx])]
            roiw = np.concatenate([roi.reshape(-1), np.zeros(16, np.uint8)]).view(np.uint32)   # + stage slack
            rpw = box_w >> 2
            # unpack
            U = np.full(box_h * P + 64, 0xABABABAB, dtype=np.uint32)
            it = np.arange((nrows + 6) * K8)
            row = it // K8; k = it - row * K8
            src = row * rpw + abw + 2 * k
            w0, w1, w2 = roiw[src].astype(np.uint32), roiw[src + 1].astype(np.uint32), roiw[src + 2].astype(np.uint32)
            dst = row * P + 8 * k
            z = np.zeros_like(w0)
            fs = lambda lo, hi: ((lo >> 24) | (hi << 8)) & 0xFFFFFFFF
            U[dst + 0] = byte_perm(w0, z, 0x4140); U[dst + 1] = byte_perm(w0, z, 0x4241); U[dst + 2] = byte_perm(w0, z, 0x4342)
            U[dst + 3] = byte_perm(fs(w0, w1), z, 0x4140)
            U[dst + 4] = byte_perm(w1, z, 0x4140); U[dst + 5] = byte_perm(w1, z, 0x4241); U[dst + 6] = byte_perm(w1, z, 0x4342)
            U[dst + 7] = byte_perm(fs(w1, w2), z, 0x4140)
            if c > 0:
                i = np.arange(2 * SP); rr = (i >= SP).astype(int); col = i - rr * SP
                m = col < npair + 3
                sc[(rr * SP + col)[m]] = sc[((prev_rows + rr) * SP + col)[m]]
            # scores
            H = (nrows + 1) >> 1
            it = np.arange(npair * H); j = it // H; rr = it - j * H
            ad = sh + rr * P + 2 * j

            def score_at(ad):
                r = [U[ad + (dy + 3) * P + dx + 3].astype(np.uint32) for (dx, dy) in RING]
                v = U[ad + 3 * P + 3].astype(np.uint32)
                assert not np.any(v == 0xABABABAB) and not any(np.any(x == 0xABABABAB) for x in r)
                return score_pair(r, v)
            so = (rr + 2) * SP + j + 1
            sc[so] = score_at(ad)
            m = rr + H < nrows
            sc[so[m] + H * SP] = score_at(ad[m] + H * P)
            if iw & 1:
                t = np.arange(2, nrows + 2); sc[t * SP + npair] &= 0xFFFF
            if last:
                sc[(nrows + 2) * SP + np.arange(npair + 3)] = 0
            # NMS
            first = 0 if c == 0 else a - 1; lastrow = ih - 1 if last else b - 2
            QH = (lastrow - first + 2) >> 1; JH = (npair + 1) >> 1
            it = np.arange(JH * QH); q = it // JH; jj = it - q * JH
            rt = first + 2 * q
            pa = (rt - a + 1) * SP + 2 * jj
            wv = [[sc[pa + r4 * SP + cc].astype(np.uint32) for cc in range(4)] for r4 in range(4)]
            thr2 = np.uint32(min_th | (min_th << 16))
            for rr2 in range(2):
                valid_r = rt + rr2 <= lastrow
                T = [vmax(wv[rr2][cc], wv[rr2 + 1][cc], wv[rr2 + 2][cc]) for cc in range(4)]
                x01, x12, x23 = byte_perm(T[0], T[1], 0x5432), byte_perm(T[1], T[2], 0x5432), byte_perm(T[2], T[3], 0x5432)
                for cc in range(2):
                    mid = wv[rr2 + 1][cc + 1]
                    vv = vmax(wv[rr2][cc + 1], wv[rr2 + 2][cc + 1], np.full_like(mid, thr2))
                    m_ = vmax(x12 if cc else x01, x23 if cc else x12, vv)
                    fl = vmax(mid, m_) ^ m_
                    jcol = 2 * jj + cc
                    ok = valid_r & (jcol < npair) & (fl != 0)
                    for idx in np.nonzero(ok)[0]:
                        assert mid[idx] != 0xDEADBEEF
                        for kk in range(2):
                            if (int(fl[idx]) >> (16 * kk)) & 0xFFFF:
                                s = (int(mid[idx]) >> (16 * kk)) & 0xFFFF
                                e = (int(rt[idx]) + rr2, 2 * int(jcol[idx]) + kk, s)
                                (lst_hi if s > ini_th else lst_lo).append(e)
            prev_rows = nrows
        use = lst_hi if lst_hi else lst_lo
        for (ry, cx, s) in use:
            out.append((cx + x0 + 3 - 16, ry + y0 + 3 - 16, s - 1))
    return out


def main():
    CH = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    rng = np.random.default_rng(1)
    cases = [(640, 480, "textured"), (309, 231, "textured"), (179, 134, "mixed"), (533, 400, "lowcontrast"), (214, 161, "sparse"),
             (147, 101, "textured"), (100, 75, "textured"), (257, 193, "textured"), (1920 // 4, 1080 // 4, "mixed")]
    for (w, h, kind) in cases:
        img = synth.textured_frame(int(rng.integers(1000)), w, h, kind)
        cl = cells_of(w, h)
        if not cl:
            continue
        max_iw = max(c[2] - c[0] - 6 for c in cl); max_ih = max(c[3] - c[1] - 6 for c in cl)
        maxnp = (max_iw + 1) // 2
        K = (2 * maxnp + 8 + 3 + 7) // 8
        box_w = (12 + 4 * (2 * K + 1) + 15) // 16 * 16
        P = next(p for p in (49, 57, 73, 89) if p >= 8 * K + 1)
        ch = min(CH, max_ih)
        got = run_level(img, 20, 7, ch, P, box_w, maxnp)
        ref = ol.fast_cells(img, 20, 7)
        gs = sorted((int(a), int(b), int(c)) for a, b, c in got)
        rs = sorted((int(a), int(b), int(c)) for a, b, c in ref)
        print(f"{w}x{h} {kind:12s} cells {len(cl):4d} max cell {max_iw}x{max_ih} P {P} box_w {box_w} ch {ch}: emulation {len(gs)} candidates, checker {len(rs)}: "
              f"{'EQUAL' if gs == rs else 'DIFFERENT'}")
        if gs != rs:
            d1 = sorted(set(gs) - set(rs))[:5]; d2 = sorted(set(rs) - set(gs))[:5]
            print("  only emulation:", d1, " only checker:", d2)
            sys.exit(1)


if __name__ == "__main__":
    main()

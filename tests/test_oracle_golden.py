"""CPU: the C oracle against the committed golden vectors (made by tests/golden/make_golden.py from real OpenCV code
through cv2 + the Python restatement) and, where cv2 is importable, directly against the OpenCV primitives."""
import glob
import hashlib
import os

import numpy as np
import pytest

from send_slam_b200 import synth


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def golden_cases(golden_dir):
    return sorted(p for p in glob.glob(os.path.join(golden_dir, "*.npz")) if not os.path.basename(p).startswith(("knn2", "pnm")))


def test_golden_files_present(golden_dir):
    assert len(golden_cases(golden_dir)) >= 5
    assert os.path.exists(os.path.join(golden_dir, "knn2_4096x128.npz"))


def test_oracle_matches_golden_extraction(oracle, golden_dir):
    for path in golden_cases(golden_dir):
        g = np.load(path)
        frame = synth.textured_frame(int(g["seed"]), int(g["width"]), int(g["height"]), str(g["kind"]))
        assert sha(frame) == str(g["frame_sha"]), "synthetic generator drifted: regenerate the fixtures"
        o = oracle.Oracle(int(g["nfeatures"]), 1.2, 8, 20, 7)
        kps, desc, mono = o.extract(frame, lap=tuple(int(v) for v in g["lap"]))
        ref = g["kps"]
        assert len(kps) == len(ref), path
        assert mono == int(g["mono_index"])
        for i, name in enumerate(["x", "y", "size", "angle", "response", "octave", "class_id"]):
            assert np.array_equal(kps[name].astype(np.float32), ref[:, i]), (path, name)
        assert np.array_equal(desc, g["desc"]), path
        for l in range(8):
            lv, bl = o.stage_level(l)
            assert sha(lv) == str(g["level_sha"][l]), (path, "pyramid level", l)
            if bl is not None:
                assert sha(bl) == str(g["blur_sha"][l]), (path, "blurred level", l)
            c, s, a, d = o.stage_keys(l)
            assert len(c) == int(g["ncand"][l]) and len(s) == int(g["nsel"][l]), (path, l)


def test_oracle_matches_golden_knn(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "knn2_4096x128.npz"))
    db = synth.descriptor_db(4096, seed=int(g["db_seed"]))
    q, src = synth.queries_from_db(db, 128, seed=int(g["q_seed"]))
    db[100] = db[7]
    db[2000] = db[7]
    idx, dist = oracle.knn2(q, db, nthreads=2)
    assert np.array_equal(idx, g["idx"]) and np.array_equal(dist, g["dist"])
    # DescriptorDistance (SWAR) == popcount(xor) == BFMatcher distance
    for i in range(0, 128, 17):
        assert oracle.distance(q[i], db[idx[i, 0]]) == dist[i, 0]
        assert int(np.unpackbits(q[i] ^ db[idx[i, 1]]).sum()) == dist[i, 1]


def test_oracle_tables(oracle):
    o = oracle.Oracle(1000, 1.2, 8, 20, 7)
    t = o.tables()
    assert t["quota"].tolist() == [217, 181, 151, 126, 105, 87, 73, 60]          # SURVEY.md §8 cfg 1
    assert t["umax"].tolist() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert [o.level_size(640, 480, l) for l in range(8)] == [(640, 480), (533, 400), (444, 333), (370, 278), (309, 231),
                                                             (257, 193), (214, 161), (179, 134)]
    assert oracle.Oracle(1200).tables()["quota"].tolist() == [261, 217, 181, 151, 126, 105, 87, 72]
    assert [oracle.Oracle(2000).level_size(1920, 1080, l) for l in (1, 7)] == [(1600, 900), (536, 301)]
    np.testing.assert_allclose(t["scale"], [1, 1.2000000477, 1.4400000572, 1.7280001640, 2.0736002922, 2.4883203506,
                                            2.9859845638, 3.5831816196], rtol=1e-7)


def test_oracle_edge_cases(oracle):
    o = oracle.Oracle(500)
    # constant frame: no keypoints, monoIndex 0 (the reference releases the descriptor matrix)
    kps, desc, mono = o.extract(np.full((240, 320), 77, np.uint8))
    assert len(kps) == 0 and desc.shape == (0, 32) and mono == 0
    # frame too small for any FAST cell on the upper levels still works
    f = synth.textured_frame(1, 120, 100)
    kps, desc, mono = o.extract(f)
    assert 0 < len(kps) <= 500 + 24
    assert np.all(kps["x"] >= 19 * 0.999) and np.all(kps["octave"] >= 0)
    # octree with N = 0 and with a single key
    keys = np.array([[10, 10, 50], [100, 40, 60], [200, 90, 70]], np.float32)
    assert len(oracle.octree(keys, 16, 16 + 320, 16, 16 + 200, 0)) >= 1
    assert oracle.octree(keys[:1], 16, 16 + 320, 16, 16 + 200, 5).tolist() == [0]
    assert len(oracle.octree(np.zeros((0, 3), np.float32), 16, 336, 16, 216, 5)) == 0


def test_trig_rule_moves_few_bits(oracle):
    """The reference calls libm cosf/sinf, which is not correctly rounded; the canonical oracle rule is the correctly
    rounded value.  The two must agree on >= 99.9 % of descriptor bits (north star tolerance)."""
    f = synth.textured_frame(5, 640, 480)
    o = oracle.Oracle(1000)
    _, d_cr, _ = o.extract(f)
    oracle.set_trig_mode(0)
    try:
        _, d_libm, _ = o.extract(f)
    finally:
        oracle.set_trig_mode(1)
    diff = int(np.unpackbits(d_cr ^ d_libm).sum())
    assert diff <= 1e-3 * d_cr.size * 8


cv2 = pytest.importorskip("cv2", reason="cv2 pins the OpenCV primitives; fixtures cover the rest")


def test_primitives_against_opencv(oracle):
    rng = np.random.default_rng(0)
    for (w, h) in [(640, 480), (333, 217), (64, 48)]:
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        tex = synth.textured_frame(w + h, w, h)
        for src in (img, tex):
            dw, dh = int(round(w / 1.2)), int(round(h / 1.2))
            assert np.array_equal(oracle.resize(src, dw, dh), cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR))
            assert np.array_equal(oracle.blur7(src), cv2.GaussianBlur(src.copy(), (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101))
            for t in (20, 7):
                det = cv2.FastFeatureDetector_create(t, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
                ref = np.array([(k.pt[0], k.pt[1], k.response) for k in det.detect(src)], np.float32).reshape(-1, 3)
                assert np.array_equal(oracle.fast(src, t), ref)
    ys = rng.integers(-3_000_000, 3_000_000, 20000).astype(np.float32)
    xs = rng.integers(-3_000_000, 3_000_000, 20000).astype(np.float32)
    for y, x in list(zip(ys, xs))[:5000] + [(0.0, 0.0), (0.0, -1.0), (-1.0, 0.0), (5.0, 5.0), (-5.0, 5.0)]:
        assert oracle.fast_atan2(y, x) == np.float32(cv2.fastAtan2(float(y), float(x)))


def test_brief_against_cv2_orb(oracle):
    """Steered BRIEF: same pattern + steering formula as cv2.ORB -> identical bits for given keypoints and angles
    (up to the libm-vs-correctly-rounded trig rule, < 0.1 % of bits)."""
    f = synth.textured_frame(9, 320, 240)
    bl = cv2.GaussianBlur(f.copy(), (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101)
    rng = np.random.default_rng(3)
    pts = np.stack([rng.integers(40, 280, 200), rng.integers(40, 200, 200)], 1).astype(np.float32)
    orb = cv2.ORB_create(nfeatures=500, scaleFactor=1.2, nlevels=1, edgeThreshold=19, patchSize=31)
    o = oracle.Oracle(500)
    kps = [cv2.KeyPoint(float(x), float(y), 31, float(o.L.orb_oracle_ic_angle(o.h, oracle._p(f), f.shape[1], float(x), float(y))), 1, 0)
           for x, y in pts]
    # cv2.ORB.compute blurs internally with the same 7x7 sigma-2 kernel when run on the raw level
    kps2, desc = orb.compute(f, kps)
    assert len(kps2) == len(kps)
    mine = np.stack([oracle.brief(bl, k.pt[0], k.pt[1], k.angle) for k in kps2])
    diff = int(np.unpackbits(mine ^ desc).sum())
    assert diff <= 1e-3 * desc.size * 8, diff


def test_oracle_gray_matches_cv2_and_pins(oracle):
    """cv::cvtColor(*2GRAY) restatement: bit-identical to cv2 where it is importable; SHA pins (made in the
    build container from the cv2-verified output) travel to machines without it."""
    rng = np.random.default_rng(2024)
    img = rng.integers(0, 256, (97, 131, 3), dtype=np.uint8)
    img4 = rng.integers(0, 256, (97, 131, 4), dtype=np.uint8)
    pins = {1: ("e7335b921ef03000", "b352c2ddaf675db7"), 2: ("da27676cd09f3c3f", "0ed6f66dacaa68de"),
            3: ("0d491f1097b7e592", "b41cea11e9741c38"), 4: ("02a4fe3fc4cd48dc", "1c4bd514d10ed9f9")}
    try:
        import cv2
        codes = {1: cv2.COLOR_RGB2GRAY, 2: cv2.COLOR_BGR2GRAY, 3: cv2.COLOR_RGBA2GRAY, 4: cv2.COLOR_BGRA2GRAY}
    except ImportError:
        cv2, codes = None, {}
    for fmt in (1, 2, 3, 4):
        src = img if fmt < 3 else img4
        g15, g14 = oracle.gray(src, fmt, 15), oracle.gray(src, fmt, 14)
        assert sha(g15)[:16] == pins[fmt][0] and sha(g14)[:16] == pins[fmt][1], fmt
        if cv2 is not None:
            assert np.array_equal(g15, cv2.cvtColor(src, codes[fmt])), fmt
    # the two fixed-point forms differ on a fraction of a percent of the pixels
    frac = float((oracle.gray(img, 1, 15) != oracle.gray(img, 1, 14)).mean())
    assert 0 < frac < 0.01


UNDISTORT_PINS = ['b72560e635450f41', '2df5dc6e0f1f241a', '0a789e7802a91ad2']
CAMERAS = [((458.654, 457.296, 367.215, 248.375, -0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0), (752, 480)),   # EuRoC cam0
           ((520.9, 521.0, 325.1, 249.7, 0.2624, -0.9531, -0.0054, 0.0026, 1.1633), (640, 480)),                           # TUM fr1-like, k3 != 0
           ((900.0, 905.0, 640.0, 360.0, -0.12, 0.05, 0.001, -0.0007, 0.0), (1280, 720))]


def test_oracle_undistort_matches_cv2_and_pins(oracle):
    """Frame::UndistortKeyPoints = cv::undistortPoints(K, D, P=K): the restatement equals cv2 bit for bit
    where cv2 is importable; SHA pins of the verified output travel."""
    try:
        import cv2
    except ImportError:
        cv2 = None
    rng = np.random.default_rng(5)
    shas = []
    for cam, (w, h) in CAMERAS:
        xy = np.stack([rng.uniform(0, w, 5000), rng.uniform(0, h, 5000)], 1).astype(np.float32)
        got = oracle.undistort_points(xy, cam)
        shas.append(sha(got)[:16])
        if cv2 is not None:
            K = np.array([[cam[0], 0, cam[2]], [0, cam[1], cam[3]], [0, 0, 1]], np.float32)
            ref = cv2.undistortPoints(xy.reshape(-1, 1, 2), K, np.array(cam[4:9], np.float32), None, K).reshape(-1, 2)
            assert np.array_equal(got, ref), cam
        b = oracle.image_bounds(cam, w, h)
        assert b[0] < b[2] and b[1] < b[3]
    assert shas == UNDISTORT_PINS, shas
    # no distortion: bounds are the image, keypoints are copied
    assert np.array_equal(oracle.image_bounds((500, 500, 320, 240, 0, 0, 0, 0), 640, 480), np.array([0, 0, 640, 480], np.float32))


def test_oracle_frame_grid_properties(oracle):
    cam, (w, h) = CAMERAS[0]
    o = oracle.Oracle(1200)
    kps, desc, mono = o.extract(synth.textured_frame(3, w, h))
    b = oracle.image_bounds(cam, w, h)
    un, start, items = oracle.frame_grid(kps, cam, b)
    assert len(un) == len(kps) and start[0] == 0 and start[-1] == len(items) <= len(kps)
    assert np.array_equal(un["x"], oracle.undistort_points(np.stack([kps["x"], kps["y"]], 1), cam)[:, 0])
    invw, invh = np.float32(64) / (b[2] - b[0]), np.float32(48) / (b[3] - b[1])
    for c in np.nonzero(np.diff(start))[0][:200]:
        seg = items[start[c]:start[c + 1]]
        assert np.all(np.diff(seg) > 0)                                          # push_back order
        px = np.floor((un["x"][seg] - b[0]) * invw + np.float32(0.5)).astype(int)  # round() for non-negative arguments
        assert np.all(px == c // 48)


def _pnm_cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "pnm_cases.npz"))
    blob, out, pos = g["blob"].tobytes(), [], 0
    for i, ln in enumerate(g["lengths"]):
        out.append((blob[pos:pos + ln], int(g["verdict"][i]), tuple(int(v) for v in g["shape"][i]), str(g["digest"][i])))
        pos += ln
    return g, out


def test_oracle_pnm_decode_matches_cv2_fixture(oracle, golden_dir):
    """cv::imdecode of the wire's binary PNM (orbslam3_mono_networked.cc:546): the restatement returns, case by case, what cv2 4.13
    returned when the fixture was made (tests/golden/make_golden_pnm.py) -- same accept / reject verdict, same Mat bytes."""
    g, cases = _pnm_cases(golden_dir)
    assert len(cases) == int(g["n"]) and sum(v == 1 for _, v, _, _ in cases) >= 10
    for data, verdict, shape, digest in cases:
        if verdict == 2:
            with pytest.raises(ValueError):
                oracle.pnm_decode(data)
            continue
        m = oracle.pnm_decode(data)
        if verdict == 0:
            assert m is None, data[:16]
        else:
            assert m is not None and (m.shape[0], m.shape[1], 1 if m.ndim == 2 else 3) == shape and sha(m) == digest, data[:16]
    # the wire frame: imencode(".ppm") header form, decoded Mat and both gray conversions
    w, h = (int(v) for v in g["wire_size"])
    bgr = np.stack([synth.textured_frame(int(s), w, h) for s in g["wire_seeds"]], axis=2)
    wire = g["wire_header"].tobytes() + bgr[:, :, ::-1].tobytes()
    assert sha(np.frombuffer(wire, np.uint8)) == str(g["wire_sha"])
    dec = oracle.pnm_decode(wire)
    assert sha(dec) == str(g["decoded_sha"]) and np.array_equal(dec, bgr)
    assert sha(oracle.gray(dec, 1)) == str(g["gray_rgb1_sha"]) and sha(oracle.gray(dec, 2)) == str(g["gray_rgb0_sha"])
    try:
        import cv2
    except ImportError:
        return
    for data, verdict, shape, digest in cases:          # live cv2 where it exists
        try:
            ref = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_UNCHANGED) if data else None
        except cv2.error:
            ref = None
        assert (ref is None) == (verdict == 0), data[:16]

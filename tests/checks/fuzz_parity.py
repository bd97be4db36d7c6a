#!/usr/bin/env python
"""Randomised parity sweep on a B200: random frame sizes / kinds / extractor parameters through orbx_extract (and batches through
the blocking, submit/collect, device-resident, colour, PPM and frame-message paths), every result compared with the CPU oracle bit for bit.  Usage: python tests/checks/fuzz_parity.py [seconds] [seed]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import msgpack
import torch
from send_slam_b200 import orbx, synth
from oracle import oracle_lib as ol

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
t0, ncase, nkp, nskip = time.time(), 0, 0, 0
kinds = ["textured", "mixed", "lowcontrast", "sparse"]
while time.time() - t0 < budget:
    w, h = int(rng.integers(64, 1400)), int(rng.integers(64, 900))
    if not (0.6 <= w / h <= 4.0):      # quadtree roots = round(w / h) must be 1..8 (and >= 1 in the reference itself)
        continue
    nf = int(rng.choice([300, 500, 1000, 1250, 2000, 4000]))
    scale = float(rng.choice([1.2, 1.2, 1.2, 1.1, 1.35, 1.5]))
    nlev = int(rng.choice([8, 8, 8, 4, 6, 10]))
    ini, mn = int(rng.choice([20, 20, 30, 12])), int(rng.choice([7, 7, 5, 10]))
    if mn > ini: mn = ini
    kind = kinds[int(rng.integers(0, len(kinds)))]
    B = int(rng.choice([1, 1, 1, 3, 9]))
    try:
        o = ol.Oracle(nf, scale, nlev, ini, mn)
    except Exception:
        continue
    frames = np.stack([synth.textured_frame(int(rng.integers(0, 1 << 30)), w, h, kind) for _ in range(B)])
    e = orbx.ORBextractor(nf, scale, nlev, ini, mn, max_width=w, max_height=h, max_batch=B)
    mode = int(rng.integers(0, 4)) if B > 1 else 0
    try:
        e(frames[0])          # geometry the reference itself cannot handle (degenerate upper levels) is refused with ORBX_E_INVALID
    except orbx.OrbxError as err:
        if err.code == orbx.ORBX_E_INVALID:
            nskip += 1; e.close(); continue
        raise
    tag = (w, h, nf, scale, nlev, ini, mn, kind, B, mode)
    if mode == 0:
        got = [e(frames[i]) for i in range(B)]
    elif mode == 1:
        mono, n, kps, desc = e.extract_batch(frames)
        got = [(int(mono[i]), kps[i, :n[i]], desc[i, :n[i]]) for i in range(B)]
    elif mode == 3:       # asynchronous pair, pageable or page-locked buffers, twice (the second call may replay a graph)
        src = torch.from_numpy(frames).pin_memory().numpy() if ncase & 1 else frames
        for rep in range(2):
            e.extract_batch_submit(src)
            mono, n, kps, desc = e.extract_batch_collect()
        got = [(int(mono[i]), kps[i, :n[i]], desc[i, :n[i]]) for i in range(B)]
    else:
        cap = e.capacity
        d_in = torch.from_numpy(frames).cuda()
        d_kp = torch.zeros((B, cap, 7), dtype=torch.float32, device="cuda"); d_desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device="cuda")
        d_n = torch.zeros(B, dtype=torch.int32, device="cuda"); d_mono = torch.zeros(B, dtype=torch.int32, device="cuda")
        for rep in range(3):   # third call replays the CUDA graph
            e.extract_batch_device(d_in.data_ptr(), w * h, B, w, h, w, d_kp.data_ptr(), d_desc.data_ptr(), cap, d_n.data_ptr(), d_mono.data_ptr())
        e.sync()
        n = d_n.cpu().numpy(); mono = d_mono.cpu().numpy()
        kp = d_kp.cpu().numpy().view(orbx.KP_DTYPE).reshape(B, cap); dd = d_desc.cpu().numpy()
        got = [(int(mono[i]), kp[i, :n[i]], dd[i, :n[i]]) for i in range(B)]
    # every third case also goes through the colour path (gray conversion on the device) and the undistort + grid kernel
    if ncase % 3 == 0:
        fmt = int(rng.integers(1, 5))
        ch = 4 if fmt >= 3 else 3
        col = rng.integers(0, 256, (h, w, ch), dtype=np.uint8)
        gray = ol.gray(col, fmt, 15)
        e.set_input_format(fmt, 15)
        mono_c, kps_c, desc_c = e(col)
        e.set_input_format(orbx.FMT_GRAY8)
        k_g, d_g, m_g = o.extract(gray)
        if not (mono_c == m_g and np.array_equal(desc_c, d_g) and np.array_equal(kps_c.view(np.int32), k_g.view(np.int32))):
            print("MISMATCH colour", tag, fmt); sys.exit(1)
        if ch == 3:           # the same pixels as a wire frame: binary PPM, bare and inside the MessagePack frame message
            ppm = b"P6\n# fuzz\n%d %d\n255\n" % (w, h) + col.tobytes()
            mono_p, kps_p, desc_p, size = e.extract_pnm(ppm, camera_rgb=(fmt == 2))
            r = e.process_frame_message(msgpack.packb({"type": "frame", "camera_id": 1 + ncase, "timestamp": 0.5 * ncase, "frame": ppm}, use_bin_type=True),
                                        camera_rgb=(fmt == 2))
            if not (size == (w, h) and mono_p == m_g and np.array_equal(desc_p, d_g) and np.array_equal(kps_p.view(np.int32), k_g.view(np.int32)) and
                    r["mono_index"] == m_g and np.array_equal(r["descriptors"], d_g) and r["camera_id"] == 1 + ncase):
                print("MISMATCH ppm", tag, fmt); sys.exit(1)
        cam = (0.9 * w, 0.92 * w, 0.5 * w + 3, 0.5 * h - 2, float(rng.uniform(-0.3, 0.3)), float(rng.uniform(-0.1, 0.1)), 1e-3, -5e-4, 0.0)
        b = e.image_bounds(cam, w, h)
        if np.array_equal(b, ol.image_bounds(cam, w, h)) and b[2] > b[0] and b[3] > b[1]:
            un, st, it = e.frame_grid(kps_c, cam, b)
            un_o, st_o, it_o = ol.frame_grid(kps_c, cam, b)
            if not (np.array_equal(un.view(np.int32), un_o.view(np.int32)) and np.array_equal(st, st_o) and np.array_equal(it, it_o)):
                print("MISMATCH grid", tag, cam); sys.exit(1)
        elif not np.array_equal(b, ol.image_bounds(cam, w, h)):
            print("MISMATCH bounds", tag, cam); sys.exit(1)
    for i in range(B):
        k_o, d_o, m_o = o.extract(frames[i])
        mono, kps, desc = got[i]
        ok = mono == m_o and len(kps) == len(k_o) and np.array_equal(np.ascontiguousarray(kps).view(np.int32), k_o.view(np.int32)) and np.array_equal(desc, d_o)
        if not ok:
            print("MISMATCH", tag, "frame", i, "n", len(kps), len(k_o), "mono", mono, m_o)
            sys.exit(1)
        nkp += len(kps)
    e.close()
    ncase += 1
print(f"fuzz ok: {ncase} configurations ({nskip} refused as unsupported geometry), {nkp} keypoints + descriptors bit-identical to the oracle in {time.time() - t0:.0f} s")

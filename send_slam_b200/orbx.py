"""Host-side mirror of the reference's operator interface for the ORB hot path, on top of the C ABI (include/orbx.h).

`ORBextractor` / `ORBmatcher` keep the names, argument meaning and error behaviour of UPSTREAM ORB-SLAM3
include/ORBextractor.h / include/ORBmatcher.h (the classes the reference builds from src/ORBextractor.cc and
src/ORBmatcher.cc, slam_backends/orb_slam_3/CMakeLists.txt:52-53,80-81, and drives from
orbslam3_mono_networked.cc:594).  Everything that computes goes through liborbx.so (hand-written CUDA, sm_100a);
there is no CPU fallback: importing works without a GPU, constructing an extractor without one raises OrbxError.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liborbx.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])  # == cv::KeyPoint == orbx_keypoint

ORBX_OK, ORBX_E_INVALID, ORBX_E_CUDA, ORBX_E_CAPACITY, ORBX_E_EMPTY, ORBX_E_OVERFLOW = 0, -1, -2, -3, -4, -5
FMT_GRAY8, FMT_RGB8, FMT_BGR8, FMT_RGBA8, FMT_BGRA8 = 0, 1, 2, 3, 4     # ORBX_FMT_*
GRAY_Q15, GRAY_Q14 = 15, 14                                              # ORBX_GRAY_*


class OrbxError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"orbx error {code}: {msg}")
        self.code = code


class _Config(C.Structure):
    _fields_ = [("nfeatures", C.c_int), ("scale_factor", C.c_float), ("nlevels", C.c_int), ("ini_th_fast", C.c_int),
                ("min_th_fast", C.c_int), ("device", C.c_int), ("max_width", C.c_int), ("max_height", C.c_int),
                ("max_batch", C.c_int)]


class _WireFrame(C.Structure):          # orbx_wire_frame (include/orbx_wire.h)
    _fields_ = [("type", C.c_void_p), ("type_len", C.c_size_t), ("image", C.c_void_p), ("image_bytes", C.c_size_t),
                ("timestamp", C.c_double), ("camera_id", C.c_int), ("has_timestamp", C.c_int), ("has_camera_id", C.c_int)]


class _WireFeatures(C.Structure):       # orbx_wire_features
    _fields_ = [("timestamp", C.c_double), ("camera_id", C.c_int), ("width", C.c_int), ("height", C.c_int), ("mono_index", C.c_int),
                ("n", C.c_int), ("keypoints", C.c_void_p), ("descriptors", C.c_void_p)]


# every symbol include/orbx.h declares (tests check the library exports all of them)
EXPORTS = [
    "orbx_create", "orbx_destroy", "orbx_last_error", "orbx_keypoint_capacity", "orbx_get_tables", "orbx_get_level_sizes",
    "orbx_extract", "orbx_extract_batch", "orbx_extract_batch_submit", "orbx_extract_batch_collect",
    "orbx_extract_batch_device", "orbx_sync", "orbx_launch_count", "orbx_pnm_header", "orbx_extract_pnm",
    "orbx_wire_parse_frame", "orbx_wire_process_frame", "orbx_wire_features_bound", "orbx_wire_pack_features", "orbx_wire_parse_features",
    "orbx_wire_copy_keypoints", "orbx_comm_unique_id", "orbx_comm_create", "orbx_comm_adopt", "orbx_comm_destroy", "orbx_comm_last_error",
    "orbx_knn2_query_sharded_device", "orbx_knn2_query_sharded", "orbx_host_alloc", "orbx_host_free",
    "orbx_debug_get_level", "orbx_debug_get_candidates", "orbx_debug_get_level_keypoints", "orbx_debug_resize",
    "orbx_debug_blur", "orbx_debug_octree", "orbx_debug_describe", "orbx_distance_batch", "orbx_match_windowed",
    "orbx_knn2_create_db", "orbx_knn2_create_db_device", "orbx_knn2_destroy_db", "orbx_knn2_last_error", "orbx_knn2_query",
    "orbx_knn2_query_device", "orbx_knn2_merge_device", "orbx_knn2_sync", "orbx_knn2_launch_count", "orbx_plan_probe",
    "orbx_version", "orbx_set_profiling", "orbx_get_stage_times", "orbx_set_stream",
    "orbx_knn2_set_stream", "orbx_knn2_set_backend", "orbx_set_input_format", "orbx_debug_gray",
    "orbx_undistort_points", "orbx_image_bounds", "orbx_frame_grid", "orbx_frame_grid_batch_device",
    "orbx_match_windowed_grid_device", "orbx_match_windowed_grid_batch_device",
    "orbx_vocab_create", "orbx_vocab_destroy", "orbx_vocab_last_error", "orbx_vocab_depth", "orbx_vocab_transform",
]

_lib = None


def pnm_header(data: bytes):
    """(width, height, channels, payload_offset) of a binary PNM, None where cv::imdecode would return an empty Mat."""
    buf = np.frombuffer(data, np.uint8)
    w, h, ch, off = C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
    rc = lib().orbx_pnm_header(_p(buf) if len(buf) else None, len(buf), C.byref(w), C.byref(h), C.byref(ch), C.byref(off))
    if rc == ORBX_E_EMPTY or (rc and not len(buf)):
        return None
    if rc:
        raise OrbxError(rc, "PNM variant outside binary 8-bit P5 / P6")
    return w.value, h.value, ch.value, off.value


def wire_parse_frame(payload: bytes):
    """ParseMessage of the backend (orbslam3_mono_networked.cc:302-337) for a non-calibration message: dict with type, image
    (bytes or None), timestamp / camera_id (None if absent); raises OrbxError where ParseMessage throws or returns false."""
    buf = np.frombuffer(payload, np.uint8)
    m = _WireFrame()
    rc = lib().orbx_wire_parse_frame(_p(buf) if len(buf) else None, len(buf), C.byref(m))
    if rc:
        raise OrbxError(rc, "Failed to parse MessagePack payload")
    base = buf.ctypes.data
    return {"type": bytes(buf[m.type - base:m.type - base + m.type_len]).decode("utf-8", "replace"),
            "image": bytes(buf[m.image - base:m.image - base + m.image_bytes]) if m.image else None,
            "timestamp": m.timestamp if m.has_timestamp else None, "camera_id": m.camera_id if m.has_camera_id else None}


def wire_pack_features(timestamp, camera_id, width, height, mono_index, kps, desc, framed=True) -> bytes:
    """The 'features' message of SURVEY.md §8f-3 (include/orbx_wire.h)."""
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    n = len(kps)
    out = np.zeros(lib().orbx_wire_features_bound(n, int(framed)), np.uint8)
    wr = C.c_size_t()
    rc = lib().orbx_wire_pack_features(float(timestamp), int(camera_id), int(width), int(height), int(mono_index), _p(kps) if n else None,
                                       _p(desc) if n else None, n, int(framed), _p(out), len(out), C.byref(wr))
    if rc:
        raise OrbxError(rc, "orbx_wire_pack_features")
    return out[:wr.value].tobytes()


def wire_parse_features(payload: bytes):
    buf = np.frombuffer(payload, np.uint8)
    m = _WireFeatures()
    rc = lib().orbx_wire_parse_features(_p(buf) if len(buf) else None, len(buf), C.byref(m))
    if rc:
        raise OrbxError(rc, "not a features message")
    base = buf.ctypes.data
    kp = np.frombuffer(bytes(buf[m.keypoints - base:m.keypoints - base + m.n * 28]), KP_DTYPE)
    de = np.frombuffer(bytes(buf[m.descriptors - base:m.descriptors - base + m.n * 32]), np.uint8).reshape(m.n, 32)
    return {"timestamp": m.timestamp, "camera_id": m.camera_id, "width": m.width, "height": m.height, "mono_index": m.mono_index,
            "keypoints": kp, "descriptors": de}


def lib():
    """Loads liborbx.so (built in-tree by __graft_entry__.build()); fails loudly if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise OrbxError(ORBX_E_CUDA, f"{_SO} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                     "(there is no CPU fallback)")
    L = C.CDLL(_SO)
    vp, ip, fp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float)
    L.orbx_create.argtypes = [C.POINTER(_Config), C.POINTER(vp)]
    L.orbx_destroy.argtypes = [vp]
    L.orbx_destroy.restype = None
    L.orbx_last_error.argtypes = [vp]
    L.orbx_last_error.restype = C.c_char_p
    L.orbx_keypoint_capacity.argtypes = [vp]
    L.orbx_get_tables.argtypes = [vp, fp, fp, fp, fp, ip]
    L.orbx_get_level_sizes.argtypes = [vp, C.c_int, C.c_int, ip, ip]
    L.orbx_extract.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, ip, ip]
    L.orbx_extract_batch.argtypes = [vp, C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp,
                                     C.c_int, vp, vp]
    L.orbx_extract_batch_submit.argtypes = [vp, C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp,
                                            C.c_int]
    L.orbx_extract_batch_collect.argtypes = [vp, vp, vp]
    L.orbx_wire_parse_frame.argtypes = [vp, C.c_size_t, C.POINTER(_WireFrame)]
    L.orbx_wire_process_frame.argtypes = [vp, vp, C.c_size_t, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, ip, ip, ip, ip,
                                          C.POINTER(C.c_double), ip]
    L.orbx_wire_features_bound.argtypes = [C.c_int, C.c_int]
    L.orbx_wire_features_bound.restype = C.c_size_t
    L.orbx_wire_pack_features.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int, vp, C.c_size_t,
                                          C.POINTER(C.c_size_t)]
    L.orbx_wire_parse_features.argtypes = [vp, C.c_size_t, C.POINTER(_WireFeatures)]
    L.orbx_pnm_header.argtypes = [vp, C.c_size_t, ip, ip, ip, C.POINTER(C.c_size_t)]
    L.orbx_extract_pnm.argtypes = [vp, vp, C.c_size_t, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, ip, ip, ip, ip]
    L.orbx_extract_batch_device.argtypes = [vp, vp, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp,
                                            vp, C.c_int, vp, vp]
    L.orbx_sync.argtypes = [vp]
    L.orbx_launch_count.argtypes = [vp]
    L.orbx_launch_count.restype = C.c_longlong
    L.orbx_debug_get_level.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, ip, ip]
    L.orbx_debug_get_candidates.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int, ip]
    L.orbx_debug_get_level_keypoints.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int, ip]
    L.orbx_debug_resize.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int]
    L.orbx_debug_blur.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int]
    L.orbx_set_input_format.argtypes = [vp, C.c_int, C.c_int]
    L.orbx_undistort_points.argtypes = [vp, vp, C.c_int, vp, vp]
    L.orbx_image_bounds.argtypes = [vp, vp, C.c_int, C.c_int, vp]
    L.orbx_frame_grid.argtypes = [vp, vp, C.c_int, vp, vp, vp, vp, vp]
    L.orbx_frame_grid_batch_device.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp]
    L.orbx_match_windowed_grid_device.argtypes = [vp, vp, vp, vp, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.orbx_match_windowed_grid_batch_device.argtypes = [vp, C.c_int, vp, vp, C.c_int, C.c_int] + [vp] * 13
    L.orbx_vocab_create.argtypes = [C.c_int, vp, vp, vp, C.c_int, C.POINTER(vp)]
    L.orbx_vocab_destroy.argtypes = [vp]
    L.orbx_vocab_last_error.restype = C.c_char_p
    L.orbx_vocab_last_error.argtypes = [vp]
    L.orbx_vocab_depth.argtypes = [vp]
    L.orbx_vocab_transform.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp, vp]
    L.orbx_debug_gray.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int]
    L.orbx_debug_octree.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int, ip]
    L.orbx_debug_describe.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp, vp, vp]
    L.orbx_distance_batch.argtypes = [vp, vp, vp, C.c_int, vp]
    L.orbx_match_windowed.argtypes = [vp, vp, vp, vp, C.c_int, vp, vp, C.c_int, vp, vp, vp, vp, vp]
    L.orbx_knn2_create_db.argtypes = [C.c_int, vp, C.c_longlong, C.c_longlong, C.POINTER(vp)]
    L.orbx_knn2_create_db_device.argtypes = [C.c_int, vp, C.c_longlong, C.c_longlong, C.POINTER(vp)]
    L.orbx_knn2_destroy_db.argtypes = [vp]
    L.orbx_knn2_destroy_db.restype = None
    L.orbx_knn2_last_error.argtypes = [vp]
    L.orbx_knn2_last_error.restype = C.c_char_p
    L.orbx_knn2_query.argtypes = [vp, vp, C.c_int, vp, vp]
    L.orbx_knn2_query_device.argtypes = [vp, vp, C.c_int, vp]
    L.orbx_knn2_merge_device.argtypes = [vp, vp, C.c_int, C.c_int, vp]
    L.orbx_knn2_sync.argtypes = [vp]
    L.orbx_knn2_launch_count.argtypes = [vp]
    L.orbx_knn2_launch_count.restype = C.c_longlong
    L.orbx_plan_probe.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, ip, ip, ip, ip, ip, ip,
                                  C.POINTER(C.c_longlong)]
    L.orbx_version.restype = C.c_char_p
    L.orbx_host_alloc.argtypes = [C.c_size_t, C.c_int]
    L.orbx_host_alloc.restype = vp
    L.orbx_host_free.argtypes = [vp]
    L.orbx_host_free.restype = None
    L.orbx_comm_unique_id.argtypes = [vp]
    L.orbx_comm_create.argtypes = [C.c_int, C.c_int, C.c_int, vp, C.POINTER(vp)]
    L.orbx_comm_adopt.argtypes = [vp, C.c_int, C.c_int, C.POINTER(vp)]
    L.orbx_comm_destroy.argtypes = [vp]
    L.orbx_comm_destroy.restype = None
    L.orbx_comm_last_error.argtypes = [vp]
    L.orbx_comm_last_error.restype = C.c_char_p
    L.orbx_knn2_query_sharded_device.argtypes = [vp, vp, vp, C.c_int, vp]
    L.orbx_knn2_query_sharded.argtypes = [vp, vp, vp, C.c_int, vp, vp]
    L.orbx_wire_copy_keypoints.argtypes = [C.POINTER(_WireFeatures), vp, C.c_int]
    L.orbx_set_stream.argtypes = [vp, vp]
    L.orbx_knn2_set_stream.argtypes = [vp, vp]
    L.orbx_knn2_set_backend.argtypes = [vp, C.c_int]
    L.orbx_set_profiling.argtypes = [vp, C.c_int]
    L.orbx_get_stage_times.argtypes = [vp, fp]
    _lib = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def plan_probe(nfeatures, scale_factor, nlevels, ini_th, min_th, width, height):
    """Geometry plan of the library for one frame size; needs no GPU."""
    L = lib()
    arrs = [np.zeros(nlevels, np.int32) for _ in range(6)]
    ab = C.c_longlong()
    ip = C.POINTER(C.c_int)
    rc = L.orbx_plan_probe(nfeatures, scale_factor, nlevels, ini_th, min_th, width, height,
                           *[a.ctypes.data_as(ip) for a in arrs], C.byref(ab))
    if rc < 0:
        raise OrbxError(rc, (L.orbx_last_error(None) or b"").decode())
    keys = ["widths", "heights", "ncells", "quota", "n_ini", "depth0"]
    out = dict(zip(keys, arrs))
    out["algorithmic_bytes"] = ab.value
    return out


class ORBextractor:
    """ORB_SLAM3::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST) on one B200.

    Extra keyword arguments size the GPU workspace (device ordinal, largest frame, frames per batch call)."""

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, *, device=0, max_width=1920,
                 max_height=1080, max_batch=1):
        self._L = lib()
        self._h = C.c_void_p()
        self._fmt = FMT_GRAY8
        cfg = _Config(int(nfeatures), float(scaleFactor), int(nlevels), int(iniThFAST), int(minThFAST), int(device),
                      int(max_width), int(max_height), int(max_batch))
        rc = self._L.orbx_create(C.byref(cfg), C.byref(self._h))
        if rc != ORBX_OK:
            raise OrbxError(rc, (self._L.orbx_last_error(None) or b"").decode())
        self.nfeatures, self.nlevels, self.max_batch, self.device = int(nfeatures), int(nlevels), int(max_batch), int(device)
        n = self.nlevels
        self._scale, self._inv, self._s2, self._is2 = (np.zeros(n, np.float32) for _ in range(4))
        self._quota = np.zeros(n, np.int32)
        fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
        self._L.orbx_get_tables(self._h, self._scale.ctypes.data_as(fp), self._inv.ctypes.data_as(fp),
                                self._s2.ctypes.data_as(fp), self._is2.ctypes.data_as(fp), self._quota.ctypes.data_as(ip))
        self.capacity = int(self._L.orbx_keypoint_capacity(self._h))
        self.mvImagePyramid = []  # as in the shim: left empty for mono (only stereo matching reads it)

    # -- lifecycle
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.orbx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != ORBX_OK:
            raise OrbxError(rc, (self._L.orbx_last_error(self._h) or b"").decode())

    # -- getters of include/ORBextractor.h
    def GetLevels(self):
        return self.nlevels

    def GetScaleFactor(self):
        return float(self._scale[1]) if self.nlevels > 1 else 1.0

    def GetScaleFactors(self):
        return self._scale.copy()

    def GetInverseScaleFactors(self):
        return self._inv.copy()

    def GetScaleSigmaSquares(self):
        return self._s2.copy()

    def GetInverseScaleSigmaSquares(self):
        return self._is2.copy()

    def features_per_level(self):
        return self._quota.copy()

    def level_sizes(self, width, height):
        w, h = np.zeros(self.nlevels, np.int32), np.zeros(self.nlevels, np.int32)
        ip = C.POINTER(C.c_int)
        self._L.orbx_get_level_sizes(self._h, width, height, w.ctypes.data_as(ip), h.ctypes.data_as(ip))
        return list(zip(w.tolist(), h.tolist()))

    # -- operator()
    def __call__(self, image, mask=None, vLappingArea=(0, 1000)):
        """Returns (monoIndex, keypoints, descriptors).  Empty image: (-1, empty, empty) like the reference."""
        if image is None or image.size == 0:
            return -1, np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        bpp = {FMT_GRAY8: 1, FMT_RGB8: 3, FMT_BGR8: 3, FMT_RGBA8: 4, FMT_BGRA8: 4}[self._fmt]
        if image.dtype != np.uint8 or (bpp == 1 and image.ndim != 2) or (bpp > 1 and (image.ndim != 3 or image.shape[2] != bpp)):
            raise OrbxError(ORBX_E_INVALID, "image must be CV_8UC1 (the reference asserts image.type() == CV_8UC1), or "
                                            "[H,W,3|4] after set_input_format")
        if image.strides[1] != bpp or (bpp > 1 and image.strides[2] != 1):
            image = np.ascontiguousarray(image)
        h, w = image.shape[:2]
        kps = np.zeros(self.capacity, KP_DTYPE)
        desc = np.zeros((self.capacity, 32), np.uint8)
        n, mono = C.c_int(), C.c_int()
        rc = self._L.orbx_extract(self._h, _p(image), w, h, image.strides[0], int(vLappingArea[0]), int(vLappingArea[1]),
                                  _p(kps), _p(desc), self.capacity, C.byref(n), C.byref(mono))
        self._check(rc)
        return mono.value, kps[:n.value].copy(), desc[:n.value].copy()

    def extract_pnm(self, data: bytes, camera_rgb: bool = True, vLappingArea=(0, 1000)):
        """A frame as it arrives on the reference's wire (binary PPM / PGM): cv::imdecode + cvtColor + operator().
        Returns (monoIndex, keypoints, descriptors, (width, height)); (-1, empty, empty, None) where imdecode fails."""
        buf = np.frombuffer(data, np.uint8)
        kps = np.zeros(self.capacity, KP_DTYPE)
        desc = np.zeros((self.capacity, 32), np.uint8)
        n, mono, w, h = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        rc = self._L.orbx_extract_pnm(self._h, _p(buf) if len(buf) else None, len(buf), int(bool(camera_rgb)), int(vLappingArea[0]),
                                      int(vLappingArea[1]), _p(kps), _p(desc), self.capacity, C.byref(n), C.byref(mono),
                                      C.byref(w), C.byref(h))
        if rc == ORBX_E_EMPTY:
            return -1, np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8), None
        self._check(rc)
        return mono.value, kps[:n.value].copy(), desc[:n.value].copy(), (w.value, h.value)

    def process_frame_message(self, payload: bytes, camera_rgb: bool = True, vLappingArea=(0, 1000)):
        """One 'frame' message as it comes off the socket (the MessagePack map after the 4-byte length): what the backend's receive
        loop does up to operator().  Returns dict(mono_index, keypoints, descriptors, size, timestamp, camera_id), or None where the
        reference logs and skips the message; raises OrbxError where its ParseMessage throws / the message is not a frame."""
        buf = np.frombuffer(payload, np.uint8)
        kps = np.zeros(self.capacity, KP_DTYPE)
        desc = np.zeros((self.capacity, 32), np.uint8)
        n, mono, w, h, cam, ts = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_double()
        rc = self._L.orbx_wire_process_frame(self._h, _p(buf) if len(buf) else None, len(buf), int(bool(camera_rgb)), int(vLappingArea[0]),
                                             int(vLappingArea[1]), _p(kps), _p(desc), self.capacity, C.byref(n), C.byref(mono),
                                             C.byref(w), C.byref(h), C.byref(ts), C.byref(cam))
        if rc == ORBX_E_EMPTY:
            return None
        self._check(rc)
        return {"mono_index": mono.value, "keypoints": kps[:n.value].copy(), "descriptors": desc[:n.value].copy(),
                "size": (w.value, h.value), "timestamp": ts.value, "camera_id": cam.value}

    def _batch_args(self, frames, out):
        frames = np.asarray(frames)
        bpp = {FMT_GRAY8: 1, FMT_RGB8: 3, FMT_BGR8: 3, FMT_RGBA8: 4, FMT_BGRA8: 4}[self._fmt]
        if frames.dtype != np.uint8 or frames.ndim != (3 if bpp == 1 else 4) or (bpp > 1 and frames.shape[3] != bpp):
            raise OrbxError(ORBX_E_INVALID, "frames must be uint8 [B,H,W] (or [B,H,W,3|4] after set_input_format)")
        if frames.strides[2] != bpp or (bpp > 1 and frames.strides[3] != 1):
            frames = np.ascontiguousarray(frames)
        B, h, w = frames.shape[:3]
        base, step = frames.ctypes.data, frames.strides[0]
        ptrs = (C.c_void_p * B)(*range(base, base + B * step, step)) if step > 0 else (C.c_void_p * B)(*[base] * B)
        if out is None:
            kps = np.zeros((B, self.capacity), KP_DTYPE)
            desc = np.zeros((B, self.capacity, 32), np.uint8)
        else:
            kps, desc = out
            if kps.shape != (B, self.capacity) or desc.shape != (B, self.capacity, 32) or kps.dtype != KP_DTYPE or desc.dtype != np.uint8:
                raise OrbxError(ORBX_E_INVALID, "out arrays must be [B,cap] KP_DTYPE and [B,cap,32] uint8")
        return frames, ptrs, B, w, h, kps, desc

    def extract_batch(self, frames, vLappingArea=(0, 1000), out=None):
        """frames: uint8 [B,H,W] (C-contiguous rows).  Returns (mono[B], n[B], kps[B,cap], desc[B,cap,32]).
        out = (kps, desc) lets the caller supply (e.g. page-locked) result arrays of shape [B,cap] / [B,cap,32]."""
        frames, ptrs, B, w, h, kps, desc = self._batch_args(frames, out)
        n, mono = np.zeros(B, np.int32), np.zeros(B, np.int32)
        rc = self._L.orbx_extract_batch(self._h, ptrs, B, w, h, frames.strides[1], int(vLappingArea[0]),
                                        int(vLappingArea[1]), _p(kps), _p(desc), self.capacity, _p(n), _p(mono))
        self._check(rc)
        return mono, n, kps, desc

    def extract_batch_submit(self, frames, vLappingArea=(0, 1000), out=None):
        """Queue one batch (orbx_extract_batch_submit) and return without waiting; extract_batch_collect() returns the
        results.  The frame and result arrays are kept alive by the extractor until then."""
        frames, ptrs, B, w, h, kps, desc = self._batch_args(frames, out)
        rc = self._L.orbx_extract_batch_submit(self._h, ptrs, B, w, h, frames.strides[1], int(vLappingArea[0]),
                                               int(vLappingArea[1]), _p(kps), _p(desc), self.capacity)
        self._check(rc)
        self._inflight = (frames, ptrs, B, kps, desc)

    def extract_batch_collect(self):
        """Wait for the submitted batch: (mono[B], n[B], kps[B,cap], desc[B,cap,32])."""
        if getattr(self, "_inflight", None) is None:
            raise OrbxError(ORBX_E_INVALID, "no submitted batch to collect")
        frames, ptrs, B, kps, desc = self._inflight
        self._inflight = None
        n, mono = np.zeros(B, np.int32), np.zeros(B, np.int32)
        self._check(self._L.orbx_extract_batch_collect(self._h, _p(n), _p(mono)))
        return mono, n, kps, desc

    def extract_batch_device(self, d_frames_ptr, frame_stride, batch, width, height, stride, d_kp_ptr, d_desc_ptr, cap,
                             d_n_ptr, d_mono_ptr, vLappingArea=(0, 1000)):
        """Device-resident batch: raw device pointers (e.g. torch.Tensor.data_ptr()); asynchronous, see sync()."""
        rc = self._L.orbx_extract_batch_device(self._h, C.c_void_p(d_frames_ptr), frame_stride, batch, width, height, stride,
                                               int(vLappingArea[0]), int(vLappingArea[1]), C.c_void_p(d_kp_ptr),
                                               C.c_void_p(d_desc_ptr), cap, C.c_void_p(d_n_ptr), C.c_void_p(d_mono_ptr))
        self._check(rc)

    def sync(self):
        self._check(self._L.orbx_sync(self._h))

    def set_input_format(self, fmt: int, gray_shift: int = 15):
        """FMT_GRAY8 (default) or FMT_RGB8 / FMT_BGR8 / FMT_RGBA8 / FMT_BGRA8: colour frames are converted on the device
        (cv::cvtColor *2GRAY of UPSTREAM Tracking::GrabImageMonocular); gray_shift 15 = cv2 4.13 fixed point, 14 = older."""
        self._check(self._L.orbx_set_input_format(self._h, int(fmt), int(gray_shift)))
        self._fmt = int(fmt)

    # -- Frame post-extraction steps (Frame::UndistortKeyPoints / ComputeImageBounds / AssignFeaturesToGrid)
    GRID_COLS, GRID_ROWS = 64, 48

    @staticmethod
    def _camera(cam):
        """(fx, fy, cx, cy, k1, k2, p1, p2[, k3]) -> orbx_camera"""
        c = np.zeros(9, np.float32)
        c[:len(cam)] = np.asarray(cam, np.float32)
        return c

    def undistort_points(self, xy, cam):
        xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
        out = np.zeros_like(xy)
        c = self._camera(cam)
        self._check(self._L.orbx_undistort_points(self._h, _p(xy), len(xy), _p(c), _p(out)))
        return out

    def image_bounds(self, cam, width, height):
        b = np.zeros(4, np.float32)
        c = self._camera(cam)
        self._check(self._L.orbx_image_bounds(self._h, _p(c), int(width), int(height), _p(b)))
        return b

    def frame_grid(self, kps, cam, bounds):
        """Returns (mvKeysUn, cell_start[3073], cell_items): the feature grid as CSR, cell = posX * 48 + posY."""
        kps = np.ascontiguousarray(kps, KP_DTYPE)
        un = np.zeros_like(kps)
        start = np.zeros(self.GRID_COLS * self.GRID_ROWS + 1, np.int32)
        items = np.zeros(max(len(kps), 1), np.int32)
        c, b = self._camera(cam), np.ascontiguousarray(bounds, np.float32)
        self._check(self._L.orbx_frame_grid(self._h, _p(kps), len(kps), _p(c), _p(b), _p(un), _p(start), _p(items)))
        return un, start, items[:start[-1]]

    def frame_grid_batch_device(self, d_kp_ptr, d_n_ptr, batch, cap, cam, bounds, d_kp_un_ptr, d_cell_start_ptr, d_cell_items_ptr):
        c, b = self._camera(cam), np.ascontiguousarray(bounds, np.float32)
        self._check(self._L.orbx_frame_grid_batch_device(self._h, C.c_void_p(d_kp_ptr), C.c_void_p(d_n_ptr), int(batch), int(cap), _p(c),
                                                         _p(b), C.c_void_p(d_kp_un_ptr), C.c_void_p(d_cell_start_ptr),
                                                         C.c_void_p(d_cell_items_ptr)))

    def match_windowed_grid_device(self, d_q_desc, d_q_uvr, d_q_levels, nq, d_t_kp_un, d_t_desc, d_cell_start, d_cell_items, bounds,
                                   d_best_idx, d_best_dist, d_second_idx, d_second_dist):
        """Windowed search on device-resident data through the feature grid (raw device pointers); asynchronous, see sync()."""
        b = np.ascontiguousarray(bounds, np.float32)
        p = [C.c_void_p(x) for x in (d_q_desc, d_q_uvr, d_q_levels)]
        t = [C.c_void_p(x) for x in (d_t_kp_un, d_t_desc, d_cell_start, d_cell_items)]
        o = [C.c_void_p(x) for x in (d_best_idx, d_best_dist, d_second_idx, d_second_dist)]
        self._check(self._L.orbx_match_windowed_grid_device(self._h, p[0], p[1], p[2], int(nq), t[0], t[1], t[2], t[3], _p(b), *o))

    def match_windowed_grid_batch_device(self, pairs, batch, cap, d_q_desc, d_q_uvr, d_q_levels, d_n, d_t_kp_un, d_t_desc, d_cell_start,
                                         d_cell_items, bounds, d_best_idx, d_best_dist, d_second_idx, d_second_dist):
        """pairs: sequence of (query frame, train frame) inside one batch; every device array is [batch][cap] (raw pointers); one launch
        per 64 pairs, asynchronous, see sync()."""
        b = np.ascontiguousarray(bounds, np.float32)
        pq = np.ascontiguousarray([p[0] for p in pairs], np.int32)
        pt = np.ascontiguousarray([p[1] for p in pairs], np.int32)
        ptrs = [C.c_void_p(x) for x in (d_q_desc, d_q_uvr, d_q_levels, d_n, d_t_kp_un, d_t_desc, d_cell_start, d_cell_items)]
        o = [C.c_void_p(x) for x in (d_best_idx, d_best_dist, d_second_idx, d_second_dist)]
        self._check(self._L.orbx_match_windowed_grid_batch_device(self._h, len(pairs), _p(pq), _p(pt), int(batch), int(cap), *ptrs[:8], _p(b), *o))

    def debug_gray(self, src, fmt, gray_shift=15):
        src = np.ascontiguousarray(src, np.uint8)
        dst = np.zeros(src.shape[:2], np.uint8)
        self._check(self._L.orbx_debug_gray(self._h, _p(src), src.shape[1], src.shape[0], src.strides[0], int(fmt), int(gray_shift),
                                            _p(dst), dst.shape[1]))
        return dst

    def set_stream(self, cuda_stream: int):
        """cuda_stream: raw cudaStream_t (e.g. torch.cuda.Stream().cuda_stream); 0 = the handle's own stream.  torch's
        default stream also has handle 0: pass 1 (cudaStreamLegacy) to order the work with it."""
        self._check(self._L.orbx_set_stream(self._h, C.c_void_p(cuda_stream)))

    def launch_count(self):
        return int(self._L.orbx_launch_count(self._h))

    STAGES = ("pyramid", "blur", "fast", "quadtree", "finalize", "describe")

    def set_profiling(self, enable: bool):
        self._check(self._L.orbx_set_profiling(self._h, int(bool(enable))))

    def stage_times_ms(self):
        ms = np.zeros(6, np.float32)
        self._check(self._L.orbx_get_stage_times(self._h, ms.ctypes.data_as(C.POINTER(C.c_float))))
        return dict(zip(self.STAGES, ms.tolist()))

    # -- stage inspection (parity tests)
    def debug_level(self, frame, level, blurred=False):
        """Pyramid plane (or its blurred copy) of a frame of the last extract call."""
        w, h = C.c_int(), C.c_int()
        tmp = np.zeros((4095, 4096), np.uint8)   # the library checks out_stride >= level width
        rc = self._L.orbx_debug_get_level(self._h, frame, level, int(blurred), _p(tmp), 4096, C.byref(w), C.byref(h))
        self._check(rc)
        return tmp[:h.value, :w.value].copy()

    def debug_candidates(self, frame, level, cap=1 << 20):
        out = np.zeros((cap, 3), np.float32)
        n = C.c_int()
        self._check(self._L.orbx_debug_get_candidates(self._h, frame, level, _p(out), cap, C.byref(n)))
        return out[:n.value].copy()

    def debug_level_keypoints(self, frame, level, cap=1 << 16):
        out = np.zeros((cap, 4), np.float32)
        n = C.c_int()
        self._check(self._L.orbx_debug_get_level_keypoints(self._h, frame, level, _p(out), cap, C.byref(n)))
        return out[:n.value].copy()

    def debug_resize(self, src, dw, dh):
        src = np.ascontiguousarray(src, np.uint8)
        dst = np.zeros((dh, dw), np.uint8)
        self._check(self._L.orbx_debug_resize(self._h, _p(src), src.shape[1], src.shape[0], src.shape[1], _p(dst), dw, dh, dw))
        return dst

    def debug_blur(self, src):
        src = np.ascontiguousarray(src, np.uint8)
        dst = np.zeros_like(src)
        self._check(self._L.orbx_debug_blur(self._h, _p(src), src.shape[1], src.shape[0], src.shape[1], _p(dst), src.shape[1]))
        return dst

    def debug_octree(self, keys, minX, maxX, minY, maxY, N):
        keys = np.ascontiguousarray(keys, np.float32).reshape(-1, 3)
        out = np.zeros(max(len(keys), 64), np.int32)
        n = C.c_int()
        self._check(self._L.orbx_debug_octree(self._h, _p(keys), len(keys), minX, maxX, minY, maxY, N, _p(out), len(out), C.byref(n)))
        return out[:n.value].copy()

    def debug_describe(self, img, blurred, xy, angles=None):
        ref = img if img is not None else blurred
        h, w = ref.shape
        img_c = None if img is None else np.ascontiguousarray(img, np.uint8)
        bl_c = None if blurred is None else np.ascontiguousarray(blurred, np.uint8)
        xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
        n = len(xy)
        ang_in = None if angles is None else np.ascontiguousarray(angles, np.float32)
        ang_out = np.zeros(n, np.float32)
        desc = np.zeros((n, 32), np.uint8)
        self._check(self._L.orbx_debug_describe(self._h, _p(img_c), _p(bl_c), w, h, w, _p(xy), n, _p(ang_in), _p(ang_out), _p(desc)))
        return ang_out, desc

    # -- matching entry points that live on the extractor handle
    def distance_batch(self, a, b):
        a = np.ascontiguousarray(a, np.uint8).reshape(-1, 32)
        b = np.ascontiguousarray(b, np.uint8).reshape(-1, 32)
        if a.shape != b.shape:
            raise OrbxError(ORBX_E_INVALID, "descriptor arrays differ in shape")
        out = np.zeros(len(a), np.int32)
        self._check(self._L.orbx_distance_batch(self._h, _p(a), _p(b), len(a), _p(out)))
        return out

    def match_windowed(self, q_desc, q_uvr, q_levels, t_kp, t_desc, bounds):
        q_desc = np.ascontiguousarray(q_desc, np.uint8).reshape(-1, 32)
        q_uvr = np.ascontiguousarray(q_uvr, np.float32).reshape(-1, 3)
        q_levels = np.ascontiguousarray(q_levels, np.int32).reshape(-1, 2)
        t_kp = np.ascontiguousarray(t_kp, KP_DTYPE)
        t_desc = np.ascontiguousarray(t_desc, np.uint8).reshape(-1, 32)
        bounds = np.ascontiguousarray(bounds, np.float32)
        nq = len(q_desc)
        outs = [np.zeros(nq, np.int32) for _ in range(4)]
        self._check(self._L.orbx_match_windowed(self._h, _p(q_desc), _p(q_uvr), _p(q_levels), nq, _p(t_kp), _p(t_desc), len(t_kp),
                                                _p(bounds), *[_p(o) for o in outs]))
        return tuple(outs)  # best_idx, best_dist, second_idx, second_dist


class ORBmatcher:
    """The arithmetic core of ORB_SLAM3::ORBmatcher: DescriptorDistance and the windowed candidate search.
    TH_HIGH / TH_LOW / HISTO_LENGTH and the nnratio ctor argument are kept for the callers that gate on them."""
    TH_HIGH, TH_LOW, HISTO_LENGTH = 100, 50, 30

    def __init__(self, nnratio=0.6, checkOri=True, *, extractor: ORBextractor):
        self.mfNNratio, self.mbCheckOrientation, self._ex = float(nnratio), bool(checkOri), extractor

    def DescriptorDistance(self, a, b) -> int:
        return int(self._ex.distance_batch(np.asarray(a).reshape(1, 32), np.asarray(b).reshape(1, 32))[0])

    def DescriptorDistances(self, a, b):
        return self._ex.distance_batch(a, b)

    def SearchInWindows(self, q_desc, q_uvr, q_levels, t_kp, t_desc, bounds):
        """GetFeaturesInArea + best / second-best per query (the loop body of SearchByProjection /
        SearchForInitialization); thresholds, ratio test and rotation histogram stay with the caller."""
        return self._ex.match_windowed(q_desc, q_uvr, q_levels, t_kp, t_desc, bounds)


class Knn2Index:
    """One row shard of a descriptor database resident in HBM: cv::BFMatcher(NORM_HAMMING).knnMatch(k=2)."""

    def __init__(self, rows=None, *, device=0, row_offset=0, device_ptr=None, nrows=None):
        self._L = lib()
        self._db = C.c_void_p()
        if device_ptr is not None:
            rc = self._L.orbx_knn2_create_db_device(device, C.c_void_p(device_ptr), int(nrows), int(row_offset), C.byref(self._db))
            self.nrows = int(nrows)
        else:
            rows = np.ascontiguousarray(rows, np.uint8).reshape(-1, 32)
            rc = self._L.orbx_knn2_create_db(device, _p(rows), len(rows), int(row_offset), C.byref(self._db))
            self.nrows = len(rows)
        if rc != ORBX_OK:
            raise OrbxError(rc, (self._L.orbx_knn2_last_error(None) or b"").decode())

    def close(self):
        if getattr(self, "_db", None) and self._db.value:
            self._L.orbx_knn2_destroy_db(self._db)
            self._db = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != ORBX_OK:
            raise OrbxError(rc, (self._L.orbx_knn2_last_error(self._db) or b"").decode())

    def knnMatch(self, queries):
        """Returns (idx [nq,2] int32 global rows, dist [nq,2] int32); -1 where the shard has fewer than 2 rows."""
        q = np.ascontiguousarray(queries, np.uint8).reshape(-1, 32)
        idx, dist = np.zeros((len(q), 2), np.int32), np.zeros((len(q), 2), np.int32)
        self._check(self._L.orbx_knn2_query(self._db, _p(q), len(q), _p(idx), _p(dist)))
        return idx, dist

    def query_device(self, d_queries_ptr, nq, d_packed_out_ptr):
        self._check(self._L.orbx_knn2_query_device(self._db, C.c_void_p(d_queries_ptr), nq, C.c_void_p(d_packed_out_ptr)))

    def merge_device(self, d_partials_ptr, nparts, nq, d_packed_out_ptr):
        self._check(self._L.orbx_knn2_merge_device(self._db, C.c_void_p(d_partials_ptr), nparts, nq, C.c_void_p(d_packed_out_ptr)))

    def sync(self):
        self._check(self._L.orbx_knn2_sync(self._db))

    def set_stream(self, cuda_stream: int):
        self._check(self._L.orbx_knn2_set_stream(self._db, C.c_void_p(cuda_stream)))

    def query_sharded_device(self, comm: "Comm", d_queries_ptr, nq, d_packed_out_ptr):
        """Top-2 over the WHOLE row-sharded database on every rank: local query + ncclAllGather + merge on this shard's stream (collective)."""
        self._check(self._L.orbx_knn2_query_sharded_device(self._db, comm._c, C.c_void_p(d_queries_ptr), nq, C.c_void_p(d_packed_out_ptr)))

    def knnMatch_sharded(self, comm: "Comm", queries):
        q = np.ascontiguousarray(queries, np.uint8).reshape(-1, 32)
        idx, dist = np.zeros((len(q), 2), np.int32), np.zeros((len(q), 2), np.int32)
        self._check(self._L.orbx_knn2_query_sharded(self._db, comm._c, _p(q), len(q), _p(idx), _p(dist)))
        return idx, dist

    POPC, TENSOR, TENSOR_FP4 = 0, 1, 2

    def set_backend(self, backend: int):
        """Knn2Index.TENSOR (default: tcgen05 int8 GEMM tiles) or Knn2Index.POPC (CUDA-core XOR + POPC)."""
        self._check(self._L.orbx_knn2_set_backend(self._db, int(backend)))

    def launch_count(self):
        return int(self._L.orbx_knn2_launch_count(self._db))


class PinnedArray:
    """A numpy array in page-locked host memory from orbx_host_alloc (optionally write-combined, for upload-only frame buffers)."""

    def __init__(self, shape, dtype=np.uint8, write_combined=False):
        self._L = lib()
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self._p = self._L.orbx_host_alloc(self.nbytes, 1 if write_combined else 0)
        if not self._p:
            raise OrbxError(ORBX_E_CUDA, "cudaHostAlloc failed")
        buf = (C.c_uint8 * self.nbytes).from_address(self._p)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def close(self):
        if getattr(self, "_p", None):
            self.array = None
            self._L.orbx_host_free(C.c_void_p(self._p))
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


NCCL_ID_BYTES = 128


def comm_unique_id() -> bytes:
    """ncclGetUniqueId through the library (call on ONE rank, ship the bytes to the others)."""
    buf = (C.c_uint8 * NCCL_ID_BYTES)()
    rc = lib().orbx_comm_unique_id(buf)
    if rc != ORBX_OK:
        raise OrbxError(rc, (lib().orbx_comm_last_error(None) or b"").decode())
    return bytes(buf)


class Comm:
    """One NCCL communicator of this process (one rank = one GPU) for the row-sharded kNN: orbx_comm of include/orbx.h."""

    def __init__(self, device: int, rank: int, nranks: int, unique_id: bytes):
        self._L = lib()
        self._c = C.c_void_p()
        idb = (C.c_uint8 * NCCL_ID_BYTES).from_buffer_copy(unique_id)
        rc = self._L.orbx_comm_create(device, rank, nranks, idb, C.byref(self._c))
        if rc != ORBX_OK:
            raise OrbxError(rc, (self._L.orbx_comm_last_error(None) or b"").decode())
        self.rank, self.nranks = rank, nranks

    def close(self):
        if getattr(self, "_c", None) and self._c.value:
            self._L.orbx_comm_destroy(self._c)
            self._c = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def unpack_knn(packed: np.ndarray):
    """(dist << 32 | row) uint64 [nq,2] -> (idx int64, dist int32), -1 where missing."""
    packed = np.asarray(packed, np.uint64).reshape(-1, 2)
    none = packed == np.uint64(0xFFFFFFFFFFFFFFFF)
    idx = (packed & np.uint64(0xFFFFFFFF)).astype(np.int64)
    dist = (packed >> np.uint64(32)).astype(np.int64).astype(np.int32)
    idx[none] = -1
    dist[none] = -1
    return idx, dist


class ORBVocabulary:
    """DBoW2 vocabulary tree on one B200: transform() = the per-descriptor descent of TemplatedVocabulary::transform that UPSTREAM
    Frame::ComputeBoW runs (word id, word weight, node id `levelsup` levels above the leaves).  The tree is given as arrays in node
    order (node 0 = root, parents before children -- the order of an ORBvoc.txt file)."""

    def __init__(self, parent, desc, weight, device=0):
        self._L = lib()
        self._v = C.c_void_p()
        parent = np.ascontiguousarray(parent, np.int32); desc = np.ascontiguousarray(desc, np.uint8)
        weight = np.ascontiguousarray(weight, np.float32)
        if desc.shape != (len(parent), 32) or weight.shape != (len(parent),):
            raise OrbxError(ORBX_E_INVALID, "desc must be [n,32] uint8 and weight [n] float32")
        rc = self._L.orbx_vocab_create(int(device), _p(parent), _p(desc), _p(weight), len(parent), C.byref(self._v))
        if rc != ORBX_OK:
            raise OrbxError(rc, (self._L.orbx_vocab_last_error(None) or b"").decode())
        self.depth = int(self._L.orbx_vocab_depth(self._v))

    def transform(self, desc, levelsup=4):
        desc = np.ascontiguousarray(desc, np.uint8)
        n = len(desc)
        w, wt, nid = np.zeros(n, np.int32), np.zeros(n, np.float32), np.zeros(n, np.int32)
        rc = self._L.orbx_vocab_transform(self._v, _p(desc), n, int(levelsup), _p(w), _p(wt), _p(nid))
        if rc != ORBX_OK:
            raise OrbxError(rc, (self._L.orbx_vocab_last_error(self._v) or b"").decode())
        return w, wt, nid

    def close(self):
        if self._v:
            self._L.orbx_vocab_destroy(self._v)
            self._v = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

"""Host-side mirror of the reference interface over the C ABI (include/rmcv_b200.h).

The reference's API for this path is three free functions (include/imgproc.h:29, include/objdetect.h:47-49,
70-71) returning the records of include/core.h:89-130.  This module binds librmcv_b200.so with ctypes and
offers

  * `Context` — one per host thread and GPU; batched device-resident calls (`detect_batch`,
    `extract_color_batch`, ...) and the host-buffer end-to-end call (`detect_batch_host`);
  * `extract_color`, `filter_lightblobs`, `filter_armours` — single-frame functions with the reference's
    names, argument order and meaning, returning numpy/cv2-shaped values so the parity tests read like the
    reference's call site at executable/main.cpp:172-176.

There is no CPU fallback: if the CUDA library is missing or no device is present every call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _abi as A

_LIB_NAME = "librmcv_b200.so"
_lib = None


class RmcvError(RuntimeError):
    def __init__(self, status: int, where: str, detail: str = ""):
        self.status = status
        super().__init__(f"{where}: status {status} ({_status_string(status)}) {detail}".strip())


def lib_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), _LIB_NAME)


def load_library():
    """Load librmcv_b200.so (built by __graft_entry__.build()).  Raises if it is missing: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not found - run `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback exists)")
    lib = C.CDLL(path)
    for name, (res, args) in A.PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.rmcv_abi_version() != A.ABI_VERSION:
        raise RuntimeError("librmcv_b200.so ABI version mismatch")
    _lib = lib
    return lib


def _status_string(s: int) -> str:
    try:
        return load_library().rmcv_status_string(s).decode()
    except Exception:
        return "?"


def default_params(**over) -> A.Params:
    p = A.Params()
    load_library().rmcv_default_params(C.byref(p))
    for k, v in over.items():
        if k == "ratio_range":
            p.ratio_min, p.ratio_max = v
        elif k == "area_range":
            p.area_min, p.area_max = v
        else:
            setattr(p, k, v)
    return p


# --------------------------------------------------------------------------------------------- records
@dataclass
class LightBlob:  # rm::lightblob, include/core.h:89-99
    angle: float
    target: int
    center: Tuple[float, float]
    vertices: np.ndarray  # 4x2 float32
    size: Tuple[float, float]

    @staticmethod
    def from_c(b: A.LightBlob) -> "LightBlob":
        v = np.array([[b.vertices[i][0], b.vertices[i][1]] for i in range(4)], np.float32)
        return LightBlob(float(b.angle), int(b.target), (float(b.center[0]), float(b.center[1])), v,
                         (float(b.size[0]), float(b.size[1])))

    def to_c(self) -> A.LightBlob:
        b = A.LightBlob()
        b.angle, b.target = self.angle, self.target
        b.center[0], b.center[1] = self.center
        for i in range(4):
            b.vertices[i][0], b.vertices[i][1] = float(self.vertices[i][0]), float(self.vertices[i][1])
        b.size[0], b.size[1] = self.size
        return b


@dataclass
class Armour:  # rm::armour public geometry, include/core.h:110-112
    icon: np.ndarray
    vertices: np.ndarray
    bounding_box: Tuple[float, float, float, float]
    i: int
    j: int
    gates: Tuple[float, ...]

    @staticmethod
    def from_c(a: A.Armour) -> "Armour":
        ic = np.array([[a.icon[k][0], a.icon[k][1]] for k in range(4)], np.float32)
        ve = np.array([[a.vertices[k][0], a.vertices[k][1]] for k in range(4)], np.float32)
        return Armour(ic, ve, tuple(float(x) for x in a.bounding_box), int(a.i), int(a.j), tuple(float(g) for g in a.gates))

    def to_c(self) -> A.Armour:
        a = A.Armour()
        for k in range(4):
            a.icon[k][0], a.icon[k][1] = float(self.icon[k][0]), float(self.icon[k][1])
            a.vertices[k][0], a.vertices[k][1] = float(self.vertices[k][0]), float(self.vertices[k][1])
        for k in range(4):
            a.bounding_box[k] = float(self.bounding_box[k])
        a.i, a.j = int(self.i), int(self.j)
        for k in range(6):
            a.gates[k] = float(self.gates[k]) if k < len(self.gates) else 0.0
        return a


@dataclass
class ContourInfo:
    first: Tuple[int, int]
    n_points: int
    status: int
    area2: int
    bbox: Tuple[int, int, int, int]
    ellipse: Tuple[float, float, float, float, float]
    fit_branch: int
    det0: float
    blob_index: int

    @staticmethod
    def from_c(c: A.ContourInfo) -> "ContourInfo":
        e = c.ellipse
        return ContourInfo((int(c.first_x), int(c.first_y)), int(c.n_points), int(c.status), int(c.area2),
                           tuple(int(v) for v in c.bbox), (float(e.cx), float(e.cy), float(e.w), float(e.h), float(e.angle)),
                           int(c.fit_branch), float(c.det0), int(c.blob_index))


@dataclass
class FrameDetections:
    contours: List[ContourInfo]
    positive: List[LightBlob]
    armours: List[Armour]
    n_negative: int
    flags: int
    poses: Optional[list] = None   # per armour (rvec, tvec, position, ok) when a camera is set (Context.set_camera)


# --------------------------------------------------------------------------------------------- context
class DeviceBuffer:
    def __init__(self, ctx: "Context", nbytes: int):
        self.ctx, self.nbytes = ctx, int(nbytes)
        p = C.c_void_p()
        ctx._check(ctx.lib.rmcv_device_alloc(ctx.h, self.nbytes, C.byref(p)), "rmcv_device_alloc")
        self.ptr = p.value

    def upload(self, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        self.ctx._check(self.ctx.lib.rmcv_memcpy_h2d(self.ctx.h, self.ptr, arr.ctypes.data, arr.nbytes), "rmcv_memcpy_h2d")
        self.ctx.sync()

    def download(self, shape, dtype=np.uint8) -> np.ndarray:
        out = np.empty(shape, dtype)
        assert out.nbytes <= self.nbytes
        self.ctx._check(self.ctx.lib.rmcv_memcpy_d2h(self.ctx.h, out.ctypes.data, self.ptr, out.nbytes), "rmcv_memcpy_d2h")
        self.ctx.sync()
        return out

    def free(self):
        if self.ptr:
            self.ctx.lib.rmcv_device_free(self.ctx.h, self.ptr)
            self.ptr = None


class PinnedArray:
    """numpy view over pinned host memory owned by the ctx."""

    def __init__(self, ctx: "Context", shape, dtype=np.uint8):
        self.ctx = ctx
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        ctx._check(ctx.lib.rmcv_host_alloc(ctx.h, n, C.byref(p)), "rmcv_host_alloc")
        self.ptr = p.value
        buf = (C.c_uint8 * n).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            self.ctx.lib.rmcv_host_free(self.ctx.h, self.ptr)
            self.ptr = None


class Context:
    """rmcv_ctx wrapper.  One per host thread and GPU."""

    def __init__(self, max_width=1280, max_height=1024, max_batch=64, device=0, chunk_frames=0, max_runs_per_frame=0,
                 max_blobs_per_frame=0, max_armours_per_frame=0, stream: Optional[int] = None):
        self.lib = load_library()
        cfg = A.Config()
        self.lib.rmcv_default_config(C.byref(cfg))
        cfg.device, cfg.max_width, cfg.max_height, cfg.max_batch = device, max_width, max_height, max_batch
        cfg.chunk_frames, cfg.max_runs_per_frame = chunk_frames, max_runs_per_frame
        cfg.max_blobs_per_frame, cfg.max_armours_per_frame = max_blobs_per_frame, max_armours_per_frame
        cfg.stream = stream
        h = C.c_void_p()
        rc = self.lib.rmcv_ctx_create(C.byref(cfg), C.byref(h))
        if rc != A.RMCV_OK:
            raise RmcvError(rc, "rmcv_ctx_create", "(is a CUDA device present? this library has no CPU path)")
        self.h = h
        self.cfg = cfg
        self.chunk_frames = int(self.lib.rmcv_chunk_frames(self.h))

    # -- plumbing
    def _check(self, rc: int, where: str):
        if rc != A.RMCV_OK:
            raise RmcvError(rc, where, self.lib.rmcv_last_error(self.h).decode(errors="replace"))

    def close(self):
        if getattr(self, "h", None):
            self.lib.rmcv_ctx_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self._check(self.lib.rmcv_sync(self.h), "rmcv_sync")

    def stream(self) -> int:
        return int(self.lib.rmcv_stream(self.h) or 0)

    def device_buffer(self, nbytes: int) -> DeviceBuffer:
        return DeviceBuffer(self, nbytes)

    def pinned(self, shape, dtype=np.uint8) -> PinnedArray:
        return PinnedArray(self, shape, dtype)

    def kernel_launches(self) -> int:
        return int(self.lib.rmcv_kernel_launches(self.h))

    def timer_start(self):
        self._check(self.lib.rmcv_timer_start(self.h), "rmcv_timer_start")

    def timer_stop(self) -> float:
        ms = C.c_double()
        self._check(self.lib.rmcv_timer_stop(self.h, C.byref(ms)), "rmcv_timer_stop")
        return ms.value

    def profile(self, on=True):
        self._check(self.lib.rmcv_profile_enable(self.h, 1 if on else 0), "rmcv_profile_enable")

    def profile_read(self, reset=True):
        ms = (C.c_double * len(A.STAGE_NAMES))()
        ln = (C.c_int64 * len(A.STAGE_NAMES))()
        self._check(self.lib.rmcv_profile_read(self.h, ms, ln, 1 if reset else 0), "rmcv_profile_read")
        return {n: (ms[i], int(ln[i])) for i, n in enumerate(A.STAGE_NAMES)}

    # -- device-resident batched calls (pointers are raw device addresses)
    def extract_color_batch(self, d_bgr: int, width: int, height: int, batch: int, target: int, lower_bound: int,
                            d_mask: Optional[int], pitch: Optional[int] = None, frame_stride: Optional[int] = None,
                            mask_pitch: Optional[int] = None, mask_frame_stride: Optional[int] = None):
        pitch = pitch or width * 3
        frame_stride = frame_stride or pitch * height
        mask_pitch = mask_pitch or width
        mask_frame_stride = mask_frame_stride or mask_pitch * height
        self._check(self.lib.rmcv_extract_color_batch(self.h, d_bgr, pitch, frame_stride, width, height, batch, target,
                                                      lower_bound, d_mask, mask_pitch, mask_frame_stride),
                    "rmcv_extract_color_batch")

    def bayer_extract_color_batch(self, d_raw: int, width: int, height: int, batch: int, layout: int, target: int,
                                  lower_bound: int, d_mask: Optional[int], pitch: Optional[int] = None,
                                  frame_stride: Optional[int] = None):
        pitch = pitch or width
        frame_stride = frame_stride or pitch * height
        self._check(self.lib.rmcv_bayer_extract_color_batch(self.h, d_raw, pitch, frame_stride, width, height, batch, layout,
                                                            target, lower_bound, d_mask, width, width * height),
                    "rmcv_bayer_extract_color_batch")

    def detect_batch(self, d_bgr: int, width: int, height: int, batch: int, params: A.Params, d_mask: Optional[int] = None,
                     pitch: Optional[int] = None, frame_stride: Optional[int] = None):
        pitch = pitch or width * 3
        frame_stride = frame_stride or pitch * height
        self._check(self.lib.rmcv_detect_batch(self.h, d_bgr, pitch, frame_stride, width, height, batch, C.byref(params),
                                               d_mask, width, width * height), "rmcv_detect_batch")

    def bayer_detect_batch(self, d_raw: int, width: int, height: int, batch: int, layout: int, params: A.Params,
                           d_mask: Optional[int] = None):
        self._check(self.lib.rmcv_bayer_detect_batch(self.h, d_raw, width, width * height, width, height, batch, layout,
                                                     C.byref(params), d_mask, width, width * height), "rmcv_bayer_detect_batch")

    def fetch_results(self, allow_overflow=False) -> A.Results:
        res = A.Results()
        rc = self.lib.rmcv_fetch_results(self.h, C.byref(res))
        if rc != A.RMCV_OK and not (allow_overflow and rc == A.RMCV_ERR_CAPACITY):
            self._check(rc, "rmcv_fetch_results")
        return res

    def detect_batch_host(self, frames: np.ndarray, params: A.Params, masks: Optional[np.ndarray] = None) -> A.Results:
        """frames: B×H×W×3 uint8 host array (pinned for full PCIe speed).  masks: optional B×H×W output."""
        assert frames.dtype == np.uint8 and frames.ndim == 4 and frames.shape[3] == 3 and frames.flags.c_contiguous
        B, H, W, _ = frames.shape
        res = A.Results()
        mptr = masks.ctypes.data if masks is not None else None
        self._check(self.lib.rmcv_detect_batch_host(self.h, frames.ctypes.data, W * 3, W * 3 * H, W, H, B, C.byref(params),
                                                    mptr, W, W * H, C.byref(res)), "rmcv_detect_batch_host")
        return res

    def bayer_detect_batch_host(self, raw: np.ndarray, layout: int, params: A.Params, masks: Optional[np.ndarray] = None) -> A.Results:
        """raw: B×H×W uint8 host mosaics (what the camera delivers, hardware/src/daheng.cpp:74-89); 1 B/px crosses PCIe."""
        assert raw.dtype == np.uint8 and raw.ndim == 3 and raw.flags.c_contiguous
        B, H, W = raw.shape
        res = A.Results()
        mptr = masks.ctypes.data if masks is not None else None
        self._check(self.lib.rmcv_bayer_detect_batch_host(self.h, raw.ctypes.data, W, W * H, W, H, B, int(layout), C.byref(params),
                                                          mptr, W, W * H, C.byref(res)), "rmcv_bayer_detect_batch_host")
        return res

    # -- result helpers
    @staticmethod
    def frame_detections(res: A.Results, f: int) -> FrameDetections:
        fi = res.frames[f]
        cs = [ContourInfo.from_c(res.contours[fi.contour_offset + k]) for k in range(fi.n_contours)]
        bs = [LightBlob.from_c(res.blobs[fi.blob_offset + k]) for k in range(fi.n_positive)]
        ar = [Armour.from_c(res.armours[fi.armour_offset + k]) for k in range(fi.n_armours)]
        poses = None
        if res.poses:
            poses = []
            for k in range(fi.n_armours):
                o = res.poses[fi.armour_offset + k]
                poses.append((np.array(o.rvec[:]), np.array(o.tvec[:]), np.array(o.position[:]), bool(o.ok)))
        return FrameDetections(cs, bs, ar, int(fi.n_negative), int(fi.flags), poses)

    def get_contour(self, frame: int, index: int, cap: int = 1 << 16) -> np.ndarray:
        n = C.c_int()
        buf = np.empty((cap, 2), np.int32)
        self._check(self.lib.rmcv_get_contour(self.h, frame, index, buf.ctypes.data, cap, C.byref(n)), "rmcv_get_contour")
        if n.value > cap:
            return self.get_contour(frame, index, n.value)
        return buf[:n.value].copy()

    def get_contours(self, frame: int) -> List[np.ndarray]:
        """All external contours of a frame, traced in one launch (list of N×2 int32 arrays, cv::findContours order)."""
        nc, npnt = C.c_int(), C.c_int()
        rc = self.lib.rmcv_get_contours(self.h, frame, None, 0, None, 0, C.byref(nc), C.byref(npnt))
        if rc not in (A.RMCV_OK, A.RMCV_ERR_CAPACITY):
            self._check(rc, "rmcv_get_contours")
        if nc.value == 0:
            return []
        xy = np.empty((max(npnt.value, 1), 2), np.int32)
        offs = np.empty(nc.value + 1, np.int32)
        self._check(self.lib.rmcv_get_contours(self.h, frame, xy.ctypes.data, npnt.value, offs.ctypes.data, nc.value,
                                               C.byref(nc), C.byref(npnt)), "rmcv_get_contours")
        return [xy[offs[k]:offs[k + 1]].copy() for k in range(nc.value)]

    def get_label_map(self, frame: int, width: int, height: int) -> np.ndarray:
        out = np.empty((height, width), np.int32)
        self._check(self.lib.rmcv_get_label_map(self.h, frame, out.ctypes.data, width), "rmcv_get_label_map")
        return out

    def get_bitmask(self, frame: int, width: int, height: int) -> np.ndarray:
        wb = (width + 31) // 32
        out = np.empty((height, wb), np.uint32)
        self._check(self.lib.rmcv_get_bitmask(self.h, frame, out.ctypes.data, wb), "rmcv_get_bitmask")
        return out

    # -- standalone a2..a5
    def filter_lightblobs_raw(self, contours: Sequence[np.ndarray], params: A.Params):
        n = len(contours)
        offs = np.zeros(n + 1, np.int32)
        for i, c in enumerate(contours):
            offs[i + 1] = offs[i] + len(c)
        xy = (np.concatenate([np.asarray(c, np.int32).reshape(-1, 2) for c in contours]) if n and offs[-1] > 0
              else np.zeros((0, 2), np.int32))
        xy = np.ascontiguousarray(xy, np.int32)
        infos = (A.ContourInfo * max(n, 1))()
        blobs = (A.LightBlob * max(n, 1))()
        nb = C.c_int()
        self._check(self.lib.rmcv_filter_lightblobs(self.h, xy.ctypes.data if xy.size else None, offs.ctypes.data, n,
                                                    C.byref(params), infos, blobs, max(n, 1), C.byref(nb)),
                    "rmcv_filter_lightblobs")
        return [ContourInfo.from_c(infos[i]) for i in range(n)], [LightBlob.from_c(blobs[i]) for i in range(nb.value)]

    def filter_armours_raw(self, blobs: Sequence[LightBlob], params: A.Params, cap: int = 4096) -> List[Armour]:
        n = len(blobs)
        arr = (A.LightBlob * max(n, 1))(*[b.to_c() for b in blobs])
        out = (A.Armour * cap)()
        na = C.c_int()
        self._check(self.lib.rmcv_filter_armours(self.h, arr, n, C.byref(params), out, cap, C.byref(na)), "rmcv_filter_armours")
        return [Armour.from_c(out[i]) for i in range(na.value)]

    def make_lightblobs(self, boxes: Sequence[Tuple[float, float, float, float, float]], target: int) -> List[LightBlob]:
        n = len(boxes)
        arr = (A.RotatedRect * max(n, 1))(*[A.RotatedRect(*b) for b in boxes])
        out = (A.LightBlob * max(n, 1))()
        self._check(self.lib.rmcv_make_lightblobs(self.h, arr, n, target, out), "rmcv_make_lightblobs")
        return [LightBlob.from_c(out[i]) for i in range(n)]


    # -- legacy rows (a6, a7)
    @staticmethod
    def _pack_contours(contours):
        n = len(contours)
        offs = np.zeros(n + 1, np.int32)
        for i, c in enumerate(contours):
            offs[i + 1] = offs[i] + len(c)
        xy = (np.concatenate([np.asarray(c, np.int32).reshape(-1, 2) for c in contours]) if n and offs[-1] > 0
              else np.zeros((0, 2), np.int32))
        return np.ascontiguousarray(xy, np.int32), offs

    def match_lightblobs(self, contours, min_ratio, max_ratio, tilt_angle, min_area, max_area, fit_ellipse=True):
        """rm::MatchLightBlob on every contour -> list of (ok, (cx, cy, w, h, angle))."""
        n = len(contours)
        if n == 0:
            return []
        xy, offs = self._pack_contours(contours)
        matched = np.zeros(n, np.int32)
        boxes = (A.RotatedRect * n)()
        self._check(self.lib.rmcv_match_lightblobs(self.h, xy.ctypes.data if xy.size else None, offs.ctypes.data, n, min_ratio,
                                                   max_ratio, tilt_angle, min_area, max_area, 1 if fit_ellipse else 0,
                                                   matched.ctypes.data, boxes), "rmcv_match_lightblobs")
        return [(bool(matched[k]), (boxes[k].cx, boxes[k].cy, boxes[k].w, boxes[k].h, boxes[k].angle)) for k in range(n)]

    def find_lightblobs_legacy(self, contours, min_ratio, max_ratio, tilt_angle, min_area, max_area, source, fit_ellipse=True):
        """rm::FindLightBlobs (legacy) -> list of LightBlob (camp voted from the source image)."""
        n = len(contours)
        if n == 0 or source.ndim != 3 or source.shape[2] != 3:   # src/objdetect.cpp:35
            return []
        source = np.ascontiguousarray(source, np.uint8)
        H, W, _ = source.shape
        xy, offs = self._pack_contours(contours)
        out = (A.LightBlob * n)()
        nb = C.c_int()
        self._check(self.lib.rmcv_find_lightblobs_legacy(self.h, xy.ctypes.data if xy.size else None, offs.ctypes.data, n,
                                                         min_ratio, max_ratio, tilt_angle, min_area, max_area, source.ctypes.data,
                                                         W * 3, W, H, 1 if fit_ellipse else 0, out, n, C.byref(nb)),
                    "rmcv_find_lightblobs_legacy")
        return [LightBlob.from_c(out[i]) for i in range(nb.value)]

    def min_area_rects(self, contours):
        """cv::minAreaRect of every contour -> list of (cx, cy, w, h, angle)."""
        n = len(contours)
        if n == 0:
            return []
        xy, offs = self._pack_contours(contours)
        boxes = (A.RotatedRect * n)()
        self._check(self.lib.rmcv_min_area_rects(self.h, xy.ctypes.data if xy.size else None, offs.ctypes.data, n, boxes),
                    "rmcv_min_area_rects")
        return [(boxes[k].cx, boxes[k].cy, boxes[k].w, boxes[k].h, boxes[k].angle) for k in range(n)]

    def lightblob_overlap(self, blobs: Sequence[LightBlob], left: int, right: int) -> bool:
        """rm::LightBlobOverlap."""
        n = len(blobs)
        arr = (A.LightBlob * max(n, 1))(*[b.to_c() for b in blobs])
        out = C.c_int()
        self._check(self.lib.rmcv_lightblob_overlap(self.h, arr, n, left, right, C.byref(out)), "rmcv_lightblob_overlap")
        return bool(out.value)


    # -- f4: camera front-end variants (hardware/src/daheng.cpp:91-187)
    def raw_frontend_batch(self, d_raw: int, width: int, height: int, batch: int, bits: int, mirror: bool, flip: bool,
                           d_raw8: int, pitch: Optional[int] = None, frame_stride: Optional[int] = None):
        bpp = 2 if bits > 8 else 1
        pitch = width * bpp if pitch is None else pitch
        frame_stride = pitch * height if frame_stride is None else frame_stride
        self._check(self.lib.rmcv_raw_frontend_batch(self.h, d_raw, pitch, frame_stride, width, height, batch, bits,
                                                     1 if mirror else 0, 1 if flip else 0, d_raw8, width, width * height),
                    "rmcv_raw_frontend_batch")

    def frontend_layout(self, layout: int, width: int, height: int, mirror: bool, flip: bool) -> int:
        return int(self.lib.rmcv_frontend_layout(layout, width, height, 1 if mirror else 0, 1 if flip else 0))

    # -- f1: rm::solve_PnP
    def set_camera(self, camera_matrix, dist_coeffs=None, exact_size=(27.0, 27.0), cam2world=None):
        """Fused rm::solve_PnP: every later detect call also fills Results.poses (executable/main.cpp:183-192)."""
        K = np.ascontiguousarray(np.asarray(camera_matrix, np.float64).reshape(9))
        D = None if dist_coeffs is None else np.ascontiguousarray(np.asarray(dist_coeffs, np.float64).reshape(5))
        M = None if cam2world is None else np.ascontiguousarray(np.asarray(cam2world, np.float64).reshape(16))
        self._check(self.lib.rmcv_set_camera(self.h, K.ctypes.data, None if D is None else D.ctypes.data, float(exact_size[0]),
                                             float(exact_size[1]), None if M is None else M.ctypes.data), "rmcv_set_camera")

    def clear_camera(self):
        self._check(self.lib.rmcv_clear_camera(self.h), "rmcv_clear_camera")

    def icon_batch(self, d_bgr: int, width: int, height: int, armours: Sequence[Armour], out_size=(20, 20), pitch: Optional[int] = None):
        """rm::affine_correction + flatten_image for every armour of a device-resident frame ->
        (icons uint8 n x h x w x 3, rows float32 n x (w*h*3), armours with their icon vertices clamped into the frame)."""
        n = len(armours)
        ow, oh = int(out_size[0]), int(out_size[1])
        icons = np.zeros((n, oh, ow, 3), np.uint8)
        rows = np.zeros((n, ow * oh * 3), np.float32)
        if n == 0:
            return icons, rows, []
        arr = (A.Armour * n)(*[a.to_c() for a in armours])
        self._check(self.lib.rmcv_icon_batch(self.h, d_bgr, pitch or width * 3, width, height, arr, n, ow, oh, icons.ctypes.data,
                                             rows.ctypes.data), "rmcv_icon_batch")
        return icons, rows, [Armour.from_c(arr[i]) for i in range(n)]

    def svm_predict(self, model: "SvmModel", rows: np.ndarray) -> np.ndarray:
        """labels[s] = int(svm.predict(rows[s])) (cv::ml::SVM, C_SVC + LINEAR)."""
        rows = np.ascontiguousarray(rows, np.float32).reshape(-1, model.sv.shape[1])
        out = np.zeros(len(rows), np.int32)
        if len(rows):
            self._check(self.lib.rmcv_svm_predict(self.h, C.byref(model.c), rows.ctypes.data, len(rows), out.ctypes.data), "rmcv_svm_predict")
        return out

    def identify_batch(self, d_bgr: int, width: int, height: int, armours: Sequence[Armour], model: "SvmModel", out_size=(20, 20),
                       pitch: Optional[int] = None) -> np.ndarray:
        """executable/main.cpp:178-181: icon crop + SVM identity of every armour of a device-resident frame."""
        n = len(armours)
        out = np.zeros(n, np.int32)
        if n:
            arr = (A.Armour * n)(*[a.to_c() for a in armours])
            self._check(self.lib.rmcv_identify_batch(self.h, d_bgr, pitch or width * 3, width, height, arr, n, int(out_size[0]),
                                                     int(out_size[1]), C.byref(model.c), out.ctypes.data), "rmcv_identify_batch")
        return out

    def solve_pnp(self, armours: Sequence[Armour], camera_matrix, dist_coeffs, exact_size=(27.0, 27.0), roi=(0.0, 0.0),
                  cam2world=None):
        """rm::solve_PnP for every armour -> list of (rvec[3], tvec[3], position[3], ok)."""
        n = len(armours)
        if n == 0:
            return []
        arr = (A.Armour * n)(*[a.to_c() for a in armours])
        K = np.ascontiguousarray(np.asarray(camera_matrix, np.float64).reshape(9))
        D = None if dist_coeffs is None else np.ascontiguousarray(np.asarray(dist_coeffs, np.float64).reshape(5))
        M = None if cam2world is None else np.ascontiguousarray(np.asarray(cam2world, np.float64).reshape(16))
        out = (A.Pose * n)()
        self._check(self.lib.rmcv_solve_pnp(self.h, arr, n, K.ctypes.data, None if D is None else D.ctypes.data,
                                            float(exact_size[0]), float(exact_size[1]), float(roi[0]), float(roi[1]),
                                            None if M is None else M.ctypes.data, out), "rmcv_solve_pnp")
        return [(np.array(o.rvec[:]), np.array(o.tvec[:]), np.array(o.position[:]), bool(o.ok)) for o in out]


class MultiContext:
    """rmcv_multi wrapper: ONE batch partitioned by frame across the GPUs of one box; one host thread and
    one rmcv_ctx per device inside the library, no collective, results concatenated in frame order."""

    def __init__(self, devices: Optional[Sequence[int]] = None, max_width=1280, max_height=1024, max_batch=64, chunk_frames=0,
                 max_runs_per_frame=0, max_blobs_per_frame=0, max_armours_per_frame=0):
        self.lib = load_library()
        cfg = A.Config()
        self.lib.rmcv_default_config(C.byref(cfg))
        cfg.max_width, cfg.max_height, cfg.max_batch = max_width, max_height, max_batch
        cfg.chunk_frames, cfg.max_runs_per_frame = chunk_frames, max_runs_per_frame
        cfg.max_blobs_per_frame, cfg.max_armours_per_frame = max_blobs_per_frame, max_armours_per_frame
        h = C.c_void_p()
        n = len(devices) if devices is not None else 0
        arr = (C.c_int * max(1, n))(*(devices or [0]))
        rc = self.lib.rmcv_multi_create(C.byref(cfg), arr if devices is not None else None, n, C.byref(h))
        if rc != A.RMCV_OK:
            raise RmcvError(rc, "rmcv_multi_create", "(is a CUDA device present? this library has no CPU path)")
        self.h = h
        self.n_devices = int(self.lib.rmcv_multi_device_count(self.h))

    def _check(self, rc: int, where: str):
        if rc != A.RMCV_OK:
            raise RmcvError(rc, where, self.lib.rmcv_multi_last_error(self.h).decode(errors="replace"))

    def close(self):
        if getattr(self, "h", None):
            self.lib.rmcv_multi_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def slice(self, batch: int, g: int) -> Tuple[int, int]:
        first, count = C.c_int(0), C.c_int(0)
        self.lib.rmcv_multi_slice(int(batch), self.n_devices, int(g), C.byref(first), C.byref(count))
        return first.value, count.value

    def detect_batch_host(self, frames: np.ndarray, params: A.Params, masks: Optional[np.ndarray] = None) -> A.Results:
        assert frames.dtype == np.uint8 and frames.ndim == 4 and frames.shape[3] == 3 and frames.flags.c_contiguous
        B, H, W, _ = frames.shape
        res = A.Results()
        mptr = masks.ctypes.data if masks is not None else None
        self._check(self.lib.rmcv_multi_detect_batch_host(self.h, frames.ctypes.data, W * 3, W * 3 * H, W, H, B, C.byref(params),
                                                          mptr, W, W * H, C.byref(res)), "rmcv_multi_detect_batch_host")
        return res

    def bayer_detect_batch_host(self, raw: np.ndarray, layout: int, params: A.Params, masks: Optional[np.ndarray] = None) -> A.Results:
        assert raw.dtype == np.uint8 and raw.ndim == 3 and raw.flags.c_contiguous
        B, H, W = raw.shape
        res = A.Results()
        mptr = masks.ctypes.data if masks is not None else None
        self._check(self.lib.rmcv_multi_bayer_detect_batch_host(self.h, raw.ctypes.data, W, W * H, W, H, B, int(layout), C.byref(params),
                                                                mptr, W, W * H, C.byref(res)), "rmcv_multi_bayer_detect_batch_host")
        return res

    frame_detections = staticmethod(Context.frame_detections)


class SvmModel:
    """A trained cv::ml::SVM (C_SVC, LINEAR kernel) as plain arrays: support_vectors (sv_total x var_count float32),
    class_labels (ascending), and per one-vs-one decision function rho and (alpha, support-vector index) lists.
    From cv2: SvmModel.from_cv2(svm, class_labels)."""

    def __init__(self, support_vectors, class_labels, rho, alphas, indices):
        self.sv = np.ascontiguousarray(support_vectors, np.float32)
        self.labels = np.ascontiguousarray(class_labels, np.int32)
        self.rho = np.ascontiguousarray(rho, np.float64)
        ofs = [0]
        for a in alphas:
            ofs.append(ofs[-1] + len(a))
        self.ofs = np.ascontiguousarray(ofs, np.int32)
        self.alpha = np.ascontiguousarray(np.concatenate([np.asarray(a, np.float64).reshape(-1) for a in alphas]), np.float64)
        self.index = np.ascontiguousarray(np.concatenate([np.asarray(i, np.int32).reshape(-1) for i in indices]), np.int32)
        c = A.SvmModel()
        c.var_count, c.class_count, c.sv_total = self.sv.shape[1], len(self.labels), self.sv.shape[0]
        c.support_vectors, c.class_labels, c.rho = self.sv.ctypes.data, self.labels.ctypes.data, self.rho.ctypes.data
        c.df_ofs, c.df_alpha, c.df_index = self.ofs.ctypes.data, self.alpha.ctypes.data, self.index.ctypes.data
        self.c = c

    @staticmethod
    def from_cv2(svm, class_labels) -> "SvmModel":
        k = len(class_labels)
        dfs = [svm.getDecisionFunction(i) for i in range(k * (k - 1) // 2)]
        return SvmModel(svm.getSupportVectors(), sorted(class_labels), [d[0] for d in dfs], [d[1] for d in dfs], [d[2] for d in dfs])


class Tracker:
    """f3: device-resident track list of one camera stream (executable/main.cpp:57-88, src/core.cpp:51-161)."""

    def __init__(self, ctx: "Context", capacity: int = 64):
        self.ctx = ctx
        h = C.c_void_p()
        ctx._check(ctx.lib.rmcv_tracker_create(ctx.h, capacity, C.byref(h)), "rmcv_tracker_create")
        self.h, self.capacity = h, capacity

    def close(self):
        if self.h:
            self.ctx.lib.rmcv_tracker_destroy(self.ctx.h, self.h)
            self.h = None

    def reset(self):
        self.ctx._check(self.ctx.lib.rmcv_tracker_reset(self.ctx.h, self.h), "rmcv_tracker_reset")

    def update(self, armours: Sequence[Armour], positions, identities, timestamp: int, tick_frequency: float,
               process_noise=5e-5, measurement_noise=0.5, error=0.05):
        """One iteration of the tracking loop for the armours of one frame (noise defaults: main.cpp:195)."""
        n = len(armours)
        arr = (A.Armour * max(n, 1))(*[a.to_c() for a in armours])
        pos = np.ascontiguousarray(np.asarray(positions, np.float64).reshape(n, 3)) if n else np.zeros((1, 3))
        ids = None if identities is None else np.ascontiguousarray(np.asarray(identities, np.int32).reshape(n))
        self.ctx._check(self.ctx.lib.rmcv_tracker_update(self.ctx.h, self.h, arr, pos.ctypes.data,
                                                         None if ids is None or n == 0 else ids.ctypes.data, n, int(timestamp),
                                                         float(tick_frequency), float(process_noise), float(measurement_noise),
                                                         float(error)), "rmcv_tracker_update")

    def read(self) -> list:
        out = (A.Track * self.capacity)()
        n = C.c_int()
        self.ctx._check(self.ctx.lib.rmcv_tracker_read(self.ctx.h, self.h, out, self.capacity, C.byref(n)), "rmcv_tracker_read")
        return [out[i] for i in range(n.value)]

    def identity_max(self, track) -> tuple:
        ident, prob = C.c_int32(), C.c_double()
        self.ctx._check(self.ctx.lib.rmcv_track_identity_max(C.byref(track), C.byref(ident), C.byref(prob)),
                        "rmcv_track_identity_max")
        return ident.value, prob.value


# --------------------------------------------------------------------------------------------- rm:: mirror
_default_ctx: Optional[Context] = None


def _ctx_for(width: int, height: int) -> Context:
    global _default_ctx
    c = _default_ctx
    if c is None or c.cfg.max_width < width or c.cfg.max_height < height:
        if c is not None:
            c.close()
        _default_ctx = c = Context(max_width=max(width, 1280), max_height=max(height, 1024), max_batch=1)
    return c


def extract_color(image: np.ndarray, target: int, lower_bound: int):
    """rm::extract_color (include/imgproc.h:29) -> (contours, binary).  contours: list of N×2 int32 arrays in
    cv::findContours order; binary: H×W uint8 {0,255}."""
    assert image.dtype == np.uint8 and image.ndim == 3 and image.shape[2] == 3
    image = np.ascontiguousarray(image)
    H, W, _ = image.shape
    ctx = _ctx_for(W, H)
    p = default_params(target=target, lower_bound=lower_bound, area_range=(0.0, 1e300))
    mask = np.empty((1, H, W), np.uint8)
    res = ctx.detect_batch_host(image[None], p, mask)
    return ctx.get_contours(0), mask[0]


def filter_lightblobs(contours, tilt_max, ratio_range, area_range, enemy):
    """rm::filter_lightblobs (include/objdetect.h:47-49) -> (positive, negative)."""
    if len(contours) == 0:
        return [], []
    ctx = _ctx_for(1, 1) if _default_ctx is None else _default_ctx
    p = default_params(target=enemy, tilt_max=tilt_max, ratio_range=ratio_range, area_range=area_range)
    infos, blobs = ctx.filter_lightblobs_raw(contours, p)
    negative = [c for c, i in zip(contours, infos) if i.status == A.CONTOUR_NEGATIVE]
    return blobs, negative


def filter_armours(lightblobs, angle_difference_max, shear_max, lenght_ratio_max, enemy):
    """rm::filter_armours (include/objdetect.h:70-71)."""
    ctx = _ctx_for(1, 1) if _default_ctx is None else _default_ctx
    p = default_params(target=enemy, angle_difference_max=angle_difference_max, shear_max=shear_max,
                       lenght_ratio_max=lenght_ratio_max)
    return ctx.filter_armours_raw(lightblobs, p)

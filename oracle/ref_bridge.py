"""TEST INFRASTRUCTURE ONLY — Python face of oracle/_ref/librmcv_ref.so.

`librmcv_ref.so` is the REFERENCE'S OWN C++ for the hot path (src/core.cpp, src/objdetect.cpp, src/imgproc.cpp,
src/mobility.cpp of /root/reference, compiled unmodified by oracle/Makefile against oracle/cvstub/opencv2).  Every rm::
statement in it is the reference's compiled code; every cv:: call it makes lands in `_cvcall` below and is served by the
real OpenCV of this image (cv2).  It exists to pin oracle/rm_oracle.py (tests/test_ref_pin.py), to generate
tests/golden/ (scripts/make_golden.py writes `"source": "_ref"`), and as the CPU arm of bench.py (`kind: "reference"`).

Never imported by the product path (rmcv_b200/).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Sequence, Tuple

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
LIB_PATH = os.path.join(REF_DIR, "librmcv_ref.so")
LIB_PATH_MATH_H = os.path.join(REF_DIR, "librmcv_ref_mathh.so")

#: seed installed before every fitEllipseDirect call, like oracle/rm_oracle.py does
FIT_RNG_SEED = 0

_DEPTH = {0: np.uint8, 1: np.int8, 2: np.uint16, 3: np.int16, 4: np.int32, 5: np.float32, 6: np.float64}
_DEPTH_OF = {np.dtype(v): k for k, v in _DEPTH.items()}


class _Arr(C.Structure):
    _fields_ = [("data", C.c_void_p), ("rows", C.c_int32), ("cols", C.c_int32), ("type", C.c_int32), ("step", C.c_int64),
                ("owner", C.c_int64)]


_REL = C.CFUNCTYPE(None, C.c_int64)
_CB = C.CFUNCTYPE(C.c_int, C.c_char_p, C.POINTER(_Arr), C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(_Arr), C.c_int)


class RefBlob(C.Structure):  # rm::lightblob public fields = rmcv_lightblob layout
    _fields_ = [("angle", C.c_float), ("target", C.c_int32), ("center", C.c_float * 2), ("vertices", (C.c_float * 2) * 4),
                ("size", C.c_float * 2)]


class RefArmour(C.Structure):
    _fields_ = [("icon", (C.c_float * 2) * 4), ("vertices", (C.c_float * 2) * 4), ("bounding_box", C.c_float * 4),
                ("i", C.c_int32), ("j", C.c_int32)]


def _to_np(a: _Arr) -> np.ndarray:
    depth, cn = a.type & 7, ((a.type >> 3) & 511) + 1
    dt = np.dtype(_DEPTH[depth])
    if a.rows <= 0 or a.cols <= 0 or not a.data:
        return np.zeros((0, 0) if cn == 1 else (0, 0, cn), dt)
    shape = (a.rows, a.cols) if cn == 1 else (a.rows, a.cols, cn)
    strides = (a.step, dt.itemsize * cn) if cn == 1 else (a.step, dt.itemsize * cn, dt.itemsize)
    buf = (C.c_ubyte * (a.step * (a.rows - 1) + a.cols * cn * dt.itemsize)).from_address(a.data)
    return np.ndarray(shape, dt, buf, 0, strides)


class RefLibrary:
    """One loaded variant of the reference build (default: <emmintrin.h> overload environment)."""

    def __init__(self, path: str = LIB_PATH):
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle` where /root/reference exists")
        self.lib = C.CDLL(path)
        self._owned = {}          # results the C++ side still references (released through _release)
        self._next_owner = 1
        self._cb = _CB(self._cvcall)
        self._rel = _REL(self._release)
        L = self.lib
        L.rmcv_ref_set_callback.argtypes = [_CB]
        L.rmcv_ref_set_callback(self._cb)
        L.rmcv_ref_set_release.argtypes = [_REL]
        L.rmcv_ref_set_release(self._rel)
        L.rmcv_ref_last_error.restype = C.c_char_p
        L.rmcv_ref_armour_new.restype = C.c_void_p
        L.rmcv_ref_armour_new.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_double), C.c_int, C.c_longlong]
        for name in ("rmcv_ref_armour_delete", "rmcv_ref_armour_reset", "rmcv_ref_armour_update_obs", "rmcv_ref_armour_update_time",
                     "rmcv_ref_armour_identity_max", "rmcv_ref_armour_max_iou", "rmcv_ref_armour_get", "rmcv_ref_armour_set_lost",
                     "rmcv_ref_armour_state"):
            getattr(L, name).argtypes = None
        self.with_math_h = bool(L.rmcv_ref_with_math_h())

    # ------------------------------------------------------------------ cv:: calls of the reference -> real OpenCV
    def _cvcall(self, op, ain, n_in, params, n_params, aout, n_out) -> int:
        try:
            ins = [_to_np(ain[i]) for i in range(n_in)]
            p = [params[i] for i in range(n_params)]
            outs = self._dispatch(op.decode(), ins, p)
            for i in range(n_out):
                o = np.ascontiguousarray(outs[i])
                if o.ndim == 1:
                    o = o.reshape(1, -1)
                if op in (b"kalmanPredict", b"kalmanCorrect"):   # attributes of a cv2.KalmanFilter that dies with this call
                    o = o.copy()
                cn = 1 if o.ndim == 2 else o.shape[2]
                aout[i].data = o.ctypes.data if o.size else None
                aout[i].rows, aout[i].cols = (o.shape[0], o.shape[1]) if o.size else (0, 0)
                aout[i].type = _DEPTH_OF[o.dtype] + ((cn - 1) << 3)
                aout[i].step = o.strides[0] if o.size else 0
                aout[i].owner = 0
                if o.size:          # the stub adopts the buffer (no copy) and releases it when its cv::Mat dies
                    owner = self._next_owner
                    self._next_owner += 1
                    self._owned[owner] = o
                    aout[i].owner = owner
            return 0
        except Exception as e:  # noqa: BLE001 — reported through the C side as a failed call
            self._cb_error = repr(e)
            return 1

    def _release(self, owner: int) -> None:
        self._owned.pop(owner, None)

    def _dispatch(self, op: str, a: List[np.ndarray], p: List[float]):
        if op == "boxPoints":
            b = a[0].ravel()
            return [cv2.boxPoints(((float(b[0]), float(b[1])), (float(b[2]), float(b[3])), float(b[4])))]
        if op == "boundingRect":
            return [np.array(cv2.boundingRect(np.ascontiguousarray(a[0]).reshape(-1, 1, 2)), np.int32)]
        if op == "contourArea":
            return [np.array([cv2.contourArea(np.ascontiguousarray(a[0]).reshape(-1, 1, 2), bool(p[0]))], np.float64)]
        if op in ("fitEllipseDirect", "fitEllipse", "minAreaRect"):
            pts = np.ascontiguousarray(a[0]).reshape(-1, 1, 2)
            if op == "fitEllipseDirect":
                cv2.setRNGSeed(FIT_RNG_SEED)
            (cx, cy), (w, h), ang = getattr(cv2, op)(pts)
            return [np.array([cx, cy, w, h, ang], np.float32)]
        if op == "split":
            return list(cv2.split(np.ascontiguousarray(a[0])))
        if op == "subtract":
            return [cv2.subtract(a[0], a[1])]
        if op == "matmul":
            return [a[0] @ a[1]]
        if op == "inRange":
            lo, hi = p[0:4], p[4:8]
            cn = 1 if a[0].ndim == 2 else a[0].shape[2]
            return [cv2.inRange(np.ascontiguousarray(a[0]), tuple(lo[:cn]) if cn > 1 else lo[0], tuple(hi[:cn]) if cn > 1 else hi[0])]
        if op == "getStructuringElement":
            return [cv2.getStructuringElement(int(p[0]), (int(p[1]), int(p[2])))]
        if op == "morphologyEx":
            return [cv2.morphologyEx(np.ascontiguousarray(a[0]), int(p[0]), np.ascontiguousarray(a[1]))]
        if op == "findContours":
            contours, _ = cv2.findContours(np.ascontiguousarray(a[0]), int(p[0]), int(p[1]))
            off = np.zeros(len(contours) + 1, np.int32)
            off[1:] = np.cumsum([len(c) for c in contours])
            pts = np.concatenate([c.reshape(-1, 2) for c in contours]).astype(np.int32) if contours else np.zeros((0, 2), np.int32)
            return [pts, off]
        if op == "mean":
            return [np.array(cv2.mean(np.ascontiguousarray(a[0])), np.float64)]
        if op == "getAffineTransform":
            return [cv2.getAffineTransform(np.ascontiguousarray(a[0]), np.ascontiguousarray(a[1]))]
        if op == "warpAffine":
            src = np.ascontiguousarray(a[0])
            dsize = (int(p[0]), int(p[1]))
            if dsize[0] <= 0 or dsize[1] <= 0:   # cv::warpAffine: an empty dsize means the source size
                dsize = (src.shape[1], src.shape[0])
            return [cv2.warpAffine(src, np.ascontiguousarray(a[1]), dsize, flags=int(p[2]), borderMode=int(p[3]))]
        if op == "resize":
            return [cv2.resize(np.ascontiguousarray(a[0]), (int(p[0]), int(p[1])), fx=p[2], fy=p[3], interpolation=int(p[4]))]
        if op == "LUT":
            return [cv2.LUT(np.ascontiguousarray(a[0]), np.ascontiguousarray(a[1]))]
        if op == "cvtColor":
            return [cv2.cvtColor(np.ascontiguousarray(a[0]), int(p[0]))]
        if op == "convertTo":
            depth = int(p[0]) & 7
            if depth in (5, 6):   # to float: exact for the 8-bit sources the reference converts (src/core.cpp:212)
                return [(np.ascontiguousarray(a[0]).astype(np.float64) * p[1] + p[2]).astype(_DEPTH[depth])]
            raise NotImplementedError("convertTo to an integer depth")
        if op == "Rodrigues":
            return [cv2.Rodrigues(np.ascontiguousarray(a[0]))[0]]
        if op == "solvePnP":
            ok, rvec, tvec = cv2.solvePnP(np.ascontiguousarray(a[0]).reshape(-1, 3), np.ascontiguousarray(a[1]).reshape(-1, 2),
                                          np.ascontiguousarray(a[2]), np.ascontiguousarray(a[3]),
                                          useExtrinsicGuess=bool(p[0]), flags=int(p[1]))
            if not ok:
                raise RuntimeError("solvePnP failed")
            return [rvec.reshape(3, 1), tvec.reshape(3, 1)]
        if op in ("kalmanPredict", "kalmanCorrect"):
            kf = cv2.KalmanFilter(6, 6, 0, cv2.CV_64F)
            names = ["statePre", "statePost", "transitionMatrix", "measurementMatrix", "processNoiseCov", "measurementNoiseCov",
                     "errorCovPre", "errorCovPost", "gain"]
            for name, m in zip(names, a[:9]):
                setattr(kf, name, np.ascontiguousarray(m, np.float64))
            if op == "kalmanPredict":
                kf.predict()
                return [kf.statePre, kf.statePost, kf.errorCovPre, kf.errorCovPost]
            with np.errstate(all="ignore"):
                kf.correct(np.ascontiguousarray(a[9], np.float64))
            return [kf.statePost, kf.errorCovPost, kf.gain]
        raise NotImplementedError(op)

    def _check(self, rc: int):
        if rc != 0:
            raise RuntimeError(f"reference call failed: {self.lib.rmcv_ref_last_error().decode()} / {getattr(self, '_cb_error', '')}")

    # ------------------------------------------------------------------ rm:: entry points
    @staticmethod
    def blob_pod(angle, target, center, vertices, size) -> RefBlob:
        b = RefBlob()
        b.angle, b.target = float(angle), int(target)
        b.center[0], b.center[1] = float(center[0]), float(center[1])
        for k in range(4):
            b.vertices[k][0], b.vertices[k][1] = float(vertices[k][0]), float(vertices[k][1])
        b.size[0], b.size[1] = float(size[0]), float(size[1])
        return b

    def make_lightblob(self, box: Sequence[float], target: int) -> RefBlob:
        """rm::lightblob::lightblob(cv::RotatedRect(cx, cy, w, h, angle), target) — src/core.cpp:9-19."""
        out = RefBlob()
        self._check(self.lib.rmcv_ref_make_lightblob((C.c_float * 5)(*[float(v) for v in box]), int(target), C.byref(out)))
        return out

    def make_armour(self, b0: RefBlob, b1: RefBlob) -> RefArmour:
        """rm::armour::armour({b0, b1}) — src/core.cpp:21-49."""
        out = RefArmour()
        self._check(self.lib.rmcv_ref_make_armour(C.byref(b0), C.byref(b1), C.byref(out)))
        return out

    def pair_passes(self, bi: RefBlob, bj: RefBlob, angle_difference_max, shear_max, lenght_ratio_max, enemy) -> bool:
        """rm::filter_armours on the two-blob list {bi, bj} — src/objdetect.cpp:114-166."""
        n = C.c_int(0)
        self._check(self.lib.rmcv_ref_pair_passes(C.byref(bi), C.byref(bj), C.c_float(angle_difference_max), C.c_float(shear_max),
                                                  C.c_float(lenght_ratio_max), int(enemy), C.byref(n)))
        return n.value == 1

    def filter_armours(self, blobs: Sequence[RefBlob], angle_difference_max, shear_max, lenght_ratio_max, enemy) -> List[RefArmour]:
        n = len(blobs)
        arr = (RefBlob * max(1, n))(*blobs)
        cap = max(1, n * (n - 1) // 2)
        out = (RefArmour * cap)()
        cnt = C.c_int(0)
        self._check(self.lib.rmcv_ref_filter_armours(arr, n, C.c_float(angle_difference_max), C.c_float(shear_max),
                                                     C.c_float(lenght_ratio_max), int(enemy), out, cap, C.byref(cnt)))
        return [out[k] for k in range(cnt.value)]

    @staticmethod
    def _flatten(contours: Sequence[np.ndarray]):
        off = np.zeros(len(contours) + 1, np.int32)
        off[1:] = np.cumsum([len(c) for c in contours])
        xy = (np.concatenate([np.asarray(c, np.int32).reshape(-1, 2) for c in contours]) if len(contours) else np.zeros((0, 2), np.int32))
        return np.ascontiguousarray(xy, np.int32), off

    def filter_lightblobs(self, contours, tilt_max, ratio_range, area_range, enemy) -> Tuple[List[RefBlob], List[int]]:
        """rm::filter_lightblobs — src/objdetect.cpp:55-87.  Returns (positive blobs, indices of the negative contours)."""
        xy, off = self._flatten(contours)
        n = len(contours)
        pos = (RefBlob * max(1, n))()
        neg = (C.c_int32 * max(1, n))()
        npos, nneg = C.c_int(0), C.c_int(0)
        self._check(self.lib.rmcv_ref_filter_lightblobs(
            xy.ctypes.data_as(C.POINTER(C.c_int32)), off.ctypes.data_as(C.POINTER(C.c_int32)), n, C.c_float(tilt_max),
            C.c_float(ratio_range[0]), C.c_float(ratio_range[1]), C.c_double(area_range[0]), C.c_double(area_range[1]), int(enemy),
            pos, C.byref(npos), neg, C.byref(nneg)))
        return [pos[k] for k in range(npos.value)], [int(neg[k]) for k in range(nneg.value)]

    def match_lightblob(self, contour, min_ratio, max_ratio, tilt_angle, min_area, max_area, fit_ellipse=True):
        """rm::MatchLightBlob — src/objdetect.cpp:9-28 -> (matched, box[5])."""
        xy = np.ascontiguousarray(np.asarray(contour, np.int32).reshape(-1, 2))
        m = C.c_int(0)
        box = (C.c_float * 5)()
        self._check(self.lib.rmcv_ref_match_lightblob(xy.ctypes.data_as(C.POINTER(C.c_int32)), len(xy), C.c_float(min_ratio),
                                                      C.c_float(max_ratio), C.c_float(tilt_angle), C.c_float(min_area),
                                                      C.c_float(max_area), int(bool(fit_ellipse)), C.byref(m), box))
        return bool(m.value), [float(v) for v in box]

    def find_lightblobs(self, contours, min_ratio, max_ratio, tilt_angle, min_area, max_area, source, fit_ellipse=True) -> List[RefBlob]:
        """rm::FindLightBlobs — src/objdetect.cpp:30-53."""
        xy, off = self._flatten(contours)
        src = np.ascontiguousarray(source)
        ch = 1 if src.ndim == 2 else src.shape[2]
        out = (RefBlob * max(1, len(contours)))()
        n = C.c_int(0)
        self._check(self.lib.rmcv_ref_find_lightblobs(
            xy.ctypes.data_as(C.POINTER(C.c_int32)), off.ctypes.data_as(C.POINTER(C.c_int32)), len(contours), C.c_float(min_ratio),
            C.c_float(max_ratio), C.c_float(tilt_angle), C.c_float(min_area), C.c_float(max_area),
            src.ctypes.data_as(C.POINTER(C.c_uint8)), src.shape[0], src.shape[1], ch, int(bool(fit_ellipse)), out, C.byref(n)))
        return [out[k] for k in range(n.value)]

    def lightblob_overlap(self, blobs: Sequence[RefBlob], left: int, right: int) -> bool:
        arr = (RefBlob * max(1, len(blobs)))(*blobs)
        r = C.c_int(0)
        self._check(self.lib.rmcv_ref_lightblob_overlap(arr, len(blobs), int(left), int(right), C.byref(r)))
        return bool(r.value)

    def point_distance(self, p1, p2) -> np.float32:
        o = C.c_float(0)
        self._check(self.lib.rmcv_ref_point_distance((C.c_float * 2)(*map(float, p1)), (C.c_float * 2)(*map(float, p2)), C.byref(o)))
        return np.float32(o.value)

    def extend_cord(self, p1, p2, delta):
        d1, d2 = (C.c_float * 2)(), (C.c_float * 2)()
        self._check(self.lib.rmcv_ref_extend_cord((C.c_float * 2)(*map(float, p1)), (C.c_float * 2)(*map(float, p2)), C.c_float(delta), d1, d2))
        return np.array(d1[:], np.float32), np.array(d2[:], np.float32)

    def calc_perspective(self, inp, out_ratio=1.0) -> np.ndarray:
        i = (C.c_float * 8)(*[float(v) for v in np.asarray(inp, np.float32).ravel()])
        o = (C.c_float * 8)()
        self._check(self.lib.rmcv_ref_calc_perspective(i, C.c_float(out_ratio), o))
        return np.array(o[:], np.float32).reshape(4, 2)

    def line_center(self, p1, p2) -> np.ndarray:
        o = (C.c_float * 2)()
        self._check(self.lib.rmcv_ref_line_center((C.c_float * 2)(*map(float, p1)), (C.c_float * 2)(*map(float, p2)), o))
        return np.array(o[:], np.float32)

    def extract_color(self, image: np.ndarray, target: int, lower_bound: int):
        """rm::extract_color — src/imgproc.cpp:50-75 -> (contours, binary)."""
        img = np.ascontiguousarray(image, np.uint8)
        rows, cols = img.shape[:2]
        ch = 1 if img.ndim == 2 else img.shape[2]
        binary = np.empty((rows, cols), np.uint8)
        cap_pts, cap_off = rows * cols + 16, rows * cols // 2 + 16
        xy = np.empty((cap_pts, 2), np.int32)
        off = np.zeros(cap_off, np.int32)
        n = C.c_int(0)
        self._check(self.lib.rmcv_ref_extract_color(img.ctypes.data_as(C.POINTER(C.c_uint8)), rows, cols, ch, int(target), int(lower_bound),
                                                    binary.ctypes.data_as(C.POINTER(C.c_uint8)), xy.ctypes.data_as(C.POINTER(C.c_int32)),
                                                    cap_pts, off.ctypes.data_as(C.POINTER(C.c_int32)), cap_off, C.byref(n)))
        return [xy[off[k]:off[k + 1]].copy() for k in range(n.value)], binary

    def affine_correction(self, source: np.ndarray, vertices, out_size=(20, 20)):
        """rm::affine_correction — src/imgproc.cpp:9-35 -> (calibration, vertices clamped in place)."""
        img = np.ascontiguousarray(source, np.uint8)
        ch = 1 if img.ndim == 2 else img.shape[2]
        v = (C.c_float * 8)(*[float(x) for x in np.asarray(vertices, np.float32).ravel()])
        out = np.empty((out_size[1], out_size[0], ch) if ch > 1 else (out_size[1], out_size[0]), np.uint8)
        self._check(self.lib.rmcv_ref_affine_correction(img.ctypes.data_as(C.POINTER(C.c_uint8)), img.shape[0], img.shape[1], ch, v,
                                                        int(out_size[0]), int(out_size[1]), out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out, np.array(v[:], np.float32).reshape(4, 2)

    def solve_pnp(self, points_image, camera_matrix, dist_coeffs, exact_size=(27.0, 27.0), roi=(0, 0)):
        """rm::solve_PnP — src/mobility.cpp:166-190 -> (rvec[3], tvec[3])."""
        p = (C.c_float * 8)(*[float(x) for x in np.asarray(points_image, np.float32).ravel()])
        K = (C.c_double * 9)(*np.asarray(camera_matrix, np.float64).ravel())
        D = (C.c_double * 5)(*np.asarray(dist_coeffs, np.float64).ravel())
        r, t = (C.c_double * 3)(), (C.c_double * 3)()
        self._check(self.lib.rmcv_ref_solve_pnp(p, K, D, C.c_float(exact_size[0]), C.c_float(exact_size[1]), int(roi[0]), int(roi[1]), r, t))
        return np.array(r[:]), np.array(t[:])


class RefTrackedArmour:
    """The tracking side of one rm::armour object (src/core.cpp:51-162), held by the reference build."""

    def __init__(self, ref: RefLibrary, bounding_box, position, identity, timestamp):
        self.ref = ref
        self.h = ref.lib.rmcv_ref_armour_new((C.c_float * 4)(*[float(v) for v in bounding_box]),
                                             (C.c_double * 3)(*[float(v) for v in position]), int(identity), C.c_longlong(int(timestamp)))
        if not self.h:
            raise RuntimeError("rmcv_ref_armour_new failed")
        self.h = C.c_void_p(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.ref.lib.rmcv_ref_armour_delete(self.h)
            self.h = None

    def reset(self, q, r, err):
        self.ref._check(self.ref.lib.rmcv_ref_armour_reset(self.h, C.c_double(q), C.c_double(r), C.c_double(err)))

    def update_observation(self, obs: "RefTrackedArmour"):
        self.ref._check(self.ref.lib.rmcv_ref_armour_update_obs(self.h, obs.h))

    def update_time(self, ts: int):
        self.ref._check(self.ref.lib.rmcv_ref_armour_update_time(self.h, C.c_longlong(int(ts))))

    @property
    def timestamp(self) -> int:
        ts, lost = C.c_longlong(0), C.c_int(0)
        self.ref._check(self.ref.lib.rmcv_ref_armour_get(self.h, C.byref(ts), C.byref(lost)))
        return ts.value

    @property
    def lost_count(self) -> int:
        ts, lost = C.c_longlong(0), C.c_int(0)
        self.ref._check(self.ref.lib.rmcv_ref_armour_get(self.h, C.byref(ts), C.byref(lost)))
        return lost.value

    @lost_count.setter
    def lost_count(self, v: int):
        self.ref.lib.rmcv_ref_armour_set_lost(self.h, int(v))

    def state(self):
        """(statePost[6], errorCovPost[6][6], initialized) of the private cv::KalmanFilter observer."""
        s, c, ini = (C.c_double * 6)(), (C.c_double * 36)(), C.c_int(0)
        self.ref._check(self.ref.lib.rmcv_ref_armour_state(self.h, s, c, C.byref(ini)))
        return np.array(s[:]), np.array(c[:]).reshape(6, 6), bool(ini.value)

    def identity_max(self):
        i, p = C.c_int(0), C.c_double(0)
        self.ref._check(self.ref.lib.rmcv_ref_armour_identity_max(self.h, C.byref(i), C.byref(p)))
        return i.value, p.value

    def max_iou(self, others: Sequence["RefTrackedArmour"]):
        arr = (C.c_void_p * max(1, len(others)))(*[o.h for o in others])
        i, v = C.c_int(0), C.c_float(0)
        self.ref._check(self.ref.lib.rmcv_ref_armour_max_iou(self.h, arr, len(others), C.byref(i), C.byref(v)))
        return i.value, np.float32(v.value)


_default: RefLibrary | None = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def get(path: str | None = None) -> RefLibrary:
    global _default
    if path is not None:
        return RefLibrary(path)
    if _default is None:
        _default = RefLibrary(LIB_PATH)
    return _default


# ---------------------------------------------------------------------------------------- the whole path through the reference
class RefFrameResult:
    """executable/main.cpp:172-176 run through the reference build: rm::extract_color -> rm::filter_lightblobs ->
    rm::filter_armours.  `status[k]` per contour: 0 skipped, 1 positive, 2 negative (recovered from the order-preserving
    outputs); `pairs[k]` = (i, j) of armour k, recovered by walking rm::filter_armours over every two-blob list."""

    def __init__(self, binary, contours, status, positive, negative_index, armours, pairs):
        self.binary, self.contours, self.status = binary, contours, status
        self.positive, self.negative_index, self.armours, self.pairs = positive, negative_index, armours, pairs


def blob_arrays(b: RefBlob):
    """(angle, target, center[2], vertices[4][2], size[2]) of a RefBlob as float32 numpy values."""
    return (np.float32(b.angle), int(b.target), np.array(b.center[:], np.float32),
            np.array([[b.vertices[k][0], b.vertices[k][1]] for k in range(4)], np.float32), np.array(b.size[:], np.float32))


def armour_arrays(a: RefArmour):
    return (np.array([[a.icon[k][0], a.icon[k][1]] for k in range(4)], np.float32),
            np.array([[a.vertices[k][0], a.vertices[k][1]] for k in range(4)], np.float32),
            tuple(float(v) for v in a.bounding_box))


def detect_frame(image: np.ndarray, target=1, lower_bound=80, tilt_max=70.0, ratio_range=(1.5, 80.0), area_range=(10.0, 99999.0),
                 angle_difference_max=12.0, shear_max=22.0, lenght_ratio_max=0.4, ref: RefLibrary | None = None,
                 with_pairs: bool = True) -> RefFrameResult:
    ref = ref or get()
    contours, binary = ref.extract_color(image, target, lower_bound)
    positive, neg_idx = ref.filter_lightblobs(contours, tilt_max, ratio_range, area_range, target)
    # a contour is positive iff it produced a blob; positives keep contour order, so walk both lists
    status = [0] * len(contours)
    for k in neg_idx:
        status[k] = 2
    if positive:
        # the remaining candidates are (size >= 6, area in range, not negative): exactly len(positive) of them
        cand = [k for k, c in enumerate(contours) if status[k] == 0 and len(c) >= 6 and
                area_range[0] <= cv2.contourArea(c.reshape(-1, 1, 2)) <= area_range[1]]
        assert len(cand) == len(positive), (len(cand), len(positive))
        for k in cand:
            status[k] = 1
    armours = ref.filter_armours(positive, angle_difference_max, shear_max, lenght_ratio_max, target)
    pairs = []
    if with_pairs:
        n = len(positive)
        for i in range(n - 1):
            for j in range(i + 1, n):
                if ref.pair_passes(positive[i], positive[j], angle_difference_max, shear_max, lenght_ratio_max, target):
                    pairs.append((i, j))
        assert len(pairs) == len(armours)
    return RefFrameResult(binary, contours, status, positive, neg_idx, armours, pairs)


def tracking_step(tracking: list, armours: list, noise=(5e-5, 0.5, 0.05)) -> list:
    """One iteration of the tracking thread's loop (executable/main.cpp:60-85 — a lambda inside main(), so it cannot be
    compiled from the reference; restated like oracle/rm_oracle.py::tracking_step) over RefTrackedArmour objects: every
    rm::armour method it calls is the reference's own."""
    armours = list(armours)
    if not armours:
        return tracking
    for a in armours:
        a.reset(*noise)
    if not tracking:
        return armours
    i = 0
    while i < len(tracking):
        index, iou = tracking[i].max_iou(armours)
        if iou > 0.5:
            tracking[i].update_observation(armours[index])
            del armours[index]
        else:
            lost = tracking[i].lost_count
            tracking[i].lost_count = lost + 1
            if lost > 25:
                del tracking[i]
            else:
                tracking[i].update_time(tracking[i].timestamp)
        i += 1
    tracking.extend(armours)
    return tracking

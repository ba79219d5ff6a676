"""ctypes mirror of include/rmcv_b200.h (struct layouts and prototypes)."""
from __future__ import annotations

import ctypes as C

ABI_VERSION = 3

RMCV_OK = 0
RMCV_ERR_INVALID_ARG, RMCV_ERR_CUDA, RMCV_ERR_CAPACITY, RMCV_ERR_NO_DEVICE, RMCV_ERR_STATE = -1, -2, -3, -4, -5
CAMP_RED, CAMP_BLUE, CAMP_GUIDELIGHT, CAMP_NEUTRAL = 0, 1, 2, -1
BAYER_RG, BAYER_GB, BAYER_GR, BAYER_BG = 1, 2, 3, 4
CONTOUR_SKIPPED, CONTOUR_POSITIVE, CONTOUR_NEGATIVE = 0, 1, 2
FIT_NONE, FIT_DIRECT, FIT_FALLBACK, FIT_FALLBACK_LONG = 0, 1, 2, 3
FRAME_OVERFLOW_MOMENTS = 16
STAGE_NAMES = ("pixel", "emit", "label", "contour", "fit", "order")


class RotatedRect(C.Structure):
    _fields_ = [("cx", C.c_float), ("cy", C.c_float), ("w", C.c_float), ("h", C.c_float), ("angle", C.c_float)]


class LightBlob(C.Structure):
    _fields_ = [("angle", C.c_float), ("target", C.c_int32), ("center", C.c_float * 2),
                ("vertices", (C.c_float * 2) * 4), ("size", C.c_float * 2)]


class Armour(C.Structure):
    _fields_ = [("icon", (C.c_float * 2) * 4), ("vertices", (C.c_float * 2) * 4), ("bounding_box", C.c_float * 4),
                ("i", C.c_int32), ("j", C.c_int32), ("gates", C.c_float * 6)]


class ContourInfo(C.Structure):
    _fields_ = [("first_x", C.c_int32), ("first_y", C.c_int32), ("n_points", C.c_int32), ("status", C.c_int32),
                ("area2", C.c_int64), ("bbox", C.c_int32 * 4), ("ellipse", RotatedRect), ("fit_branch", C.c_int32),
                ("det0", C.c_float), ("blob_index", C.c_int32)]


class FrameInfo(C.Structure):
    _fields_ = [("n_contours", C.c_int32), ("n_positive", C.c_int32), ("n_negative", C.c_int32),
                ("n_armours", C.c_int32), ("contour_offset", C.c_int32), ("blob_offset", C.c_int32),
                ("armour_offset", C.c_int32), ("flags", C.c_int32)]


class Params(C.Structure):
    _fields_ = [("target", C.c_int32), ("lower_bound", C.c_int32), ("tilt_max", C.c_float),
                ("ratio_min", C.c_float), ("ratio_max", C.c_float), ("area_min", C.c_double),
                ("area_max", C.c_double), ("angle_difference_max", C.c_float), ("shear_max", C.c_float),
                ("lenght_ratio_max", C.c_float)]


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("max_width", C.c_int32), ("max_height", C.c_int32),
                ("max_batch", C.c_int32), ("chunk_frames", C.c_int32), ("max_runs_per_frame", C.c_int32),
                ("max_blobs_per_frame", C.c_int32), ("max_armours_per_frame", C.c_int32), ("flags", C.c_int32),
                ("stream", C.c_void_p)]


class Pose(C.Structure):
    _fields_ = [("rvec", C.c_double * 3), ("tvec", C.c_double * 3), ("position", C.c_double * 3), ("reproj_err", C.c_double),
                ("ok", C.c_int32), ("pad", C.c_int32)]


class Results(C.Structure):
    _fields_ = [("batch", C.c_int32), ("total_contours", C.c_int32), ("total_blobs", C.c_int32),
                ("total_armours", C.c_int32), ("frames", C.POINTER(FrameInfo)), ("contours", C.POINTER(ContourInfo)),
                ("blobs", C.POINTER(LightBlob)), ("armours", C.POINTER(Armour)), ("poses", C.POINTER(Pose))]


class SvmModel(C.Structure):
    _fields_ = [("var_count", C.c_int32), ("class_count", C.c_int32), ("sv_total", C.c_int32), ("support_vectors", C.c_void_p),
                ("class_labels", C.c_void_p), ("rho", C.c_void_p), ("df_ofs", C.c_void_p), ("df_alpha", C.c_void_p),
                ("df_index", C.c_void_p)]


TRACK_HIST = 8


class Track(C.Structure):
    _fields_ = [("bbox", C.c_float * 4), ("position", C.c_double * 3), ("timestamp", C.c_int64), ("lost_count", C.c_int32),
                ("identity", C.c_int32), ("initialized", C.c_int32), ("n_hist", C.c_int32),
                ("hist_id", C.c_int32 * TRACK_HIST), ("hist_count", C.c_int32 * TRACK_HIST),
                ("state_pre", C.c_double * 6), ("state_post", C.c_double * 6), ("cov_pre", C.c_double * 36),
                ("cov_post", C.c_double * 36), ("meas", C.c_double * 6), ("q", C.c_double), ("r", C.c_double)]


assert C.sizeof(Pose) == 88
assert C.sizeof(LightBlob) == 56 and C.sizeof(Armour) == 112 and C.sizeof(ContourInfo) == 72 and C.sizeof(FrameInfo) == 32

_vp, _sz, _i, _u8p = C.c_void_p, C.c_size_t, C.c_int, C.c_void_p

#: name -> (restype, argtypes); every symbol include/rmcv_b200.h declares
PROTOTYPES = {
    "rmcv_abi_version": (C.c_int, []),
    "rmcv_status_string": (C.c_char_p, [_i]),
    "rmcv_default_params": (None, [C.POINTER(Params)]),
    "rmcv_default_config": (None, [C.POINTER(Config)]),
    "rmcv_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "rmcv_ctx_create": (C.c_int, [C.POINTER(Config), C.POINTER(_vp)]),
    "rmcv_ctx_destroy": (C.c_int, [_vp]),
    "rmcv_last_error": (C.c_char_p, [_vp]),
    "rmcv_chunk_frames": (C.c_int, [_vp]),
    "rmcv_device_alloc": (C.c_int, [_vp, _sz, C.POINTER(_vp)]),
    "rmcv_device_free": (C.c_int, [_vp, _vp]),
    "rmcv_host_alloc": (C.c_int, [_vp, _sz, C.POINTER(_vp)]),
    "rmcv_host_free": (C.c_int, [_vp, _vp]),
    "rmcv_memcpy_h2d": (C.c_int, [_vp, _vp, _vp, _sz]),
    "rmcv_memcpy_d2h": (C.c_int, [_vp, _vp, _vp, _sz]),
    "rmcv_memset_d": (C.c_int, [_vp, _vp, _i, _sz]),
    "rmcv_sync": (C.c_int, [_vp]),
    "rmcv_stream": (_vp, [_vp]),
    "rmcv_extract_color_batch": (C.c_int, [_vp, _u8p, _sz, _sz, _i, _i, _i, _i, _i, _u8p, _sz, _sz]),
    "rmcv_bayer_extract_color_batch": (C.c_int, [_vp, _u8p, _sz, _sz, _i, _i, _i, _i, _i, _i, _u8p, _sz, _sz]),
    "rmcv_detect_batch": (C.c_int, [_vp, _u8p, _sz, _sz, _i, _i, _i, C.POINTER(Params), _u8p, _sz, _sz]),
    "rmcv_bayer_detect_batch": (C.c_int, [_vp, _u8p, _sz, _sz, _i, _i, _i, _i, C.POINTER(Params), _u8p, _sz, _sz]),
    "rmcv_detect_batch_host": (C.c_int, [_vp, _u8p, _sz, _sz, _i, _i, _i, C.POINTER(Params), _u8p, _sz, _sz,
                                         C.POINTER(Results)]),
    "rmcv_bayer_detect_batch_host": (C.c_int, [_vp, _u8p, _sz, _sz, _i, _i, _i, _i, C.POINTER(Params), _u8p, _sz, _sz,
                                               C.POINTER(Results)]),
    "rmcv_multi_create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_int), _i, C.POINTER(_vp)]),
    "rmcv_multi_destroy": (C.c_int, [_vp]),
    "rmcv_multi_device_count": (C.c_int, [_vp]),
    "rmcv_multi_last_error": (C.c_char_p, [_vp]),
    "rmcv_multi_slice": (None, [_i, _i, _i, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "rmcv_multi_detect_batch_host": (C.c_int, [_vp, _u8p, _sz, _sz, _i, _i, _i, C.POINTER(Params), _u8p, _sz, _sz,
                                               C.POINTER(Results)]),
    "rmcv_multi_bayer_detect_batch_host": (C.c_int, [_vp, _u8p, _sz, _sz, _i, _i, _i, _i, C.POINTER(Params), _u8p, _sz, _sz,
                                                     C.POINTER(Results)]),
    "rmcv_fetch_results": (C.c_int, [_vp, C.POINTER(Results)]),
    "rmcv_get_contour": (C.c_int, [_vp, _i, _i, _vp, _i, C.POINTER(C.c_int)]),
    "rmcv_get_contours": (C.c_int, [_vp, _i, _vp, _i, _vp, _i, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "rmcv_get_label_map": (C.c_int, [_vp, _i, _vp, _sz]),
    "rmcv_get_bitmask": (C.c_int, [_vp, _i, _vp, _i]),
    "rmcv_filter_lightblobs": (C.c_int, [_vp, _vp, _vp, _i, C.POINTER(Params), _vp, _vp, _i, C.POINTER(C.c_int)]),
    "rmcv_filter_armours": (C.c_int, [_vp, _vp, _i, C.POINTER(Params), _vp, _i, C.POINTER(C.c_int)]),
    "rmcv_make_lightblobs": (C.c_int, [_vp, _vp, _i, _i, _vp]),
    "rmcv_match_lightblobs": (C.c_int, [_vp, _vp, _vp, _i, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _i, _vp, _vp]),
    "rmcv_find_lightblobs_legacy": (C.c_int, [_vp, _vp, _vp, _i, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                              _u8p, _sz, _i, _i, _i, _vp, _i, C.POINTER(C.c_int)]),
    "rmcv_min_area_rects": (C.c_int, [_vp, _vp, _vp, _i, _vp]),
    "rmcv_lightblob_overlap": (C.c_int, [_vp, _vp, _i, _i, _i, C.POINTER(C.c_int)]),
    "rmcv_raw_frontend_batch": (C.c_int, [_vp, _vp, _sz, _sz, _i, _i, _i, _i, _i, _i, _u8p, _sz, _sz]),
    "rmcv_frontend_layout": (C.c_int, [_i, _i, _i, _i, _i]),
    "rmcv_set_camera": (C.c_int, [_vp, _vp, _vp, C.c_float, C.c_float, _vp]),
    "rmcv_clear_camera": (C.c_int, [_vp]),
    "rmcv_solve_pnp": (C.c_int, [_vp, _vp, _i, _vp, _vp, C.c_float, C.c_float, C.c_float, C.c_float, _vp, _vp]),
    "rmcv_icon_batch": (C.c_int, [_vp, _u8p, _sz, _i, _i, _vp, _i, _i, _i, _vp, _vp]),
    "rmcv_svm_predict": (C.c_int, [_vp, C.POINTER(SvmModel), _vp, _i, _vp]),
    "rmcv_identify_batch": (C.c_int, [_vp, _u8p, _sz, _i, _i, _vp, _i, _i, _i, C.POINTER(SvmModel), _vp]),
    "rmcv_tracker_create": (C.c_int, [_vp, _i, C.POINTER(_vp)]),
    "rmcv_tracker_destroy": (C.c_int, [_vp, _vp]),
    "rmcv_tracker_reset": (C.c_int, [_vp, _vp]),
    "rmcv_tracker_update": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double]),
    "rmcv_tracker_read": (C.c_int, [_vp, _vp, _vp, _i, C.POINTER(C.c_int)]),
    "rmcv_track_identity_max": (C.c_int, [C.POINTER(Track), C.POINTER(C.c_int32), C.POINTER(C.c_double)]),
    "rmcv_profile_enable": (C.c_int, [_vp, _i]),
    "rmcv_profile_read": (C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_int64), _i]),
    "rmcv_timer_start": (C.c_int, [_vp]),
    "rmcv_timer_stop": (C.c_int, [_vp, C.POINTER(C.c_double)]),
    "rmcv_kernel_launches": (C.c_int64, [_vp]),
}

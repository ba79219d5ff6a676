"""rmcv_b200 — B200-native implementation of rmcv's per-frame detection hot path
(rm::extract_color -> rm::filter_lightblobs -> rm::filter_armours; reference executable/main.cpp:172-176).

The product is the CUDA library `librmcv_b200.so` behind the C ABI in include/rmcv_b200.h; this package is the
Python host mirror used by the tests and the benchmark.  Nothing here computes on the CPU.
"""
from . import _abi as abi
from .api import (Armour, Context, MultiContext, Tracker, SvmModel, ContourInfo, FrameDetections, LightBlob, RmcvError, default_params, extract_color,
                  filter_armours, filter_lightblobs, lib_path, load_library)
from ._abi import (BAYER_BG, BAYER_GB, BAYER_GR, BAYER_RG, CAMP_BLUE, CAMP_GUIDELIGHT, CAMP_NEUTRAL, CAMP_RED,
                   CONTOUR_NEGATIVE, CONTOUR_POSITIVE, CONTOUR_SKIPPED, FIT_DIRECT, FIT_FALLBACK, FIT_NONE)

__all__ = [n for n in dir() if not n.startswith("_")]

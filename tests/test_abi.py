"""CPU: the C-ABI library loads, exports every symbol include/rmcv_b200.h declares, the ctypes mirror agrees with the
header, and - without a GPU - the product path fails loudly instead of falling back to anything."""
import ctypes as C
import os
import re
import subprocess

import pytest

import rmcv_b200 as rb
from rmcv_b200 import _abi as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rmcv_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rmcv_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = rb.load_library()
    names = declared_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/rmcv_b200.h but not exported"
        assert n in A.PROTOTYPES, f"{n} missing from the ctypes mirror"
    assert lib.rmcv_abi_version() == A.ABI_VERSION


def test_header_compiles_as_c_and_struct_sizes_match():
    src = r'''
#include <stdio.h>
#include "rmcv_b200.h"
int main(void){ printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(rmcv_rotated_rect), sizeof(rmcv_lightblob), sizeof(rmcv_armour),
  sizeof(rmcv_contour_info), sizeof(rmcv_frame_info), sizeof(rmcv_params), sizeof(rmcv_config), sizeof(rmcv_results)); return 0; }
'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")], check=True)
        out = subprocess.run([os.path.join(d, "t")], check=True, capture_output=True, text=True).stdout.split()
    sizes = [int(x) for x in out]
    assert sizes == [C.sizeof(A.RotatedRect), C.sizeof(A.LightBlob), C.sizeof(A.Armour), C.sizeof(A.ContourInfo), C.sizeof(A.FrameInfo),
                     C.sizeof(A.Params), C.sizeof(A.Config), C.sizeof(A.Results)]


def test_default_params_are_the_reference_literals():
    p = rb.default_params()  # executable/main.cpp:172-176
    assert (p.target, p.lower_bound, p.tilt_max, p.ratio_min, p.ratio_max) == (1, 80, 70.0, 1.5, 80.0)
    assert (p.area_min, p.area_max, p.angle_difference_max, p.shear_max) == (10.0, 99999.0, 12.0, 22.0)
    assert abs(p.lenght_ratio_max - 0.4) < 1e-7


def test_status_strings():
    lib = rb.load_library()
    assert lib.rmcv_status_string(0) == b"ok"
    assert b"capacity" in lib.rmcv_status_string(A.RMCV_ERR_CAPACITY)


def _device_count():
    n = C.c_int(0)
    rb.load_library().rmcv_device_count(C.byref(n))
    return n.value


@pytest.mark.skipif(_device_count() > 0, reason="a CUDA device is present")
def test_no_gpu_means_loud_failure_not_fallback():
    with pytest.raises(rb.RmcvError) as ei:
        rb.Context(max_width=64, max_height=64, max_batch=1)
    assert ei.value.status == A.RMCV_ERR_NO_DEVICE
    import numpy as np
    with pytest.raises(rb.RmcvError):
        rb.extract_color(np.zeros((8, 8, 3), np.uint8), rb.CAMP_BLUE, 80)


def test_product_package_never_imports_the_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "rmcv_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(root, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports the oracle"
                assert "cv2.findContours" not in txt and "cv2.fitEllipse" not in txt, f"{f} calls OpenCV algorithms"

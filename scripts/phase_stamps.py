"""GPU aid: batch-1 phase anatomy of the labelling kernels.  Builds a -DRMCV_STAMPS copy of the library into build/stamps/
(clock64 stamps by thread 0 of CTA 0 after each phase; the product library has none) and prints the median phase times."""
import ctypes, glob, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
out = os.path.join(ROOT, "build", "stamps"); os.makedirs(out, exist_ok=True)
lib = os.path.join(out, "librmcv_b200.so")
if "--build" in sys.argv or not os.path.exists(lib):
    cus = sorted(glob.glob(os.path.join(ge.CSRC, "*.cu")))
    subprocess.run([ge._nvcc()] + ge.NVCC_FLAGS + ["-DRMCV_STAMPS", "-o", lib] + cus, check=True, cwd=ge.CSRC)
    if "--build" in sys.argv: sys.exit(0)
import rmcv_b200.api as api
api.lib_path = lambda: lib
import rmcv_b200 as rb
from rmcv_b200 import synth
W, H = 1280, 1024
ctx = rb.Context(max_width=W, max_height=H, max_batch=1)
L = api.load_library()
L.rmcv_debug_stamps.restype = ctypes.c_int; L.rmcv_debug_stamps.argtypes = [ctypes.c_void_p]
frames = [synth.make_frame(s, W, H, synth.plates_for_seed(s)) for s in range(16)]
bufs = []
for f in frames:
    b = ctx.device_buffer(f.nbytes); b.upload(f); bufs.append(b)
d_mask = ctx.device_buffer(H * W)
p = rb.default_params()
for fn in ("rmcv_debug_ns_pixel", "rmcv_debug_ns_emit", "rmcv_debug_ns_frame"):
    getattr(L, fn).restype = ctypes.c_int; getattr(L, fn).argtypes = [ctypes.c_void_p, ctypes.c_int]
def read_ns(reset):
    o = np.zeros((3, 8, 2), np.uint64)
    for j, fn in enumerate(("rmcv_debug_ns_pixel", "rmcv_debug_ns_emit", "rmcv_debug_ns_frame")):
        assert getattr(L, fn)(o[j].ctypes.data, reset) == 0
    # pixel, emit, label, contour, fit, order
    return np.stack([o[0, 0], o[1, 0], o[2, 0], o[2, 1], o[2, 2], o[2, 3]]).astype(np.int64)
import time
L.rmcv_debug_fit_marks.restype = ctypes.c_int; L.rmcv_debug_fit_marks.argtypes = [ctypes.c_void_p]
fm = []
rows = []; ns = []; wall = []
read_ns(1)
for i in range(400):
    t0 = time.perf_counter()
    ctx.detect_batch(bufs[i % 16].ptr, W, H, 1, p, d_mask.ptr); ctx.fetch_results()
    wall.append(1e6 * (time.perf_counter() - t0))
    st = np.zeros((4, 32), np.int64)
    assert L.rmcv_debug_stamps(st.ctypes.data) == 0
    k = read_ns(1)
    marks = np.zeros(16, np.int64); assert L.rmcv_debug_fit_marks(marks.ctypes.data) == 0
    if i >= 100: rows.append(st.copy()); ns.append(k); fm.append(marks)
a = np.stack(rows)
n = np.stack(ns)   # [iter][kernel][begin, end]
kn = ["pixel", "emit", "label", "contour", "fit", "order"]
print("wall p50 %.1f us (with the stamp atomics)" % float(np.median(wall[100:])))
print("kernel us (first CTA begins -> last CTA ends):", {kn[j]: round(float(np.median(n[:, j, 1] - n[:, j, 0])) / 1e3, 2) for j in range(6)})
print("gap us (end of previous -> begin of next):", {kn[j + 1]: round(float(np.median(n[:, j + 1, 0] - n[:, j, 1])) / 1e3, 2) for j in range(5)})
print("pixel begin -> order end: %.2f us" % (float(np.median(n[:, 5, 1] - n[:, 0, 0])) / 1e3))
names = {0: ("label", 12), 1: ("contour (last component of warp 0)", 6), 2: ("fit (last component of thread 0)", 4), 3: ("order", 6)}
fm = np.stack(fm)
print("fit sections, cycles (sums->moments+shift, scatter/M/det, eigenvalues, eigenvectors, ellipse, gates, lightblob):",
      [int(np.median(fm[:, i + 1] - fm[:, i])) for i in range(7)], "total", int(np.median(fm[:, 7] - fm[:, 0])))
h = a[:, 0, :]
print("label hole phase cycles (init, bbox, classify, jump, unions, flatten):",
      [int(np.median(x)) for x in (h[:, 12] - h[:, 6], h[:, 13] - h[:, 12], h[:, 14] - h[:, 13], h[:, 15] - h[:, 14], h[:, 16] - h[:, 15], h[:, 7] - h[:, 16])])
for k, (nm, n) in names.items():
    d = np.diff(a[:, k, :n], axis=1)
    print(nm, "phase cycles (median):", [int(x) for x in np.median(d, axis=0)], "total", int(np.median(a[:, k, n - 1] - a[:, k, 0])))

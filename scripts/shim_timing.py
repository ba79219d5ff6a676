"""GPU aid: the reference's three calls through the rm:: shim (tests/cpp/call_site) vs the fused rm::gpu::detect, ms per frame."""
import json, os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rmcv_b200 import synth
frame = synth.make_frame(21, 1280, 1024, 9)
with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as fh:
    fh.write(np.array([1280, 1024], np.int32).tobytes()); fh.write(frame.tobytes()); path = fh.name
out = json.loads(subprocess.run([os.path.join(ROOT, "tests", "cpp", "call_site"), path], check=True, capture_output=True, text=True).stdout)
os.unlink(path)
print("shim: three-call path %.3f ms/frame, fused rm::gpu::detect %.3f ms/frame (ratio %.2f); reused light blobs %d, armours %d" % (
    out["three_call_ms"], out["fused_ms"], out["three_call_ms"] / out["fused_ms"], out["reused_lightblobs"], out["reused_armours"]))

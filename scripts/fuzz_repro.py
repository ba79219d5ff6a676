import sys, os, numpy as np
sys.path.insert(0, "/root/repo")
import rmcv_b200 as rb
from oracle import rm_oracle as O
fr = np.load("/root/repo/gpurun_out/fuzz_fail_1.npy")
H, W, _ = fr.shape
ref = O.detect_frame(fr, target=1, lower_bound=120)
for B, chunk in ((1, 0), (17, 2), (17, 0), (33, 0)):
    with rb.Context(max_width=W, max_height=H, max_batch=B, chunk_frames=chunk) as c:
        prm = rb.default_params(target=1, lower_bound=120)
        frames = np.stack([fr] * B)
        mask = np.empty((B, H, W), np.uint8)
        res = c.detect_batch_host(frames, prm, mask)
        ns = [res.frames[f].n_contours for f in range(B)]
        print("B", B, "chunk", chunk, "env", os.environ.get("RMCV_SMALL_BATCH"), "contours", sorted(set(ns)), "oracle", len(ref.contours), "mask ok", bool(np.array_equal(mask[0], ref.binary)))
        if ns[0] != len(ref.contours):
            det = c.frame_detections(res, 0)
            got = {tuple(ci.first) for ci in det.contours}
            want = {(int(p[0][0]), int(p[0][1])) for p in ref.contours}
            print("  missing", sorted(want - got), "extra", sorted(got - want))

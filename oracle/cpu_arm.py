"""TEST / BENCH INFRASTRUCTURE ONLY — the reference's CPU path timed on the host cores (bench.py's `cpu_baseline` leg and
`--impl reference` arm).  Never imported by the product path.

What runs: executable/main.cpp:172-176 — rm::extract_color -> rm::filter_lightblobs -> rm::filter_armours — as the
REFERENCE'S OWN compiled C++ (oracle/_ref/librmcv_ref.so, `kind: "reference"`), whose cv:: calls are served by this
image's OpenCV through cv2 (the C++ OpenCV libraries are not on the image).  Where oracle/_ref is missing it falls back
to the Python restatement oracle/rm_oracle.py (`kind: "port"`).

How: one worker PROCESS per host core (no GIL between them), cv2.setNumThreads(1) in each, every worker owns a fixed
slice of the batch (frames are generated from their seeds inside the worker, untimed).  A step = every worker runs the
path once over its slice; step time = wall clock from "go" to the last worker's "done".  Spawned (not forked) so that a
parent holding a CUDA context is irrelevant.
"""
from __future__ import annotations

import json
import multiprocessing as mp
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MAIN = dict(target=1, lower_bound=80, tilt_max=70.0, ratio_min=1.5, ratio_max=80.0, area_min=10.0, area_max=99999.0,
            angle_difference_max=12.0, shear_max=22.0, lenght_ratio_max=0.4)   # executable/main.cpp:172-176


class FramePath:
    """The per-frame path of one process: reference build if present, else the Python port."""

    def __init__(self, force_port: bool = False):
        import ctypes as C
        import cv2
        cv2.setNumThreads(1)
        from oracle import ref_bridge as RB
        self.C = C
        self.kind = "port"
        self.cv_s = 0.0          # time spent inside OpenCV proper (the cv2 calls), for the stage split
        self.cb_calls = 0
        if RB.available() and not force_port:
            self.ref = RB.get()
            self.kind = "reference"
            inner = self.ref._dispatch

            def timed(op, a, p, _inner=inner):
                t = time.perf_counter()
                r = _inner(op, a, p)
                self.cv_s += time.perf_counter() - t
                self.cb_calls += 1
                return r
            self.ref._dispatch = timed
            self.fn = self.ref.lib.rmcv_ref_process_frame
        else:
            from oracle import rm_oracle as O
            self.O = O

    def run(self, frame):
        """-> (contours, positive, negative, armours) counts of one H x W x 3 frame."""
        C = self.C
        if self.kind == "reference":
            counts = (C.c_int32 * 4)()
            rc = self.fn(frame.ctypes.data_as(C.POINTER(C.c_uint8)), frame.shape[0], frame.shape[1], MAIN["target"], MAIN["lower_bound"],
                         C.c_float(MAIN["tilt_max"]), C.c_float(MAIN["ratio_min"]), C.c_float(MAIN["ratio_max"]), C.c_double(MAIN["area_min"]),
                         C.c_double(MAIN["area_max"]), C.c_float(MAIN["angle_difference_max"]), C.c_float(MAIN["shear_max"]),
                         C.c_float(MAIN["lenght_ratio_max"]), counts)
            self.ref._check(rc)
            return tuple(counts)
        fr = self.O.detect_frame(frame)
        return (len(fr.contours), len(fr.positive), len(fr.negative), len(fr.armours))

    def trampoline_overhead_us(self, n=2000) -> float:
        """Cost of one cv:: call of the reference build that is NOT OpenCV work: C++ -> ctypes callback -> numpy views ->
        dispatch -> result marshalling, measured on cv::boundingRect of four points (OpenCV's share is ~1 us)."""
        if self.kind != "reference":
            return 0.0
        import numpy as np
        a, b = (0.0, 0.0), (3.0, 4.0)
        t = time.perf_counter()
        for _ in range(n):
            self.ref.line_center(a, b)          # no cv:: call: ctypes entry cost only
        base = time.perf_counter() - t
        q = np.array([[0, 0], [4, 0], [4, 9], [0, 9]], np.float32)
        t = time.perf_counter()
        for _ in range(n):
            self.ref.calc_perspective(q)        # no cv:: call either
        base2 = time.perf_counter() - t
        blob = self.ref.make_lightblob([10, 10, 4, 20, 5], 1)
        t = time.perf_counter()
        for _ in range(n):
            self.ref.make_armour(blob, blob)    # exactly one cv:: call (boundingRect of the icon)
        one = time.perf_counter() - t
        return max(0.0, (one - max(base, base2)) / n * 1e6)


def make_frame(kind: str, seed: int):
    from rmcv_b200 import synth
    if kind == "config3":
        return synth.make_frame(seed, 1280, 1024, synth.plates_for_seed(seed), blue=True)
    if kind == "config4":
        return synth.make_stress_frame(seed % 4, 4096, 3072, 250)
    raise ValueError(kind)


def _worker(conn, kind, seeds, force_port):
    try:
        path = FramePath(force_port)
        frames = [make_frame(kind, s) for s in seeds]
        for f in frames[:2]:
            path.run(f)
        conn.send(("ready", path.kind))
        while True:
            msg = conn.recv()
            if msg == "stop":
                break
            path.cv_s = 0.0
            path.cb_calls = 0
            t0 = time.perf_counter()
            tot = [0, 0, 0, 0]
            lat = []
            for f in frames:
                t1 = time.perf_counter()
                c = path.run(f)
                lat.append(time.perf_counter() - t1)
                for k in range(4):
                    tot[k] += c[k]
            conn.send(("done", time.perf_counter() - t0, tot, path.cv_s, path.cb_calls, lat if msg == "go-lat" else None))
    except Exception as e:  # noqa: BLE001
        conn.send(("error", repr(e)))


class CpuArm:
    """`procs` worker processes over `n_frames` frames of `kind` (seeds seed0 ..), static slices."""

    def __init__(self, n_frames: int, procs: int, kind: str = "config3", seed0: int = 0, force_port: bool = False):
        ctx = mp.get_context("spawn")
        self.n_frames, self.procs = n_frames, procs
        self.workers = []
        for w in range(procs):
            seeds = list(range(seed0 + w, seed0 + n_frames, procs))
            parent, child = ctx.Pipe()
            p = ctx.Process(target=_worker, args=(child, kind, seeds, force_port), daemon=True)
            p.start()
            self.workers.append((p, parent))
        self.kind = None
        for _, c in self.workers:
            msg = c.recv()
            if msg[0] != "ready":
                raise RuntimeError(f"CPU arm worker failed: {msg}")
            self.kind = msg[1]

    def step(self, latencies: bool = False):
        """One pass of every worker over its slice -> dict(seconds, counts, cv_seconds (summed over workers), ...)."""
        t0 = time.perf_counter()
        for _, c in self.workers:
            c.send("go-lat" if latencies else "go")
        tot, cv_s, busy, calls, lat = [0, 0, 0, 0], 0.0, 0.0, 0, []
        for _, c in self.workers:
            msg = c.recv()
            if msg[0] != "done":
                raise RuntimeError(f"CPU arm worker failed: {msg}")
            busy += msg[1]; cv_s += msg[3]; calls += msg[4]
            for k in range(4):
                tot[k] += msg[2][k]
            if msg[5]:
                lat += msg[5]
        return dict(seconds=time.perf_counter() - t0, counts=tot, busy_seconds=busy, cv_seconds=cv_s, cv_calls=calls, latencies=lat)

    def close(self):
        for p, c in self.workers:
            try:
                c.send("stop")
            except Exception:
                pass
        for p, _ in self.workers:
            p.join(timeout=5)
            if p.is_alive():
                p.kill()


def measure(n_frames: int, procs: int, steps: int, warmup: int = 1, kind: str = "config3", force_port: bool = False):
    arm = CpuArm(n_frames, procs, kind, force_port=force_port)
    try:
        for _ in range(warmup):
            arm.step()
        runs = [arm.step() for _ in range(steps)]
    finally:
        arm.close()
    total = sum(r["seconds"] for r in runs)
    busy = sum(r["busy_seconds"] for r in runs)
    return dict(kind=arm.kind, procs=procs, frames_per_step=n_frames, steps=steps, seconds=total, step_seconds=[r["seconds"] for r in runs],
                fps=n_frames * steps / total, counts=runs[-1]["counts"],
                opencv_share=sum(r["cv_seconds"] for r in runs) / busy if busy > 0 else None,
                cv_calls_per_frame=runs[-1]["cv_calls"] / n_frames)


def side_configs():
    """Single-process CPU figures for BASELINE configs 2, 4 and 5, printed as one JSON object."""
    import cv2
    import numpy as np
    cv2.setNumThreads(1)
    from oracle import rm_oracle as O
    from rmcv_b200 import synth
    out = {"cores": 1, "cpu_count": os.cpu_count(), "opencv": cv2.__version__}
    path = FramePath()
    out["kind"] = path.kind
    out["trampoline_overhead_us_per_cv_call"] = path.trampoline_overhead_us()
    # config 5: per-frame latency of the path, one frame at a time
    frames = [make_frame("config3", s) for s in range(32)]
    for f in frames[:4]:
        path.run(f)
    lat = []
    for rep in range(4):
        for f in frames:
            t = time.perf_counter(); path.run(f); lat.append(1e3 * (time.perf_counter() - t))
    lat.sort()
    out["config5_latency_ms"] = {"p50": lat[len(lat) // 2], "p99": lat[int(len(lat) * 0.99)], "frames": len(lat)}
    # config 2: Bayer front (cv2 bilinear stand-in for the closed SDK) + pixel stage, 1440x1080
    raws = [synth.bgr_to_bayer(synth.make_frame(s, 1440, 1080, 10), synth.BAYER_BG) for s in range(8)]
    t = time.perf_counter()
    for rep in range(3):
        for r in raws:
            O.extract_color_mask(O.bayer_to_bgr(r, 4), 1, 80)
    out["config2_bayer_pixel_stage_ms_per_frame"] = 1e3 * (time.perf_counter() - t) / 24
    # config 4: 4096x3072 stress frames
    sf = [make_frame("config4", s) for s in range(2)]
    path.run(sf[0])
    t = time.perf_counter()
    for f in sf:
        c = path.run(f)
    out["config4_stress_ms_per_frame"] = 1e3 * (time.perf_counter() - t) / len(sf)
    out["config4_counts"] = list(c)
    # interpreter overhead of the Python port's per-contour loop: the loop with a no-op body
    contours, _ = O.extract_color(frames[0], 1, 80)
    t = time.perf_counter()
    for rep in range(200):
        for c_ in contours:
            pass
    out["python_noop_contour_loop_us_per_frame"] = 1e6 * (time.perf_counter() - t) / 200
    return out


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--kind", default="config3")
    ap.add_argument("--port", action="store_true", help="time the Python port instead of the reference build")
    ap.add_argument("--side-configs", action="store_true")
    ap.add_argument("--single-too", action="store_true", help="also measure one process on a bounded sample")
    a = ap.parse_args()
    if a.side_configs:
        print(json.dumps(side_configs()))
    else:
        res = {"all_cores": measure(a.frames, a.procs, a.steps, a.warmup, a.kind, a.port)}
        if a.single_too:
            res["single_core"] = measure(min(a.frames, 48), 1, 1, 1, a.kind, a.port)
        print(json.dumps(res))

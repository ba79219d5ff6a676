#!/usr/bin/env python
"""Benchmark of the rmcv detection hot path on B200 (BASELINE.json metric: frames/s @1280x1024 full detect).

    python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's own C++ (oracle/_ref) on the host cores

One "step" = one pass of the whole path (rm::extract_color -> rm::filter_lightblobs -> rm::filter_armours,
executable/main.cpp:172-176) over one batch of synthetic 1280x1024 BGR frames per GPU.  Weak scaling: every rank owns
its own batch (frames are independent; no data-path collective); torch.distributed is used only for
the barrier around the timed region and the MAX/SUM reduction of times and unit counts.

Prints ONE JSON line on rank 0 (see the contract in the task statement):
  value     frames/s with the frames resident in HBM when the timed region starts (device time, CUDA events on the
            library's streams, max over ranks), results written to pinned host memory inside the timed region
  e2e       frames/s through rmcv_detect_batch_host with HOST (pinned) buffers: H2D of the frames, kernels, results
  roofline  the pixel-stage kernel: algorithmic bytes (3 B/px read + 1 B/px mask written) / its CUDA-event duration
  cpu_baseline  the reference's CPU path (oracle/_ref: its own rm:: C++ over this image's OpenCV) on this box's host cores,
            one worker process per core, single-core and all-core figures (reported baseline, not the target)
  strong    (N > 1) BASELINE config 3 as written: ONE 1024-frame batch cut into N contiguous slices, one per rank
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from rmcv_b200 import shard, synth  # noqa: E402

METRIC = "frames/s @1280x1024 full detect"
UNIT = "frames/s"
W_, H_ = 1280, 1024
ALG_BYTES_PER_FRAME = W_ * H_ * 4  # 3 B/px BGR read + 1 B/px mask written


def workload_name(frames):
    return (f"1280x1024 BGR full detect (extract_color+filter_lightblobs+filter_armours, main.cpp:172-176 parameters), "
            f"{frames} synthetic frames per GPU per step (BASELINE config 3)")


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_per_launch(frames_per_launch):
    """dram__bytes_read.sum + dram__bytes_write.sum of the pixel kernel from the committed `ncu --set full` capture
    (profiles/pixel_traffic.json, written by scripts/ncu_summary.py), scaled to this run's frames per launch."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "pixel_traffic.json")))
        return d["dram_bytes_per_frame"] * frames_per_launch
    except Exception:
        return None


def make_frames(n, seed0, out, threads=None):
    """Fill out[n,H,W,3] with seeded synthetic frames (plates uniform in [4,20], alternating camp is not used:
    the timed call takes one camp per batch like the reference's call site)."""
    def one(i):
        s = seed0 + i
        out[i] = synth.make_frame(s, W_, H_, synth.plates_for_seed(s), blue=True)
    with ThreadPoolExecutor(max_workers=threads or min(32, os.cpu_count() or 8)) as ex:
        list(ex.map(one, range(n)))


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_arm_measure(n_frames, procs, steps, warmup=1):
    """The reference's CPU path on `procs` worker processes over frames of seeds 0..n_frames-1 (oracle/cpu_arm.py)."""
    from oracle import cpu_arm
    return cpu_arm.measure(n_frames, procs, steps, warmup)


def cpu_side_configs():
    """Single-process CPU figures for BASELINE configs 2, 4, 5 + trampoline overhead, in a fresh interpreter."""
    try:
        out = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "cpu_arm.py"), "--side-configs"], capture_output=True, text=True,
                             timeout=300, cwd=ROOT)
        return json.loads(out.stdout.strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)}


def cpu_description(m, cores):
    import cv2
    what = ("the reference's own rm:: C++ (oracle/_ref: src/core.cpp, objdetect.cpp, imgproc.cpp compiled unmodified), its cv:: calls served "
            f"by this image's OpenCV {cv2.__version__} through cv2" if m["kind"] == "reference" else f"Python port oracle/rm_oracle.py over cv2 {cv2.__version__}")
    return (f"{m['frames_per_step']} synthetic 1280x1024 frames (seeds 0..{m['frames_per_step'] - 1}) x {m['steps']} steps, one worker process per core "
            f"({m['procs']} of {cores} host threads, cv2.setNumThreads(1) each); {what}; {m['seconds']:.1f} s wall; "
            f"OpenCV's share of the workers' busy time {100 * (m['opencv_share'] or 0):.0f} %")


def run_reference(args):
    rank, local_rank, world = env_rank()
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n = args.batch
    m = cpu_arm_measure(n, cores, args.steps, max(1, args.warmup))
    single = cpu_arm_measure(min(n, 48), 1, 1, 1)
    fps = m["fps"]
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * m["seconds"] / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(n), "frames_per_step": n, "frames_per_gpu": n,
                   "note": "reference arm = the reference's CPU implementation of the path on the host cores; one step = the whole "
                           f"{n}-frame batch of config 3, frame-parallel over all host threads (the reference itself is single-threaded, main.cpp:55)"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": m["kind"], "sample": cpu_description(m, cores),
                         "single_core": {"value": single["fps"], "frames": single["frames_per_step"]},
                         "scaling_vs_cores": fps / (single["fps"] * cores), "cv_calls_per_frame": m["cv_calls_per_frame"]},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    rank, local_rank, world = env_rank()
    if args.gpus != world and world > 1:
        pass  # torchrun decides; --gpus is informational then
    dist = None
    dev_index = local_rank
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod
    import rmcv_b200 as rb

    B = args.batch
    ctx = rb.Context(max_width=W_, max_height=H_, max_batch=B, device=dev_index, chunk_frames=args.chunk)
    params = rb.default_params()
    # synthetic frames in pinned host memory, then resident in HBM
    t_gen = time.perf_counter()
    pinned = ctx.pinned((B, H_, W_, 3))
    make_frames(B, rank * B, pinned.array)
    t_gen = time.perf_counter() - t_gen
    d_frames = ctx.device_buffer(B * H_ * W_ * 3)
    d_mask = ctx.device_buffer(B * H_ * W_)
    d_frames.upload(pinned.array)

    def barrier():
        ctx.sync()
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    def step_device():
        ctx.detect_batch(d_frames.ptr, W_, H_, B, params, d_mask.ptr)
        return ctx.fetch_results()

    # ---- warm-up, then the timed region (device-resident inputs).  nvidia-smi needs a few hundred ms to come up and the
    # timed region lasts ~10 ms, so the clock sampler starts before the warm-up and the same load is kept up (untimed
    # extra steps) until it has delivered samples, then through the timed steps and the roofline loop.
    sampler = ClockSampler(dev_index)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        res = step_device()
    t_s = time.perf_counter()
    while sampler.proc is not None and len(sampler.lines) < 3 and time.perf_counter() - t_s < 3.0:
        step_device()
    n_blobs = int(res.total_blobs); n_armours = int(res.total_armours); n_contours = int(res.total_contours)
    ctx.profile(True)
    ctx.profile_read(reset=True)
    barrier()
    launches0 = ctx.kernel_launches()
    wall0 = time.perf_counter()
    ctx.timer_start()
    if args.no_pipeline:
        for _ in range(args.steps):
            step_device()
    else:
        # three steps in flight: step k is enqueued before the results of step k-2 are fetched (the library rotates four
        # result sets), every step's results are read on the host inside the timed region.  The deeper queue only hides host
        # jitter (eight ranks share one host at N = 8); the GPU work per step is the same.
        depth = 3
        for k in range(args.steps):
            ctx.detect_batch(d_frames.ptr, W_, H_, B, params, d_mask.ptr)
            if k >= depth - 1:
                res = ctx.fetch_results()
        for _ in range(min(depth - 1, args.steps)):
            res = ctx.fetch_results()
    dev_ms = ctx.timer_stop()
    wall_ms = 1e3 * (time.perf_counter() - wall0)
    launches = ctx.kernel_launches() - launches0
    barrier()
    prof = ctx.profile_read(reset=True)
    ctx.profile(False)
    max_ms, units = shard.reduce_timing(dev_ms, B * args.steps, dist, device=None if dist is None else f"cuda:{local_rank}")
    max_wall, _ = shard.reduce_timing(wall_ms, 0, dist, device=None if dist is None else f"cuda:{local_rank}")
    tot_launches = launches
    if dist is not None:
        import torch
        t = torch.tensor([launches], dtype=torch.int64, device=f"cuda:{local_rank}")
        dist.all_reduce(t)
        tot_launches = int(t.item())
    value = units / (max_ms * 1e-3)

    # ---- pixel-stage kernel alone (same frames, same kernel) for the roofline, CUDA events
    pix_ms = []
    rep_p = 4   # calls per timed window, so that the ~15 us of host enqueue time around the event pair do not count as kernel time
    for i in range(3 + 5):
        ctx.timer_start()
        for _ in range(rep_p):
            ctx.extract_color_batch(d_frames.ptr, W_, H_, B, params.target, params.lower_bound, d_mask.ptr)
        ms = ctx.timer_stop() / rep_p
        if i >= 3:
            pix_ms.append(ms)
    pix_ms_med = statistics.median(pix_ms)
    t_s = time.perf_counter()
    while time.perf_counter() - t_s < 0.25:   # same load until the sampler has covered the end of the timed region
        step_device()
    clocks = sampler.stop()
    clocks["note"] = ("sampled every 100 ms under the bench load: from the last warm-up steps, through the timed steps and the "
                      "pixel-kernel roofline loop, to 0.25 s of the same steps afterwards")
    chunk = ctx.chunk_frames
    n_chunks = -(-B // chunk)
    peak, peak_src = measured_peak_gbs()
    # in-pipeline duration of the pixel kernel (CUDA events around the stage inside the timed steps)
    pix_in_pipe_ms = prof["pixel"][0] / max(prof["pixel"][1], 1)
    ach_iso = ALG_BYTES_PER_FRAME * B / (pix_ms_med * 1e-3) / 1e9
    frames_per_launch = B / n_chunks
    ach_pipe = ALG_BYTES_PER_FRAME * frames_per_launch / (pix_in_pipe_ms * 1e-3) / 1e9 if pix_in_pipe_ms > 0 else None

    # ---- end to end through the host-buffer entry point: H2D of the frames, kernels, D2H of the masks (rm::extract_color
    # returns `binary`, src/imgproc.cpp:74, and the caller uses it, main.cpp:200-204) and of the result records; wall
    # clock, max over ranks.  The figure without the mask download is reported beside it.
    h_masks = ctx.pinned((B, H_, W_))

    def e2e_run(masks, steps):
        for _ in range(2):
            ctx.detect_batch_host(pinned.array, params, masks)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            ctx.detect_batch_host(pinned.array, params, masks)
        ms = 1e3 * (time.perf_counter() - t0)
        barrier()
        mx, units = shard.reduce_timing(ms, B * steps, dist, device=None if dist is None else f"cuda:{local_rank}")
        return units / (mx * 1e-3)
    e2e_value = e2e_run(h_masks.array, args.e2e_steps)
    e2e_nomask = e2e_run(None, max(3, args.e2e_steps // 2))
    rec_bytes = 32 * B + 72 * n_contours + 56 * n_blobs + 112 * n_armours  # frame infos + dense records actually written
    d2h_bytes = rec_bytes + B * H_ * W_

    # ---- BASELINE config 3 as written (N > 1): ONE 1024-frame batch cut into N contiguous slices, one per rank / GPU
    strong = None
    if world > 1:
        lo, hi = shard.frame_slice(args.batch, world, rank)
        nb = hi - lo
        for _ in range(3):
            ctx.detect_batch(d_frames.ptr, W_, H_, nb, params, d_mask.ptr); ctx.fetch_results()
        barrier()
        depth = 3        # calls in flight: a slice is one short launch of each kernel, the library keeps four result sets
        ctx.timer_start()
        for k in range(args.steps):
            ctx.detect_batch(d_frames.ptr, W_, H_, nb, params, d_mask.ptr)
            if k >= depth - 1:
                ctx.fetch_results()
        for _ in range(min(depth - 1, args.steps)):
            ctx.fetch_results()
        s_ms = ctx.timer_stop()
        barrier()
        s_max, s_units = shard.reduce_timing(s_ms, nb * args.steps, dist, device=f"cuda:{local_rank}")
        for _ in range(2):
            ctx.detect_batch_host(pinned.array[:nb], params, h_masks.array[:nb])
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            ctx.detect_batch_host(pinned.array[:nb], params, h_masks.array[:nb])
        se_ms = 1e3 * (time.perf_counter() - t0)
        barrier()
        se_max, se_units = shard.reduce_timing(se_ms, nb * args.e2e_steps, dist, device=f"cuda:{local_rank}")
        strong = {"scaling": "strong", "frames_total": args.batch, "frames_per_gpu": nb, "value": s_units / (s_max * 1e-3), "unit": UNIT,
                  "ms_per_step": s_max / args.steps, "e2e": se_units / (se_max * 1e-3),
                  "note": "BASELINE config 3 as written: one batch of frames_total frames, contiguous slice per rank, device time max over ranks; "
                          "a slice is one launch of each kernel per step (a call is one chunk); three calls in flight", "calls_in_flight": 3}

    # ---- extras on rank 0: BASELINE config 5 (batch-1 latency) and config 2 (Bayer pixel stage)
    extras = None
    if rank == 0 and not args.no_extras:
        extras = {}
        lat = []
        n_lat, n_warm = args.latency_frames, 1000
        for i in range(n_warm + n_lat):
            t0 = time.perf_counter()
            ctx.detect_batch(d_frames.ptr + (i % B) * H_ * W_ * 3, W_, H_, 1, params, d_mask.ptr)
            ctx.fetch_results()
            if i >= n_warm:
                lat.append(1e6 * (time.perf_counter() - t0))
        lat.sort()
        extras["latency_batch1"] = {"p50_us": lat[len(lat) // 2], "p99_us": lat[int(len(lat) * 0.99)], "frames": n_lat, "warmup_frames": n_warm,
                                    "what": "stream of single 1280x1024 frames (seeds 0..B-1 round robin) resident in HBM; enqueue -> results readable in pinned host memory; host wall clock (BASELINE config 5)"}
        lath = []
        for i in range(200 + 2000):
            t0 = time.perf_counter()
            ctx.detect_batch_host(pinned.array[i % B:i % B + 1], params)
            if i >= 200:
                lath.append(1e6 * (time.perf_counter() - t0))
        lath.sort()
        extras["latency_batch1_from_host"] = {"p50_us": lath[len(lath) // 2], "p99_us": lath[int(len(lath) * 0.99)], "frames": len(lath),
                                              "what": "the same from pinned host memory: H2D of 3.9 MB inside the timed region (rmcv_detect_batch_host)"}
        WB_, HB_, NB_ = 1440, 1080, 64
        raw = np.stack([synth.bgr_to_bayer(synth.make_frame(s, WB_, HB_, 10), synth.BAYER_BG) for s in range(8)] * (NB_ // 8))
        ctxb = rb.Context(max_width=WB_, max_height=HB_, max_batch=NB_, device=dev_index)
        # COLD measurement: four distinct input batches and four distinct mask buffers are rotated, so that between two uses
        # of the same 199 MB working set 597 MB of other traffic has gone through the 126 MB L2 (an HBM number, not an L2 one)
        n_rot = 4
        d_raws = [ctxb.device_buffer(raw.nbytes) for _ in range(n_rot)]
        d_mbs = [ctxb.device_buffer(NB_ * HB_ * WB_) for _ in range(n_rot)]
        for k, d in enumerate(d_raws):
            d.upload(np.roll(raw, k, axis=0))
        d_raw, d_mb = d_raws[0], d_mbs[0]
        bms = []
        rep_b = 20   # launches per timed window (the event pair itself costs ~15 us of host enqueue time)
        for i in range(8):
            ctxb.timer_start()
            for r_ in range(rep_b):
                ctxb.bayer_extract_color_batch(d_raws[r_ % n_rot].ptr, WB_, HB_, NB_, synth.BAYER_BG, params.target, params.lower_bound, d_mbs[r_ % n_rot].ptr)
            ms = ctxb.timer_stop() / rep_b
            if i >= 3:
                bms.append(ms)
        bm = statistics.median(bms)
        extras["bayer_pixel_stage"] = {"workload": "64 x 1440x1080 raw BGGR: debayer(B,R) + colour difference + threshold + 3x3 close (BASELINE config 2)",
                                       "ms": bm, "algorithmic_bytes": NB_ * HB_ * WB_ * 2, "achieved_gbs": NB_ * HB_ * WB_ * 2 / (bm * 1e-3) / 1e9,
                                       "frac_of_peak": NB_ * HB_ * WB_ * 2 / (bm * 1e-3) / 1e9 / peak,
                                       "launches_per_timed_window": rep_b,
                                       "note": "average launch duration over back-to-back launches rotating 4 distinct 199 MB working sets (cold: every byte comes from / goes to HBM)"}
        for d in d_raws + d_mbs:
            d.free()
        ctxb.close()
        # the whole path from raw Bayer frames at the size of config 3 (1024 x 1280x1024 mosaics, two calls in flight)
        WF_, HF_, NF_ = 1280, 1024, 1024
        rawf = np.stack([synth.bgr_to_bayer(pinned.array[i], synth.BAYER_BG) for i in range(64)] * (NF_ // 64))
        ctxf = rb.Context(max_width=WF_, max_height=HF_, max_batch=NF_, device=dev_index)
        d_rf = ctxf.device_buffer(rawf.nbytes); d_mf = ctxf.device_buffer(NF_ * HF_ * WF_)
        d_rf.upload(rawf)
        for _ in range(3):
            ctxf.bayer_detect_batch(d_rf.ptr, WF_, HF_, NF_, synth.BAYER_BG, params, d_mf.ptr); fres = ctxf.fetch_results()
        fsteps = 6
        ctxf.timer_start()
        ctxf.bayer_detect_batch(d_rf.ptr, WF_, HF_, NF_, synth.BAYER_BG, params, d_mf.ptr)
        for _ in range(1, fsteps):
            ctxf.bayer_detect_batch(d_rf.ptr, WF_, HF_, NF_, synth.BAYER_BG, params, d_mf.ptr); fres = ctxf.fetch_results()
        fres = ctxf.fetch_results()
        fms = ctxf.timer_stop() / fsteps
        extras["bayer_full_detect"] = {"workload": "1024 x 1280x1024 raw BGGR frames: Bayer front + full detection, two calls in flight",
                                       "ms_per_step": fms, "frames_per_s": NF_ / (fms * 1e-3), "blobs_per_frame": fres.total_blobs / NF_,
                                       "full_path_frac_of_hbm_peak": NF_ * HF_ * WF_ * 2 / (fms * 1e-3) / 1e9 / peak,
                                       "note": "2 B/px algorithmic (1 raw in, 1 mask out)"}
        # the same from HOST mosaics (what the camera delivers): 1 B/px crosses PCIe instead of 3
        p_raw = ctxf.pinned(rawf.shape); p_raw.array[:] = rawf
        p_rm = ctxf.pinned(rawf.shape)
        for _ in range(2):
            ctxf.bayer_detect_batch_host(p_raw.array, synth.BAYER_BG, params, p_rm.array)
        t0 = time.perf_counter()
        for _ in range(5):
            ctxf.bayer_detect_batch_host(p_raw.array, synth.BAYER_BG, params, p_rm.array)
        bh_s = (time.perf_counter() - t0) / 5
        extras["bayer_full_detect"]["e2e_from_host_frames_per_s"] = NF_ / bh_s
        extras["bayer_full_detect"]["e2e_note"] = "rmcv_bayer_detect_batch_host: H2D of 1 B/px raw mosaics + kernels + D2H of masks and records, wall clock"
        p_raw.free(); p_rm.free()
        d_rf.free(); d_mf.free(); ctxf.close()
        # BASELINE config 4: 4096x3072 stress frames (250 plates -> ~500 light blobs, ~125k pairs per frame), batch 16 and 64
        WS_, HS_ = 4096, 3072
        stress = {}
        for NS_ in (16, 64):
            sframes = np.stack([synth.make_stress_frame(s, WS_, HS_, 250) for s in range(4)] * (NS_ // 4))
            ctxs = rb.Context(max_width=WS_, max_height=HS_, max_batch=NS_, device=dev_index, max_blobs_per_frame=1024,
                              max_armours_per_frame=4096)
            d_sf = ctxs.device_buffer(sframes.nbytes); d_sm = ctxs.device_buffer(NS_ * HS_ * WS_)
            d_sf.upload(sframes)
            del sframes
            for _ in range(2):
                ctxs.detect_batch(d_sf.ptr, WS_, HS_, NS_, params, d_sm.ptr); sres = ctxs.fetch_results()
            sms = []
            rep_s = 4
            for i in range(5):
                ctxs.timer_start()
                ctxs.detect_batch(d_sf.ptr, WS_, HS_, NS_, params, d_sm.ptr)
                for _ in range(1, rep_s):
                    ctxs.detect_batch(d_sf.ptr, WS_, HS_, NS_, params, d_sm.ptr); sres = ctxs.fetch_results()
                sres = ctxs.fetch_results()
                sms.append(ctxs.timer_stop() / rep_s)
            sm_ = statistics.median(sms)
            stress[f"batch{NS_}"] = {"ms_per_call": sm_, "frames_per_s": NS_ / (sm_ * 1e-3),
                                     "contours_per_frame": sres.total_contours / NS_, "blobs_per_frame": sres.total_blobs / NS_,
                                     "armours_per_frame": sres.total_armours / NS_,
                                     "full_path_frac_of_hbm_peak": NS_ * HS_ * WS_ * 4 / (sm_ * 1e-3) / 1e9 / peak}
            d_sf.free(); d_sm.free(); ctxs.close()
        stress["workload"] = "4096x3072 BGR full detect, 250 plates per frame (BASELINE config 4: batch >= 16)"
        stress["note"] = ("two calls in flight; 4 B/px algorithmic (3 in, 1 mask out); the label and order kernels run as thread-block "
                          "clusters of 2-8 CTAs per frame when a chunk has fewer frames than the GPU has SMs")
        extras["stress_4096x3072"] = stress

    # ---- the library's own partition of ONE host batch across every visible GPU (one process; rmcv_multi_*)
    if extras is not None and world == 1:
        try:
            import ctypes as C_
            nvis = C_.c_int(0)
            rb.load_library().rmcv_device_count(C_.byref(nvis))
            if nvis.value > 1:
                with rb.MultiContext(None, max_width=W_, max_height=H_, max_batch=B) as mc:
                    for _ in range(2):
                        mc.detect_batch_host(pinned.array, params, h_masks.array)
                    t0 = time.perf_counter()
                    for _ in range(5):
                        mc.detect_batch_host(pinned.array, params, h_masks.array)
                    dt = (time.perf_counter() - t0) / 5
                    extras["multi_gpu_one_process"] = {"devices": mc.n_devices, "frames": B, "e2e_frames_per_s": B / dt,
                                                       "what": "rmcv_multi_detect_batch_host: one host batch, contiguous slice per device, one host thread + ctx per device"}
        except Exception as e:  # noqa: BLE001
            extras["multi_gpu_one_process"] = {"error": repr(e)}

    # ---- CPU baseline on rank 0 at N == 1
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        m = cpu_arm_measure(B, cores, args.cpu_steps, 1)          # the same batch (seeds 0..B-1), all cores
        single = cpu_arm_measure(min(B, 48), 1, 1, 1)             # one core, bounded sample
        cpu = {"value": m["fps"], "unit": UNIT, "cores": cores, "kind": m["kind"], "sample": cpu_description(m, cores),
               "single_core": {"value": single["fps"], "frames": single["frames_per_step"]},
               "scaling_vs_cores": m["fps"] / (single["fps"] * cores), "cv_calls_per_frame": m["cv_calls_per_frame"],
               "side_configs": cpu_side_configs()}

    if rank == 0:
        stage_ms = {k: v[0] / args.steps for k, v in prof.items()}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_name(B),
                       "inputs": f"resident in HBM ({B * H_ * W_ * 3 / 1e9:.1f} GB per GPU > 126 MB L2, no flush needed)",
                       "frames_per_gpu": B, "chunk_frames": chunk, "parallelism": f"frame-sharded x{world}, no collective",
                       "steps_in_flight": 1 if args.no_pipeline else 3,
                       "contours_per_frame": n_contours / B, "blobs_per_frame": n_blobs / B, "armours_per_frame": n_armours / B},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * H_ * W_ * 3, "d2h_bytes_per_step": d2h_bytes,
                    "steps": args.e2e_steps, "without_mask_download": e2e_nomask,
                    "note": "rmcv_detect_batch_host from pinned host frames into pinned host masks + result records; wall clock, max over ranks"},
            "strong": strong,
            "gpu_launches": tot_launches,
            "roofline": {"bound": "hbm", "achieved": ach_iso, "peak": peak, "unit": "GB/s", "frac": ach_iso / peak,
                         "traffic": ncu_traffic_per_launch(frames_per_launch),
                         "algorithmic_bytes_per_launch": ALG_BYTES_PER_FRAME * frames_per_launch,
                         "kernel": "pixel_bgr_kernel<true, Geom1280> (fused diff/threshold/close)", "peak_source": peak_src,
                         "peak_note": "peak = measured copy bandwidth (1 read : 1 write); this kernel reads 3 bytes per byte it "
                                      "writes, and reads are cheaper for HBM than writes, so frac may exceed 1 by a few per cent",
                         "launch_ms": pix_ms_med / n_chunks, "frames_per_launch": frames_per_launch,
                         "algorithmic_bytes_per_frame": ALG_BYTES_PER_FRAME,
                         "in_pipeline": {"achieved": ach_pipe, "launch_ms": pix_in_pipe_ms,
                                         "note": "same kernel timed with CUDA events inside the timed steps, where it overlaps the other slot's labelling kernels"},
                         "full_path_frac": value / world * ALG_BYTES_PER_FRAME / 1e9 / peak},
            "cpu_baseline": cpu,
            "extras": extras,
            "clocks": clocks,
            "stage_ms_per_step": stage_ms,
            "wall_ms_per_step": max_wall / args.steps,
            "frame_generation_s": t_gen,
        }
        emit(line)
    h_masks.free(); pinned.free(); d_frames.free(); d_mask.free()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line goes to the real stdout; everything else a library prints there (e.g. NCCL's version banner)
    has been redirected to stderr."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="frames per GPU per step")
    ap.add_argument("--chunk", type=int, default=0, help="frames per internal pipeline chunk (0 = library default)")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--cpu-steps", type=int, default=2, help="steps of the cpu_baseline leg (each = the whole batch on all cores)")
    ap.add_argument("--latency-frames", type=int, default=10000, help="timed frames of the batch-1 latency extra (config 5)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the latency (config 5) and Bayer (config 2) side measurements")
    ap.add_argument("--no-pipeline", action="store_true", help="fetch every step's results before enqueueing the next step")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()

"""GPU: seeded slices of the randomised campaigns (scripts/fuzz_*.py) as tests, so that the driver's `pytest -m gpu` run
carries them and their logs are kept (gpurun_out/fuzz_*.log; summaries are copied to profiles/ per round).  Every frame
goes through the full comparator of tests/_compare.py against the oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")


def run_script(name, args, env=None, tag=""):
    os.makedirs(OUT, exist_ok=True)
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", name)] + [str(a) for a in args], capture_output=True, text=True,
                       timeout=900, env=e, cwd=ROOT)
    log = os.path.join(OUT, f"fuzz_{name[:-3]}{tag}.log")
    with open(log, "w") as fh:
        fh.write(p.stdout[-20000:] + "\n--- stderr ---\n" + p.stderr[-4000:])
    assert p.returncode == 0, f"{name} {args}: exit {p.returncode}\n{p.stdout[-3000:]}\n{p.stderr[-2000:]}"
    return p.stdout


def test_fuzz_whole_path_and_bayer_front():
    """~200 frames of random sizes (3..1440 wide), batches 1..33, both camps, three thresholds, four Bayer layouts."""
    out = run_script("fuzz_gpu.py", [30, 20261018])
    assert " 0 mismatches" in out


def test_fuzz_masks_against_findcontours():
    """noise / blob masks and drawn shapes (rings in rings, C-shapes, spirals, combs): external contours and label maps."""
    out = run_script("fuzz_masks_gpu.py", [40, 77], env={"FUZZ_FULL": "1"})
    assert "0 mismatches" in out
    out = run_script("fuzz_masks_gpu.py", [20, 78], env={"FUZZ_SHAPES": "1", "FUZZ_FULL": "1"}, tag="_shapes")
    assert "0 mismatches" in out
    out = run_script("fuzz_masks_gpu.py", [3, 79], env={"FUZZ_BIG": "1"}, tag="_big")
    assert "0 mismatches" in out


def test_fuzz_legacy_rows():
    run_script("fuzz_legacy_gpu.py", [20, 5])


def test_fuzz_stress_frames():
    """frames up to 4096x3072 with hundreds of light blobs, every frame through the comparator."""
    run_script("fuzz_stress_gpu.py", [4, 9])


def test_fused_pixel_emit_kernel_opt_in():
    """RMCV_FUSED_EMIT=1 (the band kernel with the labelling stage's emission folded in; tuning is read once per process, so
    the check runs in its own interpreter): five launches per chunk and the same results, frame for frame."""
    out = run_script("fused_emit_parity_gpu.py", [])
    assert " 0 mismatches" in out


def test_latency_mode_changes_no_result_bit():
    """Small chunks run as a chain of programmatic dependent launches on the slot stream, with the fits on the contour kernel's
    warps and 8-row emit bands.  The same batches with all of that switched off (tuning is read once per process,
    hence two interpreters) must give the same digest over every record and mask byte."""
    a = run_script("latency_mode_digest_gpu.py", [], tag="_default")
    b = run_script("latency_mode_digest_gpu.py", [], env={"RMCV_CHAINED": "0", "RMCV_FIT_IN_CONTOUR": "0", "RMCV_EMIT_BH": "32"}, tag="_plain")
    c = run_script("latency_mode_digest_gpu.py", [], env={"RMCV_WARP_FIT": "0"}, tag="_lane0")   # fits on the contour warps, lane 0 alone
    da, db, dc = a.strip().splitlines()[-1], b.strip().splitlines()[-1], c.strip().splitlines()[-1]
    assert "digest over" in da and da == db == dc, (da, db, dc)


def test_fuzz_calls_in_flight():
    """random sizes, batches, chunk sizes and depths: device-path calls with up to three in flight, each result and mask byte for
    byte what the synchronising host entry point returns for the same batch."""
    out = run_script("fuzz_inflight_gpu.py", [25, 4242])
    assert " 0 mismatches" in out

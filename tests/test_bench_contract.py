"""bench.py contract checks that need no GPU: the reference arm (the reference's CPU path = oracle/_ref on worker processes) prints exactly
one JSON line on stdout with the keys the driver reads, and non-zero ranks of a multi-rank reference run stay silent."""
import json
import os
import subprocess
import sys

ROOT_ = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT_ not in sys.path:
    sys.path.insert(0, ROOT_)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=600, env=e, cwd=ROOT)


def test_reference_arm_prints_one_json_line():
    r = run_bench(["--impl", "reference", "--steps", "1", "--warmup", "0", "--batch", "8"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["metric"] == "frames/s @1280x1024 full detect" and d["value"] > 0 and d["n_gpus"] == 1
    from oracle import ref_bridge
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_bridge.available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["single_core"]["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "1280x1024 BGR full detect" in d["config"]["workload"]


def test_reference_arm_other_ranks_exit_silently():
    r = run_bench(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--batch", "8"],
                  env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""

"""CPU test of bench.py's reference arm: one JSON line on stdout with the contract's keys (the GPU arm prints the same keys plus
roofline / clocks / gpu_launches and is exercised on the B200 by the driver)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-batch", "4"] + extra,
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    return json.loads(lines[0])


def test_reference_arm_prints_one_contract_line():
    d = _run([])
    assert d["impl"] == "reference" and d["metric"] == "clip_train_samples_per_sec" and d["unit"] == "samples/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "per_gpu_batch" in d["config"]
    # the reference arm runs the unmodified reference when its staged copy (oracle/_ref) or /root/reference is present
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "MANIFEST.json")) or os.path.isdir("/root/reference/src"):
        assert d["cpu_baseline"]["kind"] == "reference"
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_reference_arm_other_ranks_print_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_image_workload():
    d = _run(["--workload", "c5"])
    assert "trimodal" in d["config"]["workload"]

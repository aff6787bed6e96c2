"""bench.py contract on CPU: the reference arm runs without a GPU on a tiny workload and prints exactly one JSON line
with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--scan-points", "3000", "--map-points", "30000"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout          # the ikd-Tree's own printf()s must not reach stdout
    o = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in o, k
    assert o["impl"] == "reference" and o["value"] > 0 and o["higher_is_better"] is True and o["vs_baseline"] is None
    assert o["cpu_baseline"]["kind"] in ("reference", "port") and o["cpu_baseline"]["cores"] >= 1
    assert o["e2e"]["h2d_bytes_per_step"] == 0 and o["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in o["config"] and "model" not in o["config"]


def test_nonzero_rank_of_reference_arm_is_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""

"""bench.py's output contract.  CPU: the reference arm (the restated CPU path on host processes) prints one JSON line
with the keys the driver reads.  GPU: the default arm on a small batch carries value / e2e / roofline / clocks /
gpu_launches and the batch sweep."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout=900):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-decisions", "2"])
    assert r["impl"] == "reference" and r["unit"] == "decisions/s" and r["higher_is_better"] is True
    assert r["value"] > 0 and r["cpu_baseline"]["kind"] == "port" and r["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert r["e2e"] == {"value": r["value"], "unit": "decisions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert r["config"]["workload"].startswith("cfg2") and r["vs_baseline"] is None


@pytest.mark.gpu
def test_default_arm_line():
    r = _run(["--batch", "128", "--e2e-batch", "128", "--steps", "3", "--warmup", "3", "--no-cpu"])
    assert r["metric"].startswith("routing decisions/sec") and r["n_gpus"] == 1 and r["scaling"] == "weak"
    assert r["value"] > 0 and r["e2e"]["value"] > 0 and r["e2e"]["h2d_bytes_per_step"] > 0 and r["e2e"]["d2h_bytes_per_step"] > 0
    assert r["gpu_launches"] >= 3 * 8                      # at least eight of our kernels per decision step
    roof = r["roofline"]
    assert roof["bound"] == "fp32" and 0 < roof["frac"] < 2 and roof["peak"] > 30 and roof["unit"] == "TFLOP/s"
    assert roof["engine"] in ("tc", "ffma") and roof["launch_ms"] > 0
    assert set(r["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert [s["episodes"] for s in r["extras"]["batch_sweep"]] == [1, 16, 256, 1024, 4096, 16384]

"""The reference arm of bench.py (``--impl reference``: the oracle port of the path timed on the host cores) keeps the JSON contract of
the benchmark line for all three workloads.  Small shapes, so the three runs take well under a minute; no GPU involved."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
            "config", "cpu_baseline", "e2e")


@pytest.mark.parametrize("workload,unit,higher", [("train", "utt/s", True), ("decode", "RTF", False), ("mbr", "utt/s", True)])
def test_reference_arm_line(workload, unit, higher):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload, "--steps", "1", "--warmup", "0",
           "--T", "300", "--U", "12", "--V", "64", "--beam", "4", "--cpu-threads", "4"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    for k in REQUIRED:
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == unit and d["higher_is_better"] is higher and d["data"] == "synthetic"
    assert d["value"] is not None and d["value"] > 0
    cb, e2e = d["cpu_baseline"], d["e2e"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and isinstance(cb["sample"], str) and cb["sample"]
    assert e2e["value"] == d["value"] and e2e["unit"] == unit and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    """under torchrun only rank 0 runs and prints the reference line (the driver launches both arms the same way)"""
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""

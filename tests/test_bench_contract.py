"""The reference arm of bench.py runs on the CPU: its JSON line carries the contract's keys."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def test_reference_arm_json_line():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "PatchyParticles")):
        pytest.skip("oracle/_ref not built")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "64k", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().split("\n")[-1])
    assert line["impl"] == "reference" and line["metric"] == "MC roto-translation attempts/sec" and line["unit"] == "attempts/s"
    assert line["value"] > 1e4 and line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 1
    assert line["vs_baseline"] is None and line["dtype"] == "f64" and line["data"] == "synthetic"
    assert "N=65536" in line["config"]["workload"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] == 1 and cb["value"] == line["value"] and "MC steps" in cb["sample"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["unit"] == line["unit"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""

"""CPU tier: the reference arm of bench.py (`--impl reference`: the unmodified reference's own
CPU path on this box's cores) prints ONE JSON line with the contract's keys.  The GPU arm needs
a device and is run by the driver; here only its argument parsing is checked."""
import json
import os
import subprocess
import sys

import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not oracle.reference_available(), reason="oracle/_ref not built")
def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
                "higher_is_better", "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in rec, key
    assert rec["unit"] == "points/s" and rec["value"] > 0 and rec["higher_is_better"] is True
    assert rec["cpu_baseline"]["kind"] == "reference" and rec["cpu_baseline"]["cores"] >= 1
    assert rec["e2e"]["h2d_bytes_per_step"] == 0 and rec["e2e"]["d2h_bytes_per_step"] == 0
    assert rec["e2e"]["value"] == rec["value"]
    assert "N_live=4000" in rec["metric"] and rec["config"]["n_live"] == 4000 and rec["config"]["ndim"] == 20
    # both arms print the SAME config object (the driver compares them)
    sys.path.insert(0, ROOT)
    import bench
    assert rec["config"] == bench.workload_config(1, 1 << 20)
    assert "workload" in rec["config"] and "l2_policy" in rec["config"]


def test_bench_help_runs_without_a_gpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl"):
        assert flag in out.stdout

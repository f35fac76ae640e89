"""bench.py's CPU arm (`--impl reference`) runs without a GPU and prints one well-formed JSON line."""
import json
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-seconds", "0.5",
                          "--frames", "500"], capture_output=True, text=True, timeout=300, cwd=ROOT)  # fmt: skip
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    base = json.loads((ROOT / "BASELINE.json").read_text())
    assert line["impl"] == "reference" and line["metric"] == base["metric"] and line["unit"] == "frames/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    import os

    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True,
                         timeout=120, cwd=ROOT, env=env)  # fmt: skip
    assert out.returncode == 0 and out.stdout.strip() == ""

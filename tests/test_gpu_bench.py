"""bench.py on a GPU: the JSON line carries every key of the measurement contract."""
import json
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_bench_line_contract():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "3", "--frames", "2000", "--cpu-seconds", "1",
                          "--big-frames", "20000", "--big-steps", "1", "--fit-frames", "500"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)  # fmt: skip
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, "bench.py must print exactly one JSON line"
    d = json.loads(lines[0])
    base = json.loads((ROOT / "BASELINE.json").read_text())
    assert d["metric"] == base["metric"] and d["unit"] == "frames/s" and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] >= 3
    assert d["higher_is_better"] is True and d["scaling"] == "strong" and d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] == 1
    assert "workload" in d["config"] and "model" not in d["config"] and "l2" in d["config"]
    r = d["roofline"]
    assert r["unit"] == "TFLOP/s" and 0 < r["achieved"] < r["peak"] and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["hbm"]["algorithmic_bytes_per_launch"] > 0 and 0 < r["flops_executed"] < r["flops_reference_sequence"]
    assert 0 < r["frac_executed"] < r["frac"] and r["peak_nominal"] > 50
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and "sample" in c
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 2000 * 69 * 4 and e["d2h_bytes_per_step"] > e["h2d_bytes_per_step"]
    assert e["value"] <= d["value"] * 1.05  # end to end cannot beat the device-resident number
    p = d["parity"]
    assert p["vs_kernel_order_f32"]["bit_identical"] is True and p["vs_kernel_order_f32"]["qpos_abs_rad"]["max"] == 0.0
    m = p["vs_mjx_order_f32"]  # another float32 order of the same algorithm: the typical frame meets north_star's tolerances
    assert m["marker_abs_m"]["median"] <= 1e-4 and m["qpos_abs_rad"]["median"] <= 1e-3 and m["marker_rmse_m"] <= 1e-3
    assert d["config5_1e6_frames"]["value"] > 0 and d["config5_1e6_frames"]["scaling"] == "strong"
    assert d["config4_fruitfly"]["value"] > 0 and d["config5_mouse"]["value"] > 0 and "register-resident" in d["config5_mouse"]["workload"]
    assert d["config2_weak"]["value"] > 0 and d["config3_fit"]["seconds"] > 0 and d["config3_fit"]["offsets_finite"] is True
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}

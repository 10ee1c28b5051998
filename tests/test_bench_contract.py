"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`) runs the CPU oracle and prints ONE
JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--cpu-rays", "16"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train_rays_per_s" and d["unit"] == "rays/s"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1 and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "f32"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1", "--cpu-rays", "16"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    assert p.stdout.strip() == ""


def test_step_roofline_arithmetic():
    """The whole-step roofline of SURVEY.md 8d from samples per ray (host arithmetic only)."""
    import bench
    peaks = {"hbm_gbs": 6553.9, "bf16_tflops_sustained": 1344.2}
    r = bench.step_roofline(189.9, 80.8, 292946.0, peaks)
    fg_bytes = 13 * (1164 + 1164) + 6 * 1176
    assert fg_bytes == 37320
    want_bytes = 189.9 * fg_bytes + 80.8 * 2328 + 56
    assert abs(r["algorithmic_bytes_per_ray"] - want_bytes) < 1e-6
    fg_flop = 3 * (20992 + 12 * 12800 + 19712)
    bg_flop = 3 * (5504 + 11648)
    assert abs(r["algorithmic_flop_per_ray"] - (189.9 * fg_flop + 80.8 * bg_flop)) < 1e-3
    assert r["bound"] == "hbm" and abs(r["rays_per_s_roofline"] - 6553.9e9 / want_bytes) < 1e-3
    assert 0.30 < r["frac"] < 0.36
    assert r["rays_per_s_by_tensor"] > 10 * r["rays_per_s_by_hbm"]
    a = bench.step_roofline(189.9, 80.8, 190918.0, peaks, "analytic")
    assert a["algorithmic_bytes_per_ray"] < r["algorithmic_bytes_per_ray"]

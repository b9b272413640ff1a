"""CPU: the bench.py contract the round driver depends on -- the reference arm's JSON line, the loud failure of the product
arm without a CUDA device (no CPU fallback), and the internal consistency of the committed bench lines under profiles/."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e"}


def _run(*args, timeout=600):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout,
                          cwd=ROOT, env=env)


@pytest.mark.timeout(900)
def test_reference_arm_prints_one_contract_line():
    out = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--no-lidal")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "scans/sec" and d["unit"] == "scans/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["config"]["workload"] == "spvcnn_inference_SK_batch8" and "model" not in d["config"]
    assert d["value"] > 0 and abs(d["value"] * d["ms_per_step"] / 1e3 - 1.0) < 1e-6          # one scan per step
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_on_other_ranks_exits_quietly():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_product_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    out = _run("--steps", "1", "--warmup", "1", timeout=300)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
    assert not [ln for ln in out.stdout.splitlines() if ln.startswith("{")]                   # no number without the CUDA path


@pytest.mark.parametrize("name", ["r02_bench_engine_final.json", "r02_bench_8gpu_s2.json", "r02_bench_engine_s3.json"])
def test_committed_bench_lines_are_consistent(name):
    path = os.path.join(ROOT, "profiles", name)
    text = [ln for ln in open(path).read().splitlines() if ln.startswith("{")]
    d = json.loads(text[-1])
    assert BASE_KEYS <= set(d) and d["unit"] == "scans/s" and d["scaling"] == "weak"
    assert abs(d["value"] - d["n_gpus"] * d["steps"] * 8 / (d["ms_per_step"] * d["steps"] / 1e3)) < 1e-6 * d["value"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    assert d["gpu_launches"] > 0
    clk = d["clocks"]
    assert clk["sm_mhz"] >= 0.95 * clk["sm_max_mhz"] and not set(clk["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    # achieved = algorithmic FLOPs of a step / the kernel's time in that step, and the kernel fits inside the step
    assert abs(r["achieved"] - r["algorithmic_gflop_per_step"] / r["kernel_ms_per_step"]) < 1e-6 * r["achieved"]
    assert r["kernel_ms_per_step"] < d["ms_per_step"] * 1.25
    if r.get("traffic"):
        assert r["traffic"] < 2.0 * r["algorithmic_bytes_per_launch"]                          # no wasted DRAM re-reads
    if "lidal" in d and "error" not in d["lidal"]:
        li = d["lidal"]
        assert li["unit"] == "frames/s" and li["scaling"] == "strong" and abs(li["value"] - li["frames"] / (li["ms_total"] / 1e3)) < 1e-6 * li["value"]
        assert set(li["collective"]) >= {"halo_ms", "all_gather_ms"}

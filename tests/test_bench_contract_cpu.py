"""bench.py contract checks that need no GPU: the reference arm prints one well-formed JSON line,
and the main arm refuses to run (loudly) without a CUDA device."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_the_contract_line():
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                                   "--warmup", "1", "--cpu-sample-tiles", "1"], cwd=ROOT, timeout=600).decode()
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "tiles/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    # the staged UNMODIFIED reference (baseline/_ref) when present, else the oracle port; BHSR_CPU_ARM=port forces the port
    from oracle import stage_reference
    assert line["cpu_baseline"]["kind"] == ("reference" if stage_reference.available() else "port")
    assert line["cpu_baseline"]["cores"] >= 1
    for k in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert k in line, k
    assert "workload" in line["config"]


def test_main_arm_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        return
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], cwd=ROOT, capture_output=True, timeout=300)
    assert p.returncode != 0 and b"CUDA" in (p.stderr + p.stdout)


def test_committed_gpu_bench_line_carries_the_whole_contract():
    """The last bench line measured on a B200 (profiles/r02_bench_final.json) has every key the
    driver reads, and its derived numbers are self-consistent."""
    path = os.path.join(ROOT, "profiles", "r02_bench_final.json")
    line = json.loads([ln for ln in open(path) if ln.startswith("{")][-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in line, k
    assert line["unit"] == "tiles/s" and line["scaling"] == "weak" and line["vs_baseline"] is None
    assert line["steps"] >= 1 and line["warmup"] >= 3 and line["gpu_launches"] == 356 * line["steps"]
    b = line["config"]["batch_per_gpu"]
    assert abs(line["value"] - b * line["n_gpus"] / line["ms_per_step"] * 1e3) < 1e-6 * line["value"]
    e2e = line["e2e"]
    assert e2e["h2d_bytes_per_step"] == b * 6 * 64 * 64 * 4 and e2e["d2h_bytes_per_step"] > 0 and e2e["value"] > 0
    r = line["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and r["peak"] > 0
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] > 0
    assert abs(r["step"]["achieved"] - 146.630 * b / line["ms_per_step"]) < 1e-6 * r["step"]["achieved"]
    assert {k["layer"] for k in r["kernels"]} == {f"rdb.conv{i}" for i in range(1, 6)}
    c = line["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and "sample" in c
    # the dominant kernel is the one with the largest share of the step, its figure is a mean over its layers
    assert r["kernel"].startswith("conv_dx_kernel") and 0.3 < r["kernel_share_of_step"] < 0.7
    assert abs(r["achieved"] - r["algorithmic_gflop_per_launch"] / r["us_per_launch"] * 1e3) < 1e-6 * r["achieved"]
    assert r["step"]["sum_of_kernels_ms"] < line["ms_per_step"]
    # fwd+bwd sub-record (BASELINE configs[2] / [3])
    t = line["train"]
    for k in ("metric", "value", "unit", "ms_per_step", "steps", "warmup", "batch_per_gpu", "global_batch", "launch",
              "collective", "grad_bucket_floats", "roofline"):
        assert k in t, k
    assert abs(t["value"] - t["global_batch"] / t["ms_per_step"] * 1e3) < 1e-6 * t["value"]
    assert t["roofline"]["algorithmic_gflop_per_tile"] == 169.63
    if line["n_gpus"] > 1:      # one flat all-reduce per step; none at N = 1
        assert t["collective"]["bytes"] == 4 * t["grad_bucket_floats"]
    else:
        assert t["collective"] is None
    assert set(line["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(line["clocks"]["reasons"])

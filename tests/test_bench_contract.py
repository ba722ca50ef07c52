"""bench.py's reference arm runs on CPU (the compiled reference or the oracle port): check the JSON
line it prints against the contract keys. The GPU arm's line is checked on the GPU box (test_gpu_bench)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config"}


def _run(*args, env=None):
    e = dict(os.environ, **(env or {}))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                         timeout=600, env=e, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout


def test_reference_arm_prints_one_contract_line():
    lines = [l for l in _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--rows", "100000").splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert KEYS <= set(d)
    assert d["impl"] == "reference" and d["metric"] == "qxmatch_sources_per_sec" and d["unit"] == "sources/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "sample" in d["config"]


def test_reference_arm_other_ranks_are_silent():
    assert _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--gpus", "2", "--rows", "100000",
                env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}).strip() == ""


def test_cpu_baseline_leg():
    sys.path.insert(0, ROOT)
    import bench
    b = bench.measure_cpu_baseline(budget_s=1.0)
    assert b["unit"] == "sources/s" and b["kind"] in ("reference", "port") and b["cores"] >= 1
    assert b["value"] > 0 and json.dumps(b)
    if b["cores"] > 1:
        assert 0 < b["value_1_thread"] < 1.5*b["value"]


@pytest.mark.gpu
def test_gpu_bench_line_has_roofline_and_e2e():
    lines = [l for l in _run("--steps", "3", "--warmup", "3", "--rows", "1000000", "--no-cpu").splitlines()
             if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert KEYS <= set(d)
    assert d["gpu_launches"] > 0 and d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0
    r = d["roofline"]
    assert r["bound"] == "hbm" and 0 < r["frac"] < 1 and r["unit"] == "GB/s"
    assert abs(r["achieved"] - r["bytes"] / (d["ms_per_step"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]     # whole path, SURVEY 8(d)
    k = r["dominant_kernel"]
    assert 0 < k["frac"] < 1 and k["ms"]["forward_launch"] > 0 and k["ms"]["reverse_launch"] > 0
    assert d["clocks"] is not None and d["step_ms_rank0"]["min"] > 0


@pytest.mark.gpu
def test_gpu_bench_c4_line_reports_the_fp64_roofline():
    lines = [l for l in _run("--config", "c4", "--steps", "2", "--warmup", "3", "--rows", "200000", "--no-e2e").splitlines()
             if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    r = d["roofline"]
    assert r["bound"] == "fp64" and r["unit"] == "TFLOP/s" and r["achieved"] > 0 and r["peak"] > 10

"""bench.py's contract where it can be checked without a GPU: the reference arm prints one JSON line
with the agreed keys (the CPU port of the reference's algorithm, all host threads, bounded sample),
and the product arm refuses to run without a CUDA device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout=240):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


@pytest.mark.parametrize("workload", ["ed25519_mul_base", "x25519", "p256_mul"])
def test_reference_arm_prints_the_contract_line(workload):
    r = _run(["--impl", "reference", "--workload", workload, "--steps", "1", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["config"]["workload"] == workload and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["metric"] == "scalar-mults/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == (os.cpu_count() or 1) and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_on_other_ranks_prints_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_product_arm_refuses_to_run_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = _run(["--steps", "1", "--warmup", "1", "--no-cpu"])
    assert r.returncode != 0 and "no CPU fallback" in r.stderr and not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


@pytest.mark.gpu
def test_product_arm_prints_the_contract_line_on_a_gpu():
    """One short run of the default workload on cuda:0: one JSON line with the base contract's keys plus
    roofline / cpu_baseline / e2e / clocks / gpu_launches, a green parity check of that very run, and
    kernel launches counted (a zero would mean a fallback)."""
    r = _run(["--steps", "3", "--warmup", "3", "--extra", ""], timeout=600)
    assert r.returncode == 0, r.stderr[-800:]
    lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert key in d, key
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] >= 3 and d["scaling"] == "weak" and d["data"] == "synthetic"
    # the headline is BASELINE.json configs[0] exactly: Ed25519 mul_base on 2^16 scalars, several passes per step
    assert d["config"]["workload"] == "ed25519_mul_base_2p16" and d["config"]["batch_per_gpu"] == 1 << 16 and d["parity_check"] is True
    bps = d["step"]["batches_per_step"]
    assert bps >= 1 and d["step"]["timed_region_s"] > 0.4
    assert d["gpu_launches"] >= bps * d["steps"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic", "frac_executed", "hbm_frac"):
        assert key in d["roofline"], key
    assert 0 < d["roofline"]["frac_executed"] < 1
    assert d["roofline"]["peak"] > 1 and abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9
    e = d["e2e"]
    assert e["h2d_bytes_per_batch"] == 2 << 20 and e["d2h_bytes_per_batch"] == 4 << 20 and 0 < e["value"] < d["value"]
    assert e["h2d_bytes_per_step"] == bps * (2 << 20) and e["d2h_bytes_per_step"] == bps * (4 << 20)
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and 0 < c["value"] < d["value"] / 10
    assert d["clocks"]["sm_mhz"] > 0 and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}

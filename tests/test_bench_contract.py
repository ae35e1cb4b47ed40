"""bench.py's output contract, as far as it can be checked without a GPU: the reference arm prints exactly ONE JSON line on stdout with
the keys the driver reads, and the B200 arm fails loudly (no JSON, non-zero exit) when there is no device -- never a CPU fallback."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600, cwd=ROOT)


def test_reference_arm_prints_one_json_line():
    out = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--batch", "64")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "sqp_lm_iterations_per_second" and d["unit"] == "iters/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # same config keys as the B200 arm, and the sample really is the batch the label names
    assert d["config"]["workload"].startswith("van_der_pol_fd_n50_batch64") and d["config"]["instances_total"] == 64
    assert d["config"]["n_grid"] == 50 and d["config"]["n_params"] == 147
    assert "64 instances" in d["cpu_baseline"]["sample"] and d["wall_value"] > 0 and d["wall_value"] <= d["value"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_b200_arm_fails_loudly_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    out = _run("--steps", "1", "--warmup", "1")
    assert out.returncode != 0 and out.stdout.strip() == ""
    assert "no CUDA device" in out.stderr


def test_traffic_capture_carries_the_hash_of_the_kernel_sources():
    """roofline.traffic is only reported when profiles/traffic.json was captured from the tree bench.py runs from"""
    sys.path.insert(0, ROOT)
    import bench

    t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    for key, e in t.items():
        assert {"kernel", "dram_bytes_read", "dram_bytes_write", "csrc_sha256", "source"} <= set(e), key
        value, src = bench.measured_traffic(key)
        if e["csrc_sha256"] == bench.csrc_hash():
            assert value == e["dram_bytes_read"] + e["dram_bytes_write"]
        else:
            assert value is None and src.startswith("stale")
    assert bench.measured_traffic("no_such_workload")[0] is None


def test_algorithmic_flops_of_the_headline_workload():
    """SURVEY.md section 8d quotes ~35 kflop per LM iteration of Van der Pol N=50 (FD sweep ~25 k + J^T J + block Cholesky)"""
    sys.path.insert(0, ROOT)
    from control_box_rst_b200 import problems, roofline, solver

    ocp, _, _ = problems.config(1)
    f = roofline.flops_per_iteration(ocp, solver.dims_of(ocp))
    assert 25_000 <= f <= 45_000, f

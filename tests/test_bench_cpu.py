"""bench.py host-side contract checks that need no GPU: the reference arm prints one JSON line with the
required keys (run here on a small env count), and the B200 arm refuses to run without a CUDA device."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          env={**os.environ, **(env or {})}, timeout=600)


def test_reference_arm_json_line():
    out = run("--impl", "reference", "--gpus", "1", "--steps", "3", "--warmup", "3", "--envs", "2048")
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "env-steps/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("env steps/sec (9x9x5") and line["value"] > 0 and line["steps"] == 3
    from oracle import ref_tree
    # kind "reference" = the unmodified tree staged under oracle/_ref (or mounted); "port" only where it is absent
    assert line["cpu_baseline"]["kind"] == ("reference" if ref_tree.available() else "port") and line["cpu_baseline"]["cores"] >= 1
    assert set(line["config"]) == {"workload", "envs_per_gpu", "global_envs", "l2", "parallelism"}     # == the B200 arm's keys
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("cfg2")


def test_reference_arm_other_ranks_exit_quietly():
    out = run("--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "3", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_b200_arm_has_no_cpu_fallback():
    if torch.cuda.is_available():
        return
    out = run("--steps", "2", "--warmup", "3")
    assert out.returncode != 0 and "no CUDA device" in (out.stderr + out.stdout)


def test_reference_rollout_baseline_runs_the_unmodified_ppo_loop():
    """cfg3's cpu_baseline: the reference's own PPOAgent.learn rollout section (fps as ppo.py:126-129 computes it)."""
    from oracle import ref_tree
    if not ref_tree.available():
        import pytest
        pytest.skip("reference tree neither mounted nor staged (oracle/_ref)")
    sys.path.insert(0, ROOT)
    import bench
    wl = bench.Workload("cfg2", 1)
    res = bench.cpu_rollout_run(wl, envs=64, n_steps=2, budget_s=5.0)
    assert res["kind"] == "reference" and res["unit"] == "agent-steps/s" and res["value"] > 0 and res["cores"] >= 1


def test_workload_bytes_match_survey():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.Workload("cfg2", 1).alg_bytes == 814 and bench.Workload("cfg2", 1).packed_bytes == 85
    assert bench.Workload("cfg4", 4).alg_bytes == 1638 and bench.Workload("cfg4", 4).envs == 262144
    assert bench.Workload("cfg5", 8).alg_bytes == 3462 and bench.Workload("cfg5", 8).envs == 524288

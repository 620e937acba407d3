"""CPU checks of the reference-driven test plumbing: the staged tree is the unmodified reference, the
reference's own test file runs through tests/run_reference_tests.py, and the INTEGRATION.md (A) import
overlay resolves hot-path modules to the drop-in and everything else to the reference."""
import filecmp
import json
import os
import subprocess
import sys

import pytest

from oracle import ref_tree

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not ref_tree.available(), reason="reference tree neither mounted nor staged (oracle/_ref)")

REFERENCE_OUTCOMES = {      # what the UNMODIFIED reference does on its own tests (6 of 7 pass;
    "test_env_mechanics_win": "passed",            # test_env_illegal_move expects validators that are dead code)
    "test_env_illegal_move": "failed",
    "test_canonical_view": "passed",
    "test_agent_win_reward": "passed",
    "test_opponent_win_penalty": "passed",
    "test_autoreset_next_step": "passed",
    "test_opponent_starts_after_reset": "passed",
}


def run_reference_tests(mode, tmp_path, extra_env=None):
    out = tmp_path / f"{mode}.json"
    env = dict(os.environ, **(extra_env or {}))
    env.pop("PYTHONPATH", None)
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "run_reference_tests.py"), mode, str(out)],
                          cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert out.exists(), proc.stdout[-2000:] + proc.stderr[-2000:]
    return json.loads(out.read_text())


def test_staged_tree_is_the_unmodified_reference():
    if not (os.path.isdir(ref_tree.MOUNTED) and os.path.isdir(os.path.join(ref_tree.STAGED, "src"))):
        pytest.skip("needs both the mounted reference and the staged copy (build container after build())")
    cmp = filecmp.dircmp(os.path.join(ref_tree.MOUNTED, "src"), os.path.join(ref_tree.STAGED, "src"), ignore=["__pycache__"])
    stack, checked = [cmp], 0
    while stack:
        c = stack.pop()
        assert not c.diff_files and not c.left_only and not c.right_only, (c.left, c.diff_files, c.left_only, c.right_only)
        checked += len(c.same_files)
        stack.extend(c.subdirs.values())
    assert checked >= 30
    tracked = subprocess.run(["git", "ls-files", "oracle/_ref"], cwd=ROOT, capture_output=True, text=True).stdout
    assert tracked.strip() == "", "reference sources must stay out of this repository's history"


def test_reference_own_tests_on_the_reference(tmp_path):
    res = run_reference_tests("reference", tmp_path)
    assert res["env_class"].endswith("torch_vector_mnk_env.TorchVectorMnkEnv") and not res["native_lib_loaded"]
    assert {k: v["outcome"] for k, v in res["results"].items()} == REFERENCE_OUTCOMES


def test_overlay_resolves_hot_path_to_dropin_and_the_rest_to_the_reference():
    ppo, buf, sched, wrap, pool = ref_tree.load_dropin("alg.ppo", "alg.rollout_buffer", "alg.entropy_scheduler",
                                                        "selfplay.torch_self_play_wrapper", "selfplay.opponent_pool")
    ref_src = os.path.join(ref_tree.root(), "src")
    assert os.path.abspath(ppo.__file__).startswith(ref_src)                       # PPO: the reference's, unmodified
    assert os.path.abspath(sched.__file__).startswith(ref_src)
    assert os.path.abspath(buf.__file__).startswith(ref_tree.PKG)                  # buffer: packed bitboards
    assert ppo.RolloutBuffer is buf.RolloutBuffer and buf.RolloutBuffer.__module__ == "mnk_b200.rollout"
    assert wrap.TorchSelfPlayWrapper.__module__ == "mnk_b200.wrapper"
    assert os.path.abspath(pool.__file__).startswith(ref_tree.PKG)
    stock = ref_tree.load("selfplay.torch_self_play_wrapper")
    assert stock.TorchSelfPlayWrapper.__module__ == "selfplay.torch_self_play_wrapper"
    assert os.path.abspath(stock.__file__).startswith(ref_src)

"""The north-star sentence "src/train.py and src/validation.py can use it as a drop-in", tested with the
UNMODIFIED reference (staged under oracle/_ref/ by __graft_entry__.build()):

  * the reference's own test file runs against the drop-in modules (same outcomes as on the reference itself,
    and all seven green with the opt-in strict mode);
  * the reference's own ``PPOAgent.learn`` (src/alg/ppo.py:78-166) trains on the drop-in env + wrapper + packed
    buffer, and the reference's own ``validate_gpu`` (src/selfplay/validation.py:6-44) evaluates on it;
  * live lock-step: the stock reference env / wrapper on ``device="cuda"`` and the bitboard kernels fed the
    same actions at BASELINE cfg2 size -- boards, masks, rewards, dones, players, counters ``torch.equal``.
"""
import importlib.util
import os
from unittest import mock

import numpy as np
import pytest
import torch

from oracle import ref_tree
from test_reference_tree_cpu import REFERENCE_OUTCOMES, run_reference_tests

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_tree.available(), reason="reference tree neither mounted nor staged (oracle/_ref)")]
DEV = "cuda"


# ------------------------------------------------------------------------------------------------
# 1. the reference's own tests, file untouched, on the drop-in
# ------------------------------------------------------------------------------------------------
def test_reference_own_tests_on_the_dropin(tmp_path):
    res = run_reference_tests("dropin", tmp_path)
    assert res["cuda"] and res["native_lib_loaded"] and res["env_class"] == "mnk_b200.env.TorchVectorMnkEnv", res
    assert {k: v["outcome"] for k, v in res["results"].items()} == REFERENCE_OUTCOMES, res["results"]


def test_reference_own_tests_on_the_dropin_strict_mode_all_green(tmp_path):
    """test_env_illegal_move expects the validators the reference never calls (torch_vector_mnk_env.py:86-104);
    MNK_B200_STRICT=1 turns the drop-in's device-side legality flag into those ValueErrors."""
    res = run_reference_tests("dropin", tmp_path, {"MNK_B200_STRICT": "1"})
    assert res["native_lib_loaded"]
    assert {k: v["outcome"] for k, v in res["results"].items()} == {k: "passed" for k in REFERENCE_OUTCOMES}, res["results"]


# ------------------------------------------------------------------------------------------------
# 2. the reference's PPOAgent.learn and validate_gpu on the drop-in
# ------------------------------------------------------------------------------------------------
def _reference_validate_gpu():
    """src/selfplay/validation.py loaded BY FILE from the reference tree, with its imports
    (env.torch_vector_mnk_env, selfplay.torch_self_play_wrapper) resolving to the drop-in."""
    path = os.path.join(ref_tree.root(), "src", "selfplay", "validation.py")
    with ref_tree.imports(ref_tree.dropin_path()):
        spec = importlib.util.spec_from_file_location("reference_validation", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    assert mod.TorchVectorMnkEnv.__module__ == "mnk_b200.env" and mod.TorchSelfPlayWrapper.__module__ == "mnk_b200.wrapper"
    return mod.validate_gpu


def test_reference_validate_gpu_on_the_dropin_known_answer():
    """Uniformly random play on 3x3x3 has a known outcome distribution (255,168 games weighted by their
    probabilities: first player 58.49 %, second 28.81 %, draw 12.70 %); the reference's validate_gpu puts the
    agent first in half of the games."""
    from mnk_b200 import RandomPolicy
    validate_gpu = _reference_validate_gpu()
    n = 40000
    res = validate_gpu(RandomPolicy(9, seed=1), RandomPolicy(9, seed=2), (3, 3, 3), n_episodes=n, device=DEV)
    p_first, p_second, p_draw = 0.584921, 0.288095, 0.126984
    want_win, want_loss = 0.5 * (p_first + p_second), 0.5 * (p_first + p_second)
    tol = 4 * (0.25 / n) ** 0.5
    assert abs(res["validation/vs_benchmark/win_rate"] - want_win) < tol
    assert abs(res["validation/vs_benchmark/loss_rate"] - want_loss) < tol
    assert abs(res["validation/vs_benchmark/draw_rate"] - p_draw) < tol
    assert res["validation/vs_benchmark/games_played"] == n


def test_reference_ppo_agent_learns_on_the_dropin():
    """The reference's PPOAgent (rollout loop :78-133 + update :168-262), its network class and its
    HardwareConfig, all imported from the unmodified tree, driving the drop-in env / wrapper / RolloutBuffer."""
    ppo, hw, cfg, pol = ref_tree.load_dropin("alg.ppo", "utils.hardware", "alg.architectures.configs", "selfplay.policy")
    from mnk_b200 import TorchSelfPlayWrapper, TorchVectorMnkEnv
    from mnk_b200.rollout import RolloutBuffer
    assert pol.NNPolicy.__module__ == "mnk_b200.policy"
    validate_gpu = _reference_validate_gpu()
    torch.manual_seed(0)
    m = n = k = 3
    ne, n_steps = 512, 16
    net = cfg.ResNetSActorCritic((2, m, n), m * n)
    opt = torch.optim.AdamW(net.parameters(), lr=1e-3, eps=1e-5)
    agent = ppo.PPOAgent((2, m, n), m * n, net, hw_config=hw.HardwareConfig("cuda", torch.bfloat16, False, None),
                         n_steps=n_steps, optimizer=opt, batch_size=1024, num_envs=ne, ppo_epochs=3, entropy_coef=0.01)
    assert isinstance(agent.buffer, RolloutBuffer)                 # `from .rollout_buffer import RolloutBuffer` -> packed buffer
    env = TorchVectorMnkEnv(m, n, k, ne, device=DEV)
    wrapper = TorchSelfPlayWrapper(env, seed=5)
    wrapper.set_opponent(pol.RandomPolicy(m * n, seed=6))
    before = validate_gpu(pol.NNPolicy(agent.network), pol.RandomPolicy(9, seed=7), (m, n, k), n_episodes=4096, device=DEV)
    agent.network.train()
    metrics = []
    for _ in range(30):
        metrics.append(agent.learn(wrapper))
    for mt in metrics:
        for name in ("actor_loss", "critic_loss", "entropy_loss", "grad_norm", "approx_kl", "explained_variance", "fps"):
            assert np.isfinite(getattr(mt, name)), (name, mt)
        assert mt.fps > 0 and mt.mean_length > 0
    assert agent.buffer.ptr == 0                                    # learn() ends with buffer.reset()
    after = validate_gpu(pol.NNPolicy(agent.network), pol.RandomPolicy(9, seed=7), (m, n, k), n_episodes=4096, device=DEV)
    s0, s1 = before["validation/vs_benchmark/score_rate"], after["validation/vs_benchmark/score_rate"]
    print(f"reference PPOAgent on the drop-in: score vs random {s0:.3f} -> {s1:.3f}")
    assert s1 > s0 + 0.08 and s1 > 0.62, (s0, s1)
    assert metrics[-1].mean_reward > metrics[0].mean_reward


# ------------------------------------------------------------------------------------------------
# 3. live lock-step against the stock reference on the same GPU
# ------------------------------------------------------------------------------------------------
def test_live_reference_env_lockstep_cfg2_size():
    """BASELINE cfg2 (9x9x5 x 65,536 envs): identical action tensors into the stock reference env on
    device="cuda" and into the bitboard kernels for 200 steps, finished envs reset in both."""
    from mnk_b200 import TorchVectorMnkEnv
    ref_env_mod = ref_tree.load("env.torch_vector_mnk_env")
    m, n, k, ne = 9, 9, 5, 65536
    ref = ref_env_mod.TorchVectorMnkEnv(m, n, k, ne, device=DEV)
    env = TorchVectorMnkEnv(m, n, k, ne, device=DEV)
    o_ref, o = ref.reset(), env.reset()
    wins = draws = 0
    for t in range(200):
        assert torch.equal(o["observation"], o_ref["observation"]) and torch.equal(o["action_mask"], o_ref["action_mask"]), t
        a = env.random_legal_actions(seed=3, counter=t)
        o_ref, r_ref, d_ref = ref.step(a)
        o, r, d = env.step(a)
        assert torch.equal(r, r_ref) and torch.equal(d, d_ref), t
        assert r.dtype == r_ref.dtype and d.dtype == d_ref.dtype and o["action_mask"].dtype == o_ref["action_mask"].dtype
        wins += int(r.sum())
        draws += int((d & (r == 0)).sum())
        done = torch.nonzero(d).squeeze(1)
        if done.numel():
            o_ref, o = ref.reset(done), env.reset(done)
        if t % 50 == 49:
            assert torch.equal(env.boards, ref.boards)
            assert torch.equal(env.current_player, ref.current_player) and torch.equal(env.move_counts, ref.move_counts)
            env.release_mirrors()
    assert wins > 100000 and draws > 100          # the trace exercised both terminal kinds


class HashPolicy:
    """Deterministic row-wise opponent (same arithmetic as oracle/gen_golden.py::HashPolicy) usable by both wrappers."""

    def act(self, obs_dict):
        obs, mask = obs_dict["observation"], obs_dict["action_mask"]
        b, cells = mask.shape
        w = torch.arange(1, 2 * cells + 1, dtype=torch.int64, device=obs.device)
        score = (obs.reshape(b, -1).to(torch.int64) * w).sum(dim=1)
        cnt = mask.sum(dim=1)
        j = score % torch.clamp(cnt, min=1)
        rank = torch.cumsum(mask.to(torch.int64), dim=1) - 1
        picked = torch.argmax((mask & (rank == j[:, None])).to(torch.int64), dim=1)
        return torch.where(cnt == 0, score % cells, picked)


@pytest.mark.parametrize("m,n,k,ne,steps", [(9, 9, 5, 8192, 150), (3, 3, 3, 4096, 60)], ids=lambda v: str(v))
def test_live_reference_wrapper_lockstep(m, n, k, ne, steps):
    """TorchSelfPlayWrapper.reset / step (torch_self_play_wrapper.py:19-67) of the stock reference on cuda against
    the fused kernels: same agent actions, same deterministic opponent, sides injected on both sides (the
    reference's torch.randint patched; the drop-in's `next_sides`)."""
    from mnk_b200 import TorchSelfPlayWrapper, TorchVectorMnkEnv
    env_mod, wrap_mod = ref_tree.load("env.torch_vector_mnk_env", "selfplay.torch_self_play_wrapper")
    gen = torch.Generator(device="cpu").manual_seed(m * 100 + ne)
    side_table = torch.randint(0, 2, (steps + 1, ne), generator=gen).to(DEV)
    ref_wr = wrap_mod.TorchSelfPlayWrapper(env_mod.TorchVectorMnkEnv(m, n, k, ne, device=DEV))
    ref_wr.set_opponent(HashPolicy())
    wr = TorchSelfPlayWrapper(TorchVectorMnkEnv(m, n, k, ne, device=DEV), seed=1)
    wr.set_opponent(HashPolicy())
    row = {"sides": None}
    real_randint = torch.randint

    def fake_randint(low, high, size, **kw):
        assert (low, high) == (0, 2) and row["sides"] is not None and len(row["sides"]) == size[0]
        return row["sides"].clone()

    with mock.patch.object(torch, "randint", fake_randint):
        row["sides"] = side_table[0]
        o_ref, _ = ref_wr.reset()
    wr.next_sides = side_table[0]
    o, info = wr.reset()
    assert info == {}
    outcomes = torch.zeros(3, dtype=torch.long, device=DEV)
    for t in range(steps):
        assert torch.equal(o["observation"], o_ref["observation"]) and torch.equal(o["action_mask"], o_ref["action_mask"]), t
        assert torch.equal(wr.agent_side, ref_wr.agent_side) and torch.equal(wr.pending_resets, ref_wr.pending_resets), t
        # agent: a random legal cell of the canonical mask (identical on both sides by the assert above)
        actions = torch.multinomial(o["action_mask"].float(), 1, generator=None).squeeze(1)
        pending = ref_wr.pending_resets
        with mock.patch.object(torch, "randint", fake_randint):
            row["sides"] = side_table[t + 1][pending]
            o_ref, r_ref, term_ref, trunc_ref, _ = ref_wr.step(actions)
        wr.next_sides = side_table[t + 1]
        o, r, term, trunc, info = wr.step(actions)
        assert torch.equal(r, r_ref) and torch.equal(term, term_ref) and torch.equal(trunc, trunc_ref) and info == {}, t
        outcomes += torch.stack([(r > 0).sum(), (r < 0).sum(), (term & (r == 0)).sum()])
    assert torch.equal(wr.env.boards, ref_wr.env.boards) and torch.equal(wr.env.move_counts, ref_wr.env.move_counts)
    assert int(outcomes[0]) > 0 and int(outcomes[1]) > 0
    assert real_randint is torch.randint

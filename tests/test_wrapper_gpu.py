"""GPU parity tests of the fused self-play wrapper, the policies and the masked sampler."""
import numpy as np
import pytest
import torch

import golden_io as gio
from oracle import mnk_oracle as orc

from mnk_b200 import RandomPolicy, TorchSelfPlayWrapper, TorchVectorMnkEnv

pytestmark = pytest.mark.gpu
DEV = "cuda"


def make(m, n, k, ne, seed=0, **kw):
    from env.torch_vector_mnk_env import TorchVectorMnkEnv
    from selfplay.torch_self_play_wrapper import TorchSelfPlayWrapper
    env = TorchVectorMnkEnv(m, n, k, ne, device=DEV, **kw)
    return TorchSelfPlayWrapper(env, seed=seed)


def t(x):
    return torch.as_tensor(np.ascontiguousarray(x), device=DEV)


class ScriptedPolicy:
    """As in the reference's tests (src/tests/test_mnk_integration.py:11-24): takes obs_dict only."""

    def __init__(self, action_idx):
        self.action_idx = action_idx

    def act(self, obs_dict):
        b = obs_dict["action_mask"].shape[0]
        return torch.full((b,), self.action_idx, device=obs_dict["action_mask"].device, dtype=torch.long)


class HashPolicy:
    """torch twin of oracle/gen_golden.py::HashPolicy (deterministic, row-wise)."""

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


@pytest.mark.parametrize("path", gio.files("wrapper_trace_"), ids=gio.name)
def test_golden_wrapper_trace(path):
    g = gio.load(path)
    m, n, k, ne = (int(x) for x in g["geom"])
    sides = g["sides"]
    wr = make(m, n, k, ne)
    wr.set_opponent(HashPolicy())
    if int(g["options_reset"]):
        obs, info = wr.reset(options={"agent_side": t(sides[0])})
    else:
        wr.next_sides = t(sides[0])
        obs, info = wr.reset()
    assert info == {}
    assert np.array_equal(obs["observation"].cpu().numpy(), gio.unpack(g["obs0"], (2, m, n)).astype(np.float32))
    assert np.array_equal(obs["action_mask"].cpu().numpy(), gio.unpack(g["mask0"], (m * n,)))
    assert np.array_equal(wr.agent_side.cpu().numpy(), g["side0"])
    for step in range(len(g["actions"])):
        wr.next_sides = t(sides[step + 1])
        obs, r, term, trunc, info = wr.step(t(g["actions"][step]))
        assert obs["observation"].dtype == torch.float32 and obs["action_mask"].dtype == torch.bool
        assert r.dtype == torch.float32 and term.dtype == torch.bool and trunc.dtype == torch.bool and info == {}
        assert np.array_equal(obs["observation"].cpu().numpy(), gio.unpack(g["obs"][step], (2, m, n)).astype(np.float32)), step
        assert np.array_equal(obs["action_mask"].cpu().numpy(), gio.unpack(g["mask"][step], (m * n,))), step
        assert np.array_equal(r.cpu().numpy(), g["rewards"][step]), step
        assert np.array_equal(term.cpu().numpy(), g["terminated"][step]), step
        assert not trunc.any()
        assert np.array_equal(wr.agent_side.cpu().numpy(), g["agent_side"][step]), step
        assert np.array_equal(wr.pending_resets.cpu().numpy(), g["pending"][step]), step
    env = wr.env
    assert np.array_equal(env.boards.cpu().numpy().astype(bool), gio.unpack(g["boards"][-1], (2, m, n)))
    assert np.array_equal(env.current_player.cpu().numpy(), g["player"][-1])
    assert np.array_equal(env.move_counts.cpu().numpy(), g["count"][-1])


@pytest.mark.parametrize("m,n,k,ne", [(3, 3, 3, 500), (9, 9, 5, 2000), (13, 13, 5, 300), (6, 7, 4, 333), (19, 19, 5, 70)],
                         ids=lambda v: str(v))
def test_fused_random_opponent_vs_oracle(m, n, k, ne):
    """One-launch wrapper step with the on-device RandomPolicy opponent and Philox sides, against the
    numpy OracleWrapper fed with the same counter-based draws."""
    from selfplay.policy import RandomPolicy
    seed, off = 77, 1000
    wr = make(m, n, k, ne, seed=seed, env_offset=off)
    wr.set_opponent(RandomPolicy(m * n))
    episodes = np.zeros(ne, dtype=np.int64)
    clock = {"step": 0}

    def side_fn(idx):
        episodes[idx] += 1
        return orc.side_draw(seed, off + idx, episodes[idx])

    oenv = orc.OracleEnv(m, n, k, ne)
    owr = orc.OracleWrapper(oenv, side_fn=side_fn)
    owr.set_opponent(lambda od: orc.random_legal_actions(od["action_mask"], seed, clock["step"], env_offset=off,
                                                         env_ids=owr.last_active, stream=orc.STREAM_OPPONENT))
    clock["step"] = 1
    obs, _ = wr.reset()
    oobs, _ = owr.reset()
    rng = np.random.default_rng(0)
    seen = set()
    for step in range(2, min(2 * m * n, 120)):
        assert np.array_equal(obs["observation"].cpu().numpy(), oobs["observation"]), step
        assert np.array_equal(obs["action_mask"].cpu().numpy(), oobs["action_mask"]), step
        mask = oobs["action_mask"]
        cnt = mask.sum(1)
        j = (rng.random(ne) * cnt).astype(np.int64)
        a = np.argmax(mask & ((np.cumsum(mask, 1) - 1) == j[:, None]), axis=1).astype(np.int64)
        clock["step"] = step
        obs, r, term, _, _ = wr.step(t(a))
        oobs, r2, term2, _, _ = owr.step(a)
        assert np.array_equal(r.cpu().numpy(), r2) and np.array_equal(term.cpu().numpy(), term2), step
        assert np.array_equal(wr.agent_side.cpu().numpy(), owr.agent_side), step
        seen.update(np.unique(r2).tolist())
    assert seen >= {-1.0, 0.0, 1.0} or m * n > 100


def wrapper_factory(opponent_action_idx=0):
    wr = make(3, 3, 3, 1)
    wr.set_opponent(ScriptedPolicy(opponent_action_idx))
    return wr


def test_reference_wrapper_tests_port():
    """The five wrapper-level reference tests (src/tests/test_mnk_integration.py:89-207), same
    pokes into env.boards, same assertions."""
    # test_canonical_view
    wr = wrapper_factory()
    wr.reset(options={"agent_side": 0})
    wr.env.boards[0, 0, 0, 0] = 1.0
    assert wr.get_agent_obs()["observation"][0, 0, 0, 0] == 1.0
    wr.set_opponent(ScriptedPolicy(8))
    wr.reset(options={"agent_side": 1})
    wr.env.boards[0, 1, 0, 0] = 1.0
    obs = wr.get_agent_obs()
    assert obs["observation"][0, 0, 0, 0] == 1.0
    assert obs["observation"][0, 1, 2, 2] == 1.0
    # test_agent_win_reward
    wr = wrapper_factory()
    wr.reset(options={"agent_side": 0})
    wr.env.boards[0, 0, 0, 0] = 1
    wr.env.boards[0, 0, 0, 1] = 1
    obs, rewards, terms, trunc, _ = wr.step(torch.tensor([2], device=DEV))
    assert rewards[0].item() == 1.0 and terms[0].item() is True
    assert obs["observation"][0, 0].sum() == 3.0
    # test_opponent_win_penalty
    wr = wrapper_factory(opponent_action_idx=5)
    wr.reset(options={"agent_side": 0})
    wr.env.boards[0, 0, 0, 0] = 1
    wr.env.boards[0, 0, 0, 1] = 1
    wr.env.boards[0, 1, 1, 0] = 1
    wr.env.boards[0, 1, 1, 1] = 1
    obs, rewards, terms, truncs, _ = wr.step(torch.tensor([6], device=DEV))
    assert terms[0].item() is True and rewards[0].item() == -1.0
    assert obs["observation"][0, 1, 1, :].sum() == 3.0
    # test_autoreset_next_step
    wr = wrapper_factory()
    wr.reset(options={"agent_side": 0})
    wr.env.boards[0, 0, 0, 0] = 1
    wr.env.boards[0, 0, 0, 1] = 1
    obs, rewards, terms, _, _ = wr.step(torch.tensor([2], device=DEV))
    assert terms[0].item() is True and rewards[0].item() == 1.0 and obs["observation"][0, 0].sum() == 3.0
    wr.next_sides = torch.zeros(1, dtype=torch.long, device=DEV)     # the reference test implicitly needs black here
    obs_new, rewards_new, terms_new, _, _ = wr.step(torch.tensor([0], device=DEV))
    assert terms_new[0].item() is False and rewards_new[0].item() == 0.0
    assert obs_new["observation"][0, 0].sum() == 0.0
    # test_opponent_starts_after_reset
    wr = wrapper_factory(opponent_action_idx=4)
    obs, _ = wr.reset(options={"agent_side": 1})
    assert obs["observation"][0, 0].sum() == 0.0 and obs["observation"][0, 1, 1, 1] == 1.0


def test_masked_sampler_matches_categorical_arithmetic():
    from mnk_b200 import MaskedCategorical, masked_sample
    rng = np.random.default_rng(1)
    for acts in (9, 81, 169, 361, 500):
        rows = 257
        logits = rng.normal(size=(rows, acts)).astype(np.float32) * 3
        mask = rng.random((rows, acts)) < 0.4
        mask[5] = False                       # all-masked row -> uniform (resnet.py:91-92)
        mask[6] = True
        mask[7] = False
        mask[7, acts - 1] = True              # single legal action
        want = orc.masked_log_softmax(logits, mask)
        lg, mk = t(logits), t(mask)
        dist = MaskedCategorical(lg, mk)
        a, lp, ent = masked_sample(lg, mk, seed=3, counter=1, want_entropy=True)
        an = a.cpu().numpy()
        legal_or_dead = mask[np.arange(rows), an] | ~mask.any(1)
        assert legal_or_dead.all() and an[7] == acts - 1
        assert np.allclose(lp.cpu().numpy(), want[np.arange(rows), an], rtol=1e-5, atol=1e-5)
        p = np.exp(want)
        want_ent = -(np.where(p > 0, p * want, 0)).sum(1)
        assert np.allclose(ent.cpu().numpy(), want_ent, rtol=1e-4, atol=1e-5)
        assert np.allclose(dist.entropy().cpu().numpy(), want_ent, rtol=1e-4, atol=1e-5)
        fin = np.isfinite(want)
        got_logits = dist.logits.cpu().numpy()
        assert np.array_equal(np.isfinite(got_logits), fin)                    # -inf positions identical
        assert np.allclose(got_logits[fin], want[fin], rtol=1e-5, atol=1e-5)
        given = rng.integers(0, acts, size=rows)
        given = np.where(mask.any(1), np.argmax(mask, 1), given)               # a legal action per row
        assert np.allclose(dist.log_prob(t(given)).cpu().numpy(), want[np.arange(rows), given], rtol=1e-5, atol=1e-5)
        det = dist.mode().cpu().numpy()
        assert np.array_equal(det, np.argmax(want, axis=1))                    # first index on ties, like torch.argmax
        # same (seed, counter, row) => same draw; different counter => different draws
        a2 = masked_sample(lg, mk, seed=3, counter=1, want_log_prob=False)[0]
        a3 = masked_sample(lg, mk, seed=3, counter=2, want_log_prob=False)[0]
        assert torch.equal(a, a2) and not torch.equal(a, a3)
        off = masked_sample(lg[10:], mk[10:], seed=3, counter=1, row_offset=10, want_log_prob=False)[0]
        assert torch.equal(off, a[10:])                                        # sharding-invariant


def test_masked_sampler_distribution_chi_square():
    from mnk_b200 import masked_sample
    rng = np.random.default_rng(2)
    acts, rows = 81, 200000
    base = rng.normal(size=acts).astype(np.float32) * 1.5
    mask1 = rng.random(acts) < 0.5
    mask1[:3] = True
    logits = t(np.tile(base, (rows, 1)))
    mask = t(np.tile(mask1, (rows, 1)))
    a = masked_sample(logits, mask, seed=11, counter=5, want_log_prob=False)[0].cpu().numpy()
    p = np.exp(orc.masked_log_softmax(base[None], mask1[None]))[0]
    counts = np.bincount(a, minlength=acts)
    assert counts[~mask1].sum() == 0
    exp = p[mask1] * rows
    chi2 = ((counts[mask1] - exp) ** 2 / exp).sum()
    dof = mask1.sum() - 1
    assert chi2 < dof + 6 * np.sqrt(2 * dof), (chi2, dof)


def test_random_policy_uniform_and_first_legal():
    from selfplay.policy import RandomPolicy
    g = gio.load(gio.files("random_policy_first_legal")[0])
    mask = gio.unpack(g["mask"], (int(g["cells"]),))
    pol = RandomPolicy(81)
    det = pol.act({"action_mask": t(mask)}, deterministic=True)
    assert np.array_equal(det.cpu().numpy(), g["first_legal"])        # the reference's RandomPolicy(deterministic=True)
    one = np.zeros((120000, 9), dtype=bool)
    one[:, [0, 4, 5, 8]] = True
    a = pol.act({"action_mask": t(one)}).cpu().numpy()
    counts = np.bincount(a, minlength=9)
    assert counts[[1, 2, 3, 6, 7]].sum() == 0
    assert np.all(np.abs(counts[[0, 4, 5, 8]] / 120000 - 0.25) < 0.01)
    dead = np.zeros((90000, 9), dtype=bool)
    a = pol.act({"action_mask": t(dead)}).cpu().numpy()                # all-masked rows: uniform over every cell
    assert np.all(np.abs(np.bincount(a, minlength=9) / 90000 - 1 / 9) < 0.01)


def test_validate_gpu_random_vs_random_tictactoe():
    """Consumer of the path (src/selfplay/validation.py:6-44) on the drop-in: random vs random
    tic-tac-toe must reproduce the known first/second-player statistics (58.5% / 28.8% / 12.7%)."""
    from selfplay.policy import RandomPolicy
    from selfplay.validation import validate_gpu
    res = validate_gpu(RandomPolicy(9, seed=1), RandomPolicy(9, seed=2), (3, 3, 3), n_episodes=40000, device=DEV)
    win, loss, draw = (res[f"validation/vs_benchmark/{k}_rate"] for k in ("win", "loss", "draw"))
    assert abs(win + loss + draw - 1.0) < 1e-9
    assert abs(win - (0.585 + 0.288) / 2) < 0.015 and abs(loss - (0.585 + 0.288) / 2) < 0.015 and abs(draw - 0.127) < 0.01
    assert res["validation/vs_benchmark/games_played"] == 40000


def test_nnpolicy_with_stock_torch_network():
    """NNPolicy over a torch module with the reference's forward(obs, mask) -> (Categorical, value)."""
    from torch.distributions import Categorical
    from selfplay.policy import NNPolicy

    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(18, 9)

        def forward(self, obs, mask=None):
            logits = self.lin(obs.flatten(1))
            logits = torch.where(mask.bool(), logits, -torch.inf)
            dead = logits.max(dim=1, keepdim=True)[0] == -torch.inf
            logits = torch.where(dead, torch.zeros_like(logits), logits)
            return Categorical(logits=logits), logits.sum(1, keepdim=True)

    torch.manual_seed(0)
    pol = NNPolicy(Tiny().to(DEV))
    wr = make(3, 3, 3, 4096)
    wr.set_opponent(pol)
    obs, _ = wr.reset()
    for _ in range(12):
        a = pol.act(obs)
        assert bool(obs["action_mask"].gather(1, a[:, None]).all())         # always legal
        det = pol.act(obs, deterministic=True)
        assert bool(obs["action_mask"].gather(1, det[:, None]).all())
        obs, r, term, trunc, _ = wr.step(a)
    assert pol.act({"observation": obs["observation"][0], "action_mask": obs["action_mask"][0]}).shape == (1,)


def test_play_batch_games_tournament_primitive():
    """selfplay.match.play_batch_games (the fused counterpart of MatchRunner._play_batch_games,
    src/model_comparison/match_runner.py:125-218): random vs random tic-tac-toe reproduces the
    first-player / second-player statistics (58.5 % / 28.8 % / 12.7 %) for either colour."""
    from selfplay.match import play_batch_games
    from selfplay.policy import RandomPolicy
    n = 40000
    w, l, d = play_batch_games(RandomPolicy(9, seed=1), RandomPolicy(9, seed=2), (3, 3, 3), n, p1_is_black=True, device=DEV)
    assert w + l + d == n
    assert abs(w / n - 0.585) < 0.012 and abs(l / n - 0.288) < 0.012 and abs(d / n - 0.127) < 0.01
    w, l, d = play_batch_games(RandomPolicy(9, seed=3), RandomPolicy(9, seed=4), (3, 3, 3), n, p1_is_black=False, device=DEV)
    assert abs(w / n - 0.288) < 0.012 and abs(l / n - 0.585) < 0.012 and abs(d / n - 0.127) < 0.01
    assert play_batch_games(None, None, (3, 3, 3), 0, True, device=DEV) == (0, 0, 0)


def test_default_constructed_policies_draw_independently():
    """Two policies built without a seed must not hand the same noise to consecutive plies: on
    identical inputs their draws agree only as often as independent draws from the distribution would."""
    rows, cells = 20000, 81
    obs = {"observation": torch.zeros(rows, 2, 9, 9, device=DEV), "action_mask": torch.ones(rows, cells, dtype=torch.bool, device=DEV)}
    a, b = RandomPolicy(cells), RandomPolicy(cells)
    x, y = a.act(obs), b.act(obs)                       # both objects are at call count 1
    agree = (x == y).float().mean().item()
    assert abs(agree - 1.0 / cells) < 0.006, agree      # independent uniform draws agree with probability 1/81
    same = RandomPolicy(cells, seed=9).act(obs), RandomPolicy(cells, seed=9).act(obs)
    assert torch.equal(*same)                           # explicit equal seeds reproduce


def test_random_legal_counter_is_64_bit():
    """mnk_random_legal's counter is 64-bit end to end (r1: the kernel truncated it to 32 bits)."""
    env = TorchVectorMnkEnv(9, 9, 5, 4096, device=DEV)
    obs = env.reset()
    mask = obs["action_mask"].cpu().numpy()
    lo, hi = 12345, 12345 + (7 << 32)
    a_lo, a_hi = env.random_legal_actions(seed=2, counter=lo), env.random_legal_actions(seed=2, counter=hi)
    assert np.array_equal(a_lo.cpu().numpy(), orc.random_legal_actions(mask, 2, lo))
    assert np.array_equal(a_hi.cpu().numpy(), orc.random_legal_actions(mask, 2, hi))
    assert (a_lo != a_hi).float().mean().item() > 0.9


def test_agent_side_is_a_live_writable_mirror():
    """wrapper.agent_side (reference :13) can be written in place, as the reference's reset does with
    ``self.agent_side[:] = ...`` (:23), and tracks the sides the kernels assign."""
    env = TorchVectorMnkEnv(3, 3, 3, 64, device=DEV)
    wr = TorchSelfPlayWrapper(env, seed=4)
    wr.set_opponent(RandomPolicy(9, seed=1))
    wr.reset(options={"agent_side": 0})
    side = wr.agent_side
    assert side.dtype == torch.long and int(side.sum()) == 0
    side[:] = 1                                          # in-place write through the mirror
    obs = wr.get_agent_obs()
    raw = env.observe()
    assert torch.equal(obs["observation"], raw["observation"].flip(1))       # canonical view now swaps the planes
    wr.reset(options={"agent_side": torch.arange(64, device=DEV) % 2})
    assert torch.equal(side, torch.arange(64, device=DEV) % 2)               # the held tensor follows


def test_dropin_nnpolicy_runs_supported_models_on_the_native_forward():
    """selfplay.policy.NNPolicy -- the class the reference's train.py constructs for every opponent -- picks the tcgen05
    forward for the architectures that have one: same legal moves, logits-level agreement with the torch forward, weights
    re-imported after an in-place update, `native=False` keeps the torch path."""
    from selfplay.policy import NNPolicy
    from mnk_b200 import NativeConvNet, NativeResNet, NativeTransformer, TorchSelfPlayWrapper, TorchVectorMnkEnv, build_architecture
    torch.manual_seed(4)
    for arch, cls in (("resnet_b_s", NativeResNet), ("cnn_b_s", NativeConvNet), ("transformer_b_s", NativeTransformer)):
        model = build_architecture(arch, (2, 9, 9), 81).to(DEV)
        with torch.no_grad():                    # (0.01-gain initialisation: make the logits non-flat)
            (model.policy_head if hasattr(model, "policy_head") else model.actor)[7].weight.mul_(60.0)
        pol = NNPolicy(model, seed=5)
        assert isinstance(pol.net, cls) and pol.reads_bitboards and not model.training
        torch_pol = NNPolicy(model, seed=5, native=False)
        assert torch_pol.net is None and not torch_pol.reads_bitboards
        env = TorchVectorMnkEnv(9, 9, 5, 300, device=DEV, strict=True)
        wr = TorchSelfPlayWrapper(env, seed=2)
        wr.set_opponent(pol)
        obs, _ = wr.reset()
        for t in range(12):
            obs, r, term, _, _ = wr.step(torch.multinomial(obs["action_mask"].float(), 1).squeeze(1))
        pol.net.check_error()
        # deterministic actions from f32 observations: native vs torch forward agree wherever the top-2 gap is clear
        a_native, a_torch = pol.act(obs, deterministic=True), torch_pol.act(obs, deterministic=True)
        with torch.no_grad():
            d, _ = model(obs["observation"], obs["action_mask"])
        top2 = torch.topk(d.logits, 2, dim=1).values
        clear = (top2[:, 0] - top2[:, 1]) > 1e-2 * d.logits[torch.isfinite(d.logits)].abs().max()
        assert bool((a_native == a_torch)[clear].all()) and int(clear.sum()) > 0
        assert bool(obs["action_mask"].gather(1, a_native[:, None]).all())
        # an in-place update of the module is picked up
        before = pol.net.version
        with torch.no_grad():
            next(model.parameters()).mul_(1.01)
        pol.act(obs, deterministic=True)
        assert pol.net.version == before + 1

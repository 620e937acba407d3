"""GPU tests of the packed rollout buffer, GAE kernel, episode statistics and the collector."""
import numpy as np
import pytest
import torch

import golden_io as gio
from oracle import mnk_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda"


def t(x):
    return torch.as_tensor(np.ascontiguousarray(x), device=DEV)


def test_buffer_add_gae_and_gather_match_reference_buffer():
    """Same random transitions through the packed buffer as through the reference's RolloutBuffer
    (golden): stored observations / masks identical, advantages and returns bit-identical."""
    from alg.rollout_buffer import RolloutBuffer
    g = gio.load(gio.files("rollout_buffer_gae")[0])
    steps, ne, m, n = (int(x) for x in g["geom"])
    obs = gio.unpack(g["obs"], (2, m, n)).astype(np.float32).reshape(steps, ne, 2, m, n)
    masks = gio.unpack(g["masks"], (m * n,)).reshape(steps, ne, m * n)
    buf = RolloutBuffer(steps, ne, (2, m, n), m * n, device=DEV)
    for s in range(steps):
        buf.add(t(obs[s]), t(g["actions"][s]), t(g["rewards"][s]), t(g["values"][s]).view(-1, 1), t(g["log_probs"][s]),
                t(g["dones"][s]), t(masks[s]))
    with pytest.raises(IndexError):
        buf.add(t(obs[0]), t(g["actions"][0]), t(g["rewards"][0]), t(g["values"][0]), t(g["log_probs"][0]), t(g["dones"][0]))
    buf.compute_advantages_and_returns(t(g["last_values"]), round(float(g["gamma"]), 6), round(float(g["lam"]), 6))
    assert np.array_equal(buf.advantages.cpu().numpy(), g["advantages"])
    assert np.array_equal(buf.returns.cpu().numpy(), g["returns"])
    assert np.array_equal(buf.observations.cpu().numpy(), obs)
    fixed = masks.copy()
    fixed[~masks.any(-1), 0] = True            # the canonical mask carries the wrapper's all-masked fix
    assert np.array_equal(buf.action_masks.cpu().numpy(), fixed)
    seen = []
    for b_obs, b_act, b_lp, b_ret, b_adv, b_mask, b_val in buf.get_data_loader(64, normalize_advantages=False):
        assert b_obs.shape[1:] == (2, m, n) and b_mask.shape[1:] == (m * n,) and b_obs.dtype == torch.float32
        # identify each sample by (value, log_prob) and check its observation row
        flat_v, flat_lp = g["values"].reshape(-1), g["log_probs"].reshape(-1)
        for row in range(b_obs.shape[0]):
            idx = np.nonzero((flat_v == b_val[row].item()) & (flat_lp == b_lp[row].item()))[0]
            assert len(idx) == 1
            seen.append(int(idx[0]))
            assert np.array_equal(b_obs[row].cpu().numpy(), obs.reshape(-1, 2, m, n)[idx[0]])
            assert b_ret[row].item() == g["returns"].reshape(-1)[idx[0]] and b_act[row].item() == g["actions"].reshape(-1)[idx[0]]
    assert sorted(seen) == list(range(steps * ne))       # every sample exactly once
    buf.reset()
    assert buf.ptr == 0 and buf.rewards.abs().sum().item() == 0


@pytest.mark.parametrize("gamma,lam", [(0.99, 0.95), (0.997, 0.9), (0.9, 0.97), (0.993, 0.913), (1.0, 1.0), (0.95, 0.0)])
def test_gae_kernel_large_random_vs_oracle(gamma, lam):
    """The oracle's GAE is pinned to the live reference buffer for the same (gamma, lambda) pairs in
    tests/test_rollout_cpu.py, including pairs where (float)gamma * (float)lambda != (float)(gamma * lambda)."""
    from mnk_b200 import RolloutBuffer
    rng = np.random.default_rng(4)
    steps, ne = 128, 4099
    buf = RolloutBuffer(steps, ne, (2, 9, 9), 81, device=DEV)
    rewards = rng.choice([-1.0, 0.0, 1.0], size=(steps, ne)).astype(np.float32)
    values = rng.normal(size=(steps, ne)).astype(np.float32)
    dones = rng.random((steps, ne)) < 0.05
    last = rng.normal(size=ne).astype(np.float32)
    buf.rewards.copy_(t(rewards)), buf.values.copy_(t(values)), buf.dones.copy_(t(dones))
    buf.ptr = steps
    buf.compute_advantages_and_returns(t(last), gamma, lam)
    adv, ret = orc.gae(rewards, values, dones, last, gamma, lam)
    assert np.array_equal(buf.advantages.cpu().numpy(), adv) and np.array_equal(buf.returns.cpu().numpy(), ret)


class TinyNet(torch.nn.Module):
    """forward(obs, mask) -> (Categorical, value[B,1]) like the reference's networks (resnet.py:78-95)."""

    def __init__(self, cells):
        super().__init__()
        self.pi = torch.nn.Linear(2 * cells, cells)
        self.v = torch.nn.Linear(2 * cells, 1)

    def forward(self, obs, mask=None):
        x = obs.flatten(1)
        logits = torch.where(mask.bool(), self.pi(x), -torch.inf)
        dead = logits.max(dim=1, keepdim=True)[0] == -torch.inf
        logits = torch.where(dead, torch.zeros_like(logits), logits)
        return torch.distributions.Categorical(logits=logits), torch.tanh(self.v(x))


@pytest.mark.parametrize("m,n,k,ne,steps", [(3, 3, 3, 257, 40), (9, 9, 5, 512, 48)], ids=lambda v: str(v))
def test_collector_rollout_replays_on_oracle(m, n, k, ne, steps):
    """Collect a rollout with the fused wrapper (random opponent, Philox sides), then replay the
    recorded agent actions through the numpy OracleWrapper: every stored observation, mask,
    reward and done must match, as must the episode statistics."""
    from mnk_b200 import RandomPolicy, RolloutBuffer, RolloutCollector, TorchSelfPlayWrapper, TorchVectorMnkEnv
    torch.manual_seed(0)
    seed = 5
    env = TorchVectorMnkEnv(m, n, k, ne, device=DEV)
    wr = TorchSelfPlayWrapper(env, seed=seed)
    wr.set_opponent(RandomPolicy(m * n))
    net = TinyNet(m * n).to(DEV)
    buf = RolloutBuffer(steps, ne, (2, m, n), m * n, device=DEV, k=k)
    col = RolloutCollector(ne, device=DEV, seed=123)
    stats = col.collect(net, wr, buf)
    assert buf.ptr == steps and stats.agent_steps == steps * ne and stats.fps > 0
    got_obs = buf.observations.cpu().numpy()
    got_mask = buf.action_masks.cpu().numpy()
    actions = buf.actions.cpu().numpy()

    episodes = np.zeros(ne, dtype=np.int64)
    clock = {"step": 1}

    def side_fn(idx):
        episodes[idx] += 1
        return orc.side_draw(seed, idx, episodes[idx])

    owr = orc.OracleWrapper(orc.OracleEnv(m, n, k, ne), side_fn=side_fn)
    owr.set_opponent(lambda od: orc.random_legal_actions(od["action_mask"], seed, clock["step"], env_ids=owr.last_active,
                                                         stream=orc.STREAM_OPPONENT))
    oobs, _ = owr.reset()
    ep_r, ep_l = np.zeros(ne), np.zeros(ne)
    fin_r, fin_l = [], []
    for s in range(steps):
        assert np.array_equal(got_obs[s], oobs["observation"]), s
        assert np.array_equal(got_mask[s], oobs["action_mask"]), s
        legal = oobs["action_mask"][np.arange(ne), actions[s]]
        assert legal.all()                                            # sampled actions are always legal
        clock["step"] = s + 2
        oobs, r, term, _, _ = owr.step(actions[s])
        assert np.array_equal(buf.rewards[s].cpu().numpy(), r) and np.array_equal(buf.dones[s].cpu().numpy(), term), s
        ep_r += r
        ep_l += 1
        fin_r += ep_r[term].tolist()
        fin_l += ep_l[term].tolist()
        ep_r[term] = 0
        ep_l[term] = 0
    assert stats.episodes == len(fin_r) > 0
    assert abs(stats.mean_reward - np.mean(fin_r)) < 1e-6 and abs(stats.mean_length - np.mean(fin_l)) < 1e-6
    assert stats.wins + stats.losses + stats.draws == stats.episodes
    # log-probs stored are those of the sampled actions under the network's masked distribution
    with torch.no_grad():
        dist, values = net(t(got_obs[steps - 1]), t(got_mask[steps - 1]))
        assert torch.allclose(dist.log_prob(buf.actions[steps - 1]), buf.log_probs[steps - 1], atol=1e-5)
        assert torch.allclose(values.view(-1), buf.values[steps - 1], atol=1e-6)
    # the observation is carried across collect() calls (ppo.py:81-88,124): a second rollout continues the games
    buf.reset()
    col.collect(net, wr, buf)
    assert np.array_equal(buf.observations[0].cpu().numpy(), oobs["observation"])


@pytest.mark.parametrize("agent_arch,opp_arch", [("transformer_b_s", "resnet_b_l"), ("resnet_b_s", "transformer_b_s"), ("cnn_b_s", "cnn_b_s")])
def test_graph_rollout_with_wide_and_transformer_networks(agent_arch, opp_arch):
    """RolloutCollector.collect(graph=True) -- the whole rollout as one CUDA graph -- with the other native forwards on either
    side: captures (torch head tails included), replays with fresh draws, every stored action legal, flags clean."""
    from mnk_b200 import (NativeNNPolicy, RolloutBuffer, RolloutCollector, TorchSelfPlayWrapper, TorchVectorMnkEnv, build_architecture,
                          native_network, NativeResNet)
    torch.manual_seed(3)
    m, n, k, ne, steps = 9, 9, 5, 700, 6
    a_net = build_architecture(agent_arch, (2, m, n), m * n).to(DEV)
    agent = NativeResNet(a_net, device=DEV, bn_mode="train") if agent_arch == "resnet_b_s" else native_network(a_net.eval(), device=DEV)
    opp = NativeNNPolicy(build_architecture(opp_arch, (2, m, n), m * n).to(DEV), seed=4)
    env = TorchVectorMnkEnv(m, n, k, ne, device=DEV)
    wr = TorchSelfPlayWrapper(env, seed=9)
    wr.set_opponent(opp)
    buf = RolloutBuffer(steps, ne, (2, m, n), m * n, device=DEV, k=k)
    col = RolloutCollector(ne, device=DEV, seed=2)
    wr.reset(materialise=False)
    col._last_obs = {"observation": None, "action_mask": None}
    seen = []
    for rep in range(3):
        buf.reset()
        stats = col.collect(agent, wr, buf, graph=True)
        assert stats.agent_steps == steps * ne and buf.ptr == steps
        masks = buf.action_masks
        assert bool(masks.gather(2, buf.actions[:steps].unsqueeze(-1)).all())
        assert bool(torch.isfinite(buf.log_probs[:steps]).all()) and bool(torch.isfinite(buf.values[:steps]).all())
        seen.append(buf.actions[:steps].clone())
    assert not torch.equal(seen[0], seen[1]) and col._graph_replays == 3
    agent.check_error(), opp.net.check_error()


def test_rollout_on_13x13_with_the_train_mode_tower():
    """The board of the reference's second experiment (src/train_all_13.py): the agent's forward on mnk_resnet_tower_train
    (boards up to 13 rows), the frozen opponent on the tap kernel; eager and as a graph: legal actions, finite statistics,
    running statistics moving, flags clean."""
    import copy
    from mnk_b200 import NativeNNPolicy, NativeResNet, ResNetActorCritic, RolloutBuffer, RolloutCollector, TorchSelfPlayWrapper, TorchVectorMnkEnv
    torch.manual_seed(8)
    m, n, k, ne, steps = 13, 13, 5, 300, 5
    net = ResNetActorCritic((2, m, n), m * n).to(DEV)
    agent = NativeResNet(net, device=DEV, bn_mode="train")
    before = agent._params["running_mean"].clone()
    wr = TorchSelfPlayWrapper(TorchVectorMnkEnv(m, n, k, ne, device=DEV), seed=2)
    wr.set_opponent(NativeNNPolicy(copy.deepcopy(net), seed=3))
    buf = RolloutBuffer(steps, ne, (2, m, n), m * n, device=DEV, k=k)
    col = RolloutCollector(ne, device=DEV, seed=4)
    wr.reset(materialise=False)
    col._last_obs = {"observation": None, "action_mask": None}
    for graph in (False, True, True):
        buf.reset()
        stats = col.collect(agent, wr, buf, graph=graph)
        assert stats.agent_steps == steps * ne
        assert bool(buf.action_masks.gather(2, buf.actions[:steps].unsqueeze(-1)).all())
        assert bool(torch.isfinite(buf.log_probs[:steps]).all()) and bool(torch.isfinite(buf.values[:steps]).all())
    assert not torch.equal(before, agent._params["running_mean"])
    agent.check_error(), wr.opponent_policy.net.check_error()

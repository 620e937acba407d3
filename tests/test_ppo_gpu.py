"""PPO learner on the B200 rollout path: an end-to-end training smoke test."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_ppo_learns_tictactoe_against_random():
    """A few PPO iterations of the reference's default network on 3x3x3 against RandomPolicy: losses stay
    finite, the metrics have the reference's fields, and the win rate rises above random-vs-random."""
    from mnk_b200 import PPOAgent, RandomPolicy, ResNetActorCritic, TorchSelfPlayWrapper, TorchVectorMnkEnv
    from selfplay.policy import NNPolicy
    from selfplay.validation import validate_gpu
    torch.manual_seed(0)
    ne, steps = 512, 16
    env = TorchVectorMnkEnv(3, 3, 3, ne, device=DEV)
    wr = TorchSelfPlayWrapper(env, seed=1)
    wr.set_opponent(RandomPolicy(9, seed=3))
    net = ResNetActorCritic((2, 3, 3), 9).to(DEV)
    opt = torch.optim.AdamW(net.parameters(), lr=1e-3)
    agent = PPOAgent((2, 3, 3), 9, net, n_steps=steps, optimizer=opt, batch_size=1024, ppo_epochs=3, num_envs=ne,
                     device=DEV, k=3, entropy_coef=0.01)
    before = validate_gpu(NNPolicy(net), RandomPolicy(9, seed=5), (3, 3, 3), n_episodes=4096, device=DEV)
    net.train()
    history = []
    for _ in range(25):
        m = agent.learn(wr)
        history.append(m)
        for v in (m.actor_loss, m.critic_loss, m.entropy_loss, m.grad_norm, m.clip_fraction, m.explained_variance, m.approx_kl):
            assert np.isfinite(v)
        assert m.fps > 0 and m.rollout_time > 0 and m.learn_time > 0
    after = validate_gpu(NNPolicy(net), RandomPolicy(9, seed=5), (3, 3, 3), n_episodes=4096, device=DEV)
    s0, s1 = before["validation/vs_benchmark/score_rate"], after["validation/vs_benchmark/score_rate"]
    print(f"score vs random: {s0:.3f} -> {s1:.3f}; mean_reward {history[0].mean_reward:.3f} -> {history[-1].mean_reward:.3f}")
    assert s1 > s0 + 0.08 and s1 > 0.62
    assert history[-1].mean_reward > history[0].mean_reward

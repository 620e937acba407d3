"""Batched head-to-head matches on the fused wrapper.

The reference's tournament code (src/model_comparison/match_runner.py:125-218, `_play_batch_games`)
drives the raw env with per-side observation subsets, `step_subset` on the unfinished games and
four host synchronisations per ply.  The same result -- (wins, losses, draws) of player 1 over
`n_games` games with a fixed colour -- falls out of the self-play wrapper: player 1 is the "agent"
with a forced side, player 2 the opponent, a game's first terminal reward decides it, and an
m x n game ends within ceil(m*n/2)+1 agent steps, so the loop runs a fixed number of fused steps and
synchronises once.
"""
import torch

from env.torch_vector_mnk_env import TorchVectorMnkEnv
from selfplay.torch_self_play_wrapper import TorchSelfPlayWrapper


def play_batch_games(p1_policy, p2_policy, mnk_config, n_games: int, p1_is_black: bool, device="cuda"):
    """Returns (wins, losses, draws) from player 1's point of view."""
    if n_games == 0:
        return 0, 0, 0
    m, n, k = mnk_config
    env = TorchVectorMnkEnv(m, n, k, num_envs=n_games, device=device)
    wrapper = TorchSelfPlayWrapper(env)
    wrapper.set_opponent(p2_policy)
    side = torch.full((n_games,), 0 if p1_is_black else 1, dtype=torch.long, device=device)
    obs, _ = wrapper.reset(options={"agent_side": side})
    wrapper.next_sides = side                     # envs that auto-reset keep the same colours (their results are ignored)
    outcome = torch.zeros(n_games, device=device)
    decided = torch.zeros(n_games, dtype=torch.bool, device=device)
    for _ in range((m * n + 1) // 2 + 1):
        with torch.no_grad():
            actions = p1_policy.act(obs, deterministic=False)
        obs, rewards, terminated, _, _ = wrapper.step(actions)
        fresh = terminated & ~decided
        outcome = torch.where(fresh, rewards, outcome)
        decided |= terminated
    assert bool(decided.all()), "a game outlived m*n plies"
    wins = int((outcome == 1.0).sum())
    losses = int((outcome == -1.0).sum())
    return wins, losses, n_games - wins - losses

"""First-episode evaluation on the fused self-play wrapper.

Both consumers of the hot path that play "every env exactly one game" -- ``validate_gpu``
(reference: src/selfplay/validation.py:6-44) and the tournament's ``_play_batch_games`` (reference:
src/model_comparison/match_runner.py:125-218) -- reduce to the same device-side computation: step the
wrapper until every env has terminated once and keep the reward of each env's FIRST termination
(+1 / -1 / 0 from the agent's point of view).  The reference asks the host ``active_mask.any()``
before every step (and, in the match runner, three more times per ply); here the tallies stay on
the device, the loop polls a single "all decided" flag once per `poll_every` steps (a game cannot
outlive ceil(m*n/2)+1 agent steps, which also bounds the loop), and the result is read with one
3-element device->host copy.
"""
from __future__ import annotations

from typing import Tuple

import torch

from .env import TorchVectorMnkEnv
from .wrapper import TorchSelfPlayWrapper


def play_first_episodes(agent_policy, opponent_policy, mnk_config, sides: torch.Tensor, device="cuda",
                        poll_every: int = 8, deterministic: bool = False) -> Tuple[int, int, int]:
    """(wins, losses, draws) of `agent_policy` over len(sides) games, env e playing colour sides[e]
    (0 = black moves first, 1 = white)."""
    m, n, k = mnk_config
    games = int(sides.numel())
    if games == 0:
        return 0, 0, 0
    env = TorchVectorMnkEnv(m, n, k, num_envs=games, device=device)
    wrapper = TorchSelfPlayWrapper(env)
    wrapper.set_opponent(opponent_policy)
    sides = torch.as_tensor(sides, device=env._dev).long()
    obs, _ = wrapper.reset(options={"agent_side": sides})
    wrapper.next_sides = sides          # auto-reset games keep their colours; their results are never counted
    tally = torch.zeros(3, dtype=torch.int64, device=env._dev)           # wins, losses, draws
    open_games = torch.ones(games, dtype=torch.bool, device=env._dev)
    max_steps = (m * n + 1) // 2 + 1
    for step in range(1, max_steps + 1):
        with torch.no_grad():
            actions = agent_policy.act(obs, deterministic=deterministic)
        obs, rewards, terminated, _, _ = wrapper.step(actions)
        fresh = terminated & open_games
        tally += torch.stack([(fresh & (rewards > 0)).sum(), (fresh & (rewards < 0)).sum(), (fresh & (rewards == 0)).sum()])
        open_games &= ~terminated
        if step % poll_every == 0 and step < max_steps and not bool(open_games.any()):
            break
    wins, losses, draws = tally.tolist()
    if wins + losses + draws != games:
        raise RuntimeError(f"{games - wins - losses - draws} games outlived {max_steps} agent steps on a {m}x{n} board")
    return wins, losses, draws

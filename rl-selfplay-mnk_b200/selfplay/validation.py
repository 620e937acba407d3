"""``validate_gpu`` (reference: src/selfplay/validation.py:6-44) on the fused wrapper: same protocol
(first half of the episodes as black, second half as white; the first terminal reward of each env
counts) and the same result keys; the games themselves run in mnk_b200.evaluate."""
import torch

from mnk_b200.evaluate import play_first_episodes

_PREFIX = "validation/vs_benchmark/"


def validate_gpu(agent_policy, opponent_policy, mnk_config, n_episodes=1024, device="cuda"):
    sides = (torch.arange(n_episodes, device=device) >= n_episodes // 2).long()
    wins, losses, draws = play_first_episodes(agent_policy, opponent_policy, tuple(mnk_config), sides, device=device)
    rates = {"win_rate": wins, "loss_rate": losses, "draw_rate": draws, "score_rate": wins + 0.5 * draws}
    out = {_PREFIX + key: count / n_episodes for key, count in rates.items()}
    out[_PREFIX + "games_played"] = n_episodes
    return out

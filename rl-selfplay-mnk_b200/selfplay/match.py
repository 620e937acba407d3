"""Batched head-to-head matches (reference: src/model_comparison/match_runner.py:125-218,
``_play_batch_games``): (wins, losses, draws) of player 1 over `n_games` games with a fixed colour.
Player 1 is the wrapper's "agent" with a forced side, player 2 its opponent; see mnk_b200.evaluate."""
import torch

from mnk_b200.evaluate import play_first_episodes


def play_batch_games(p1_policy, p2_policy, mnk_config, n_games: int, p1_is_black: bool, device="cuda"):
    """Returns (wins, losses, draws) from player 1's point of view."""
    sides = torch.full((n_games,), 0 if p1_is_black else 1, dtype=torch.long, device=device)
    return play_first_episodes(p1_policy, p2_policy, tuple(mnk_config), sides, device=device)

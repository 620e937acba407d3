"""CPU BASELINE PORT: the reference env step restated op-for-op in PyTorch (CPU).

TEST / MEASUREMENT INFRASTRUCTURE ONLY -- never imported by the product.  The reference
is a Python program whose arithmetic lives in torch library calls; it cannot travel to
the GPU box, so this file restates the same *sequence of torch ops* (advanced-index
scatter, three valid F.conv2d cross-correlations with the same weights, threshold,
any-reduce, full-board clone in observe) so that bench.py's `cpu_baseline` /
`--impl reference` legs time what the reference would cost on the box's host cores:

  PortState / port_reset / port_observe / port_step_subset
      <- src/env/torch_vector_mnk_env.py:8-24, 34-44, 46-53, 60-84, 106-119
  port_uniform_legal
      <- src/selfplay/policy.py:17-29 (RandomPolicy.act: multinomial over mask.float())

Parity status: PINNED -- tests/test_oracle_golden.py::test_torch_port_* replays the golden
traces recorded from the unmodified reference through these functions.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.nn.functional as F


@dataclass
class PortState:
    m: int
    n: int
    k: int
    num_envs: int
    planes: torch.Tensor      # f32[N,2,m,n]   (reference: boards, :17)
    to_move: torch.Tensor     # i64[N]         (reference: current_player, :18)
    plies: torch.Tensor       # i64[N]         (reference: move_counts, :19)
    everyone: torch.Tensor    # i64[N] arange  (reference: env_indices, :22)
    w_row: torch.Tensor       # ones[1,1,1,k]  (:27)
    w_col: torch.Tensor       # ones[1,1,k,1]  (:28)
    w_diag: torch.Tensor      # [eye ; fliplr(eye)] [2,1,k,k]  (:30-32)


def port_make(m: int, n: int, k: int, num_envs: int, device="cpu") -> PortState:
    """`device="cuda"` runs the same torch op sequence on the GPU (the reference's own `device`
    argument, :8-16) -- bench.py's informative "stock PyTorch on the same B200" figure."""
    assert m >= k and n >= k
    eye = torch.eye(k, device=device)
    return PortState(
        m, n, k, num_envs,
        planes=torch.zeros((num_envs, 2, m, n), dtype=torch.float32, device=device),
        to_move=torch.zeros(num_envs, dtype=torch.long, device=device),
        plies=torch.zeros(num_envs, dtype=torch.long, device=device),
        everyone=torch.arange(num_envs, device=device),
        w_row=torch.ones((1, 1, 1, k), device=device),
        w_col=torch.ones((1, 1, k, 1), device=device),
        w_diag=torch.stack([eye, torch.fliplr(eye)]).reshape(2, 1, k, k),
    )


def port_observe(s: PortState):
    taken = (s.planes != 0.0).any(dim=1)
    return {"observation": s.planes.clone(), "action_mask": (~taken).flatten(1)}


def port_reset(s: PortState, which: Optional[torch.Tensor] = None):
    if which is None:
        s.planes.zero_()
        s.to_move.zero_()
        s.plies.zero_()
    else:
        s.planes[which] = 0
        s.to_move[which] = 0
        s.plies[which] = 0
    return port_observe(s)


def _line_found(s: PortState, which: torch.Tensor, movers: torch.Tensor) -> torch.Tensor:
    mine = s.planes[which, movers].unsqueeze(1)
    cut = s.k - 0.1
    b = which.shape[0]
    hits = [(F.conv2d(mine, w) > cut).view(b, -1).any(dim=1) for w in (s.w_row, s.w_col, s.w_diag)]
    return hits[0] | hits[1] | hits[2]


def port_step_subset(s: PortState, actions: torch.Tensor, which: torch.Tensor) -> Tuple[dict, torch.Tensor, torch.Tensor]:
    r = actions.div(s.n, rounding_mode="floor")
    c = actions % s.n
    movers = s.to_move[which]
    s.planes[which, movers, r, c] = 1.0
    s.plies[which] += 1
    won = _line_found(s, which, movers)
    drawn = (s.plies[which] >= s.m * s.n) & (~won)
    rewards = torch.zeros(s.num_envs, device=s.planes.device)
    if won.any():
        rewards[which[won]] = 1.0
    dones = torch.zeros(s.num_envs, dtype=torch.bool, device=s.planes.device)
    dones[which] = won | drawn
    s.to_move[which] ^= 1
    return port_observe(s), rewards, dones


def port_step(s: PortState, actions: torch.Tensor):
    return port_step_subset(s, actions, s.everyone)


def port_uniform_legal(mask: torch.Tensor) -> torch.Tensor:
    p = mask.float()
    empty = p.sum(dim=1, keepdim=True) == 0
    if empty.any():
        p = p + empty.float() * 1e-8
    return torch.multinomial(p, num_samples=1).squeeze(1)

"""Policies (reference: src/selfplay/policy.py:7-54).

``Policy.act(obs_dict, deterministic=False) -> int64[B]`` keeps the reference's contract.
``RandomPolicy`` draws a uniformly random legal cell with the warp-per-row sampler instead of
``torch.multinomial(mask.float())``; when it is the wrapper's opponent the wrapper bypasses
``act`` altogether and fuses the draw into its step kernel (mnk_selfplay_step_random).
``NNPolicy`` runs the model's forward (any torch module with the reference's
``forward(obs, mask) -> (dist, value)`` signature) and samples with the same kernel.
"""
from __future__ import annotations

import itertools
from abc import ABC, abstractmethod
from typing import Dict, Optional

import torch
import torch.nn as nn

from .sampling import MaskedCategorical, fresh_seed, masked_sample


class Policy(ABC):
    @abstractmethod
    def act(self, obs: Dict[str, torch.Tensor], deterministic: bool = False) -> torch.Tensor:
        pass


class RandomPolicy(Policy):
    """Uniform over legal cells; rows with no legal cell draw uniformly from all cells (the reference adds
    1e-8 to every entry of such rows, policy.py:21-24); deterministic => first legal cell (:26-27)."""

    def __init__(self, action_dim: int, seed: Optional[int] = None):
        self.action_dim = action_dim
        self.seed = fresh_seed() if seed is None else seed
        self._calls = itertools.count(1)

    def act(self, obs: Dict[str, torch.Tensor], deterministic: bool = False) -> torch.Tensor:
        mask = obs["action_mask"]
        if mask.dim() == 1:
            mask = mask.unsqueeze(0)
        zeros = torch.zeros(mask.shape, dtype=torch.float32, device=mask.device)
        return masked_sample(zeros, mask, seed=self.seed, counter=next(self._calls), deterministic=deterministic,
                             want_log_prob=False)[0]


class NNPolicy(Policy):
    def __init__(self, model: nn.Module, seed: Optional[int] = None):
        self.model = model
        self.model.eval()                       # policy.py:33-35
        self.seed = fresh_seed() if seed is None else seed
        self._calls = itertools.count(1)

    def act(self, obs: Dict[str, torch.Tensor], deterministic: bool = False) -> torch.Tensor:
        observation = obs["observation"]
        action_mask = obs["action_mask"]
        if observation.dim() == 3:              # policy.py:41-44
            observation = observation.unsqueeze(0)
        if action_mask.dim() == 1:
            action_mask = action_mask.unsqueeze(0)
        with torch.no_grad():
            dist, _ = self.model(observation, action_mask)
            if isinstance(dist, MaskedCategorical):
                return dist.mode() if deterministic else dist.sample()
            # a stock torch Categorical (the reference's networks): its logits are already masked and
            # normalised (-inf on illegal cells), so they can be sampled directly
            logits = dist.logits
            return masked_sample(logits, None, seed=self.seed, counter=next(self._calls), deterministic=deterministic,
                                 want_log_prob=False)[0]

"""Policies (reference: src/selfplay/policy.py:7-54).

``Policy.act(obs_dict, deterministic=False) -> int64[B]`` keeps the reference's contract.
``RandomPolicy`` draws a uniformly random legal cell with the warp-per-row sampler instead of
``torch.multinomial(mask.float())``; when it is the wrapper's opponent the wrapper bypasses
``act`` altogether and fuses the draw into its step kernel (mnk_selfplay_step_random).
``NNPolicy`` runs the model's forward (any torch module with the reference's
``forward(obs, mask) -> (dist, value)`` signature) and samples with the same kernel.  When the model is one of the
architectures with a tcgen05 forward (``mnk_b200.convnet.native_network``: resnet_b_s, resnet_b_l, cnn_b_s, cnn_b_l,
transformer_b_s, transformer_b_l) and lives on a CUDA device, the forward runs on that kernel instead -- so the reference's
unmodified ``NNPolicy(deepcopy(agent.network))`` opponents and validation policies (src/train.py:96-154) land on tensor
cores through the drop-in module path; ``native=False`` or ``MNK_B200_NATIVE_POLICY=0`` keeps the torch forward.
"""
from __future__ import annotations

import itertools
import os
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
    def __init__(self, model: nn.Module, seed: Optional[int] = None, native: Optional[bool] = None):
        self.model = model
        self.model.eval()                       # policy.py:33-35
        self.seed = fresh_seed() if seed is None else seed
        self._calls = itertools.count(1)
        self.counter_base: Optional[torch.Tensor] = None     # see TorchSelfPlayWrapper.counter_base
        self.net = None                         # the tcgen05 forward of `model`, where one exists
        if native is None:
            native = os.environ.get("MNK_B200_NATIVE_POLICY", "1") != "0"
        first = next(iter(model.parameters()), None)
        if native and first is not None and first.is_cuda:
            from .convnet import native_network
            try:
                self.net = native_network(model, device=first.device)
                self._versions = self._param_versions()
            except (ValueError, AttributeError, RuntimeError):
                self.net = None

    def _param_versions(self):
        return tuple(p._version for p in self.model.parameters()) + tuple(b._version for b in self.model.buffers())

    def _fresh_net(self):
        """The native forward holds a copy of the weights: re-import them if the module was updated in place since."""
        now = self._param_versions()
        if now != self._versions:
            self.net.refresh(self.model)
            self._versions = now
        return self.net

    @property
    def reads_bitboards(self) -> bool:
        """True when the wrapper may call act_from_env (the forward reads the env's packed state: no f32 opponent view)."""
        return self.net is not None

    def act_from_env(self, env, counter: int, deterministic: bool = False) -> torch.Tensor:
        """Action for the side to move of every env, straight from the bitboards (the wrapper's dense opponent call)."""
        swap = (env._meta & 1).to(torch.uint8)
        logits, _ = self._fresh_net().forward_env(env, swap, want_value=False)
        self._native_calls = getattr(self, "_native_calls", 0) + 1
        if self._native_calls % 512 == 0 and not torch.cuda.is_current_stream_capturing():
            self.net.check_error()           # one host read every 512 calls: a barrier timeout must not pass silently
        return masked_sample(logits, env.legal_mask(), seed=self.seed, counter=counter, row_offset=env.env_offset,
                             deterministic=deterministic, want_log_prob=False, counter_base=self.counter_base)[0]

    def act(self, obs: Dict[str, torch.Tensor], deterministic: bool = False) -> torch.Tensor:
        observation = obs["observation"]
        action_mask = obs["action_mask"]
        if observation.dim() == 3:              # policy.py:41-44
            observation = observation.unsqueeze(0)
        if action_mask.dim() == 1:
            action_mask = action_mask.unsqueeze(0)
        with torch.no_grad():
            if self.net is not None and observation.is_cuda:
                dist, _ = self._fresh_net().forward(observation, action_mask, want_value=False)
                return masked_sample(dist._raw, dist._mask, seed=self.seed, counter=next(self._calls),
                                     deterministic=deterministic, want_log_prob=False, counter_base=self.counter_base)[0]
            dist, _ = self.model(observation, action_mask)
            if isinstance(dist, MaskedCategorical):
                return dist.mode() if deterministic else dist.sample()
            # a stock torch Categorical (the reference's networks): its logits are already masked and
            # normalised (-inf on illegal cells), so they can be sampled directly
            logits = dist.logits
            return masked_sample(logits, None, seed=self.seed, counter=next(self._calls), deterministic=deterministic,
                                 want_log_prob=False)[0]

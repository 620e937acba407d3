"""The residual policy/value network of the reference ("resnet_b_s": 32 channels, 4 blocks, head
width 128; src/alg/architectures/resnet.py:8-95, configs.py:28-35) as a stock-PyTorch module with the
reference's parameter names, so that its ``state_dict`` files load unchanged
(``conv_in.0.weight``, ``res_blocks.N.conv1.weight``, ``policy_head.4.weight`` ...).

This module is the *learner-side* network (train-mode BatchNorm, autograd) and the fp32 yardstick
for the tcgen05 forward in ``mnk_b200.resnet``; its forward returns a ``MaskedCategorical`` whose
sample / log_prob / entropy run on the warp-per-row sampler.
"""
from __future__ import annotations

import torch.nn as nn
import torch.nn.functional as F

from .sampling import MaskedCategorical


class _ResBlock(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.conv1 = nn.Conv2d(channels, channels, kernel_size=3, padding=1)
        self.bn1 = nn.BatchNorm2d(channels)
        self.conv2 = nn.Conv2d(channels, channels, kernel_size=3, padding=1)
        self.bn2 = nn.BatchNorm2d(channels)

    def forward(self, x):
        y = F.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        return F.relu(y + x)


def _head(in_features: int, hidden: int, out_features: int, conv_out: int, channels: int, final=None):
    layers = [nn.Conv2d(channels, conv_out, kernel_size=1), nn.Flatten(), nn.LayerNorm(in_features), nn.ReLU(),
              nn.Linear(in_features, hidden), nn.LayerNorm(hidden), nn.ReLU(), nn.Linear(hidden, out_features)]
    if final is not None:
        layers.append(final)
    return nn.Sequential(*layers)


class ResNetActorCritic(nn.Module):
    def __init__(self, obs_shape, action_dim, channels: int = 32, num_blocks: int = 4, head_hidden_dim: int = 128):
        super().__init__()
        self.obs_shape = tuple(int(x) for x in obs_shape)
        self.action_dim = int(action_dim)
        self.channels, self.num_blocks = channels, num_blocks
        _, m, n = self.obs_shape
        self.conv_in = nn.Sequential(nn.Conv2d(self.obs_shape[0], channels, kernel_size=3, padding=1),
                                     nn.BatchNorm2d(channels), nn.ReLU())
        self.res_blocks = nn.Sequential(*[_ResBlock(channels) for _ in range(num_blocks)])
        self.policy_head = _head(2 * m * n, head_hidden_dim, action_dim, 2, channels)
        self.value_head = _head(m * n, head_hidden_dim, 1, 1, channels, final=nn.Tanh())
        self._init_weights()
        self._architecture_name = "resnet_b_s"
        self._architecture_params = {"obs_shape": list(self.obs_shape), "action_dim": self.action_dim}

    def _init_weights(self):
        """Orthogonal(relu gain) for conv / linear, unit norms, 0.01 / 1.0 gains on the last actor /
        critic layers -- the scheme of the reference's src/alg/weight_init.py:16-67."""
        gain = nn.init.calculate_gain("relu")
        for mod in self.modules():
            if isinstance(mod, (nn.Conv2d, nn.Linear)):
                nn.init.orthogonal_(mod.weight, gain=gain)
                nn.init.zeros_(mod.bias)
            elif isinstance(mod, (nn.BatchNorm2d, nn.LayerNorm)):
                nn.init.ones_(mod.weight)
                nn.init.zeros_(mod.bias)
        nn.init.orthogonal_(self.policy_head[7].weight, gain=0.01)
        nn.init.orthogonal_(self.value_head[7].weight, gain=1.0)

    def forward_body(self, x):
        return self.res_blocks(self.conv_in(x))

    def forward(self, obs, action_mask=None):
        features = self.forward_body(obs)
        logits = self.policy_head(features)
        value = self.value_head(features)
        if action_mask is not None and action_mask.dim() == 1 and logits.dim() == 2:
            action_mask = action_mask.unsqueeze(0)
        return MaskedCategorical(logits, action_mask), value

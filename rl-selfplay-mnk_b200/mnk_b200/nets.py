"""The convolutional policy/value networks of the reference as stock-PyTorch modules with the reference's
parameter names, so that its ``state_dict`` files load unchanged:

* ``ResNetActorCritic`` -- src/alg/architectures/resnet.py:8-95 (``conv_in.0.weight``, ``res_blocks.N.conv1.weight``,
  ``policy_head.4.weight`` ...): "resnet_b_s" (32 channels, 4 blocks, head width 128; configs.py:28-35, the default
  network) and "resnet_b_l" (80 channels, 5 blocks, head width 256; configs.py:38-45);
* ``CnnActorCritic`` -- src/alg/architectures/cnn.py:7-80 (``shared_body.N``, ``actor``, ``critic``): "cnn_b_s"
  ([56] * 4, head width 128) and "cnn_b_l" ([96] * 8, head width 256; configs.py:49-65).

``build_architecture(name, obs_shape, action_dim)`` constructs them by the reference's registry names.  These modules
are the *learner-side* networks (train-mode BatchNorm, autograd) and the fp32 yardsticks for the tcgen05 forwards in
``mnk_b200.resnet`` / ``mnk_b200.convnet``; their forward returns a ``MaskedCategorical`` whose sample / log_prob /
entropy run on the warp-per-row sampler.
"""
from __future__ import annotations

import torch
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


def _init_actor_critic(module: nn.Module, actor_last: nn.Linear, critic_last: nn.Linear):
    gain = nn.init.calculate_gain("relu")
    for mod in module.modules():
        if isinstance(mod, (nn.Conv2d, nn.Linear)):
            nn.init.orthogonal_(mod.weight, gain=gain)
            nn.init.zeros_(mod.bias)
        elif isinstance(mod, (nn.BatchNorm2d, nn.LayerNorm)):
            nn.init.ones_(mod.weight)
            nn.init.zeros_(mod.bias)
    nn.init.orthogonal_(actor_last.weight, gain=0.01)
    nn.init.orthogonal_(critic_last.weight, gain=1.0)


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
        self._architecture_name = {(32, 4, 128): "resnet_b_s", (80, 5, 256): "resnet_b_l", (64, 4, 256): "resnet_s",
                                   (128, 8, 256): "resnet_l"}.get(
            (channels, num_blocks, head_hidden_dim), f"resnet_{channels}x{num_blocks}_{head_hidden_dim}")
        self._architecture_params = {"obs_shape": list(self.obs_shape), "action_dim": self.action_dim}

    def _init_weights(self):
        """Orthogonal(relu gain) for conv / linear, unit norms, 0.01 / 1.0 gains on the last actor /
        critic layers -- the scheme of the reference's src/alg/weight_init.py:16-67."""
        _init_actor_critic(self, self.policy_head[7], self.value_head[7])

    def forward_body(self, x):
        return self.res_blocks(self.conv_in(x))

    def forward(self, obs, action_mask=None):
        features = self.forward_body(obs)
        logits = self.policy_head(features)
        value = self.value_head(features)
        if action_mask is not None and action_mask.dim() == 1 and logits.dim() == 2:
            action_mask = action_mask.unsqueeze(0)
        return MaskedCategorical(logits, action_mask), value


class CnnActorCritic(nn.Module):
    """BaseCnnActorCritic (src/alg/architectures/cnn.py:7-80): a stack of conv3x3 + BatchNorm + ReLU, then the same two
    heads as the residual network under the names ``actor`` / ``critic``."""

    def __init__(self, obs_shape, action_dim, channels=(56, 56, 56, 56), head_hidden_dim: int = 128):
        super().__init__()
        self.obs_shape = tuple(int(x) for x in obs_shape)
        self.action_dim = int(action_dim)
        self.channels = tuple(int(c) for c in channels)
        _, m, n = self.obs_shape
        layers, c_in = [], self.obs_shape[0]
        for c_out in self.channels:
            layers += [nn.Conv2d(c_in, c_out, kernel_size=3, padding=1), nn.BatchNorm2d(c_out), nn.ReLU()]
            c_in = c_out
        self.shared_body = nn.Sequential(*layers)
        self.actor = _head(2 * m * n, head_hidden_dim, action_dim, 2, c_in)
        self.critic = _head(m * n, head_hidden_dim, 1, 1, c_in, final=nn.Tanh())
        _init_actor_critic(self, self.actor[7], self.critic[7])
        self._architecture_name = {((56,) * 4, 128): "cnn_b_s", ((96,) * 8, 256): "cnn_b_l", ((64,) * 4, 256): "cnn_s",
                                   ((192,) * 6, 256): "cnn_l"}.get(
            (self.channels, head_hidden_dim), f"cnn_{'_'.join(map(str, self.channels))}_{head_hidden_dim}")
        self._architecture_params = {"obs_shape": list(self.obs_shape), "action_dim": self.action_dim}

    def forward_body(self, x):
        return self.shared_body(x)

    def forward(self, obs, action_mask=None):
        features = self.shared_body(obs)
        logits = self.actor(features)
        value = self.critic(features)
        if action_mask is not None and action_mask.dim() == 1 and logits.dim() == 2:
            action_mask = action_mask.unsqueeze(0)
        return MaskedCategorical(logits, action_mask), value


class TransformerActorCritic(nn.Module):
    """BaseTransformerActorCritic (src/alg/architectures/transformer.py:7-92): a 1x1 cell embedding + learned positions,
    pre-norm nn.TransformerEncoder layers (ReLU, feed-forward 4 x embed_dim, no dropout), Conv1d heads.  Parameter names
    are the reference's (``cell_embed``, ``pos_embed``, ``transformer.layers.N.*``, ``policy_head``, ``value_head``)."""

    def __init__(self, obs_shape, action_dim, embed_dim: int = 56, num_layers: int = 2, num_heads: int = 4,
                 head_hidden_dim: int = 128):
        super().__init__()
        self.obs_shape = tuple(int(x) for x in obs_shape)
        self.action_dim = int(action_dim)
        c, h, w = self.obs_shape
        tokens = h * w
        self.embed_dim, self.num_layers, self.num_heads = embed_dim, num_layers, num_heads
        self.cell_embed = nn.Conv2d(c, embed_dim, kernel_size=1, stride=1)
        self.pos_embed = nn.Parameter(torch.zeros(1, tokens, embed_dim))
        layer = nn.TransformerEncoderLayer(d_model=embed_dim, nhead=num_heads, dim_feedforward=embed_dim * 4,
                                           batch_first=True, norm_first=True, dropout=0.0)
        self.transformer = nn.TransformerEncoder(layer, num_layers=num_layers)

        def head(conv_out, final=None):
            layers = [nn.Conv1d(embed_dim, conv_out, kernel_size=1), nn.Flatten(), nn.LayerNorm(conv_out * tokens), nn.ReLU(),
                      nn.Linear(conv_out * tokens, head_hidden_dim), nn.LayerNorm(head_hidden_dim), nn.ReLU(),
                      nn.Linear(head_hidden_dim, action_dim if final is None else 1)]
            return nn.Sequential(*(layers + ([final] if final is not None else [])))

        self.policy_head = head(2)
        self.value_head = head(1, nn.Tanh())
        nn.init.normal_(self.pos_embed, std=0.02)            # transformer.py:53-55; the encoder keeps torch's defaults
        nn.init.normal_(self.cell_embed.weight, mean=0.0, std=0.02)
        nn.init.constant_(self.cell_embed.bias, 0.0)
        for hd, last_gain in ((self.policy_head, 0.01), (self.value_head, 1.0)):      # weight_init.py:16-67 on the heads
            for mod in hd:
                if isinstance(mod, (nn.Conv1d, nn.Linear)):
                    nn.init.orthogonal_(mod.weight, gain=nn.init.calculate_gain("relu"))
                    nn.init.zeros_(mod.bias)
            nn.init.orthogonal_(hd[7].weight, gain=last_gain)
        self._architecture_name = {(56, 2, 4, 128): "transformer_b_s", (96, 5, 8, 256): "transformer_b_l"}.get(
            (embed_dim, num_layers, num_heads, head_hidden_dim), f"transformer_{embed_dim}x{num_layers}x{num_heads}_{head_hidden_dim}")
        self._architecture_params = {"obs_shape": list(self.obs_shape), "action_dim": self.action_dim}

    def forward_body(self, x):
        x = self.cell_embed(x).flatten(2).transpose(1, 2) + self.pos_embed
        return self.transformer(x)

    def forward(self, obs, action_mask=None):
        features = self.forward_body(obs).transpose(1, 2)
        logits = self.policy_head(features)
        value = self.value_head(features)
        if action_mask is not None and action_mask.dim() == 1 and logits.dim() == 2:
            action_mask = action_mask.unsqueeze(0)
        return MaskedCategorical(logits, action_mask), value


# the reference's registry names (src/utils/model_export.py:17-35) for the convolutional families
ARCHITECTURES = {
    "resnet_b_s": lambda obs_shape, action_dim: ResNetActorCritic(obs_shape, action_dim, 32, 4, 128),
    "resnet_b_l": lambda obs_shape, action_dim: ResNetActorCritic(obs_shape, action_dim, 80, 5, 256),
    "cnn_b_s": lambda obs_shape, action_dim: CnnActorCritic(obs_shape, action_dim, (56,) * 4, 128),
    "cnn_b_l": lambda obs_shape, action_dim: CnnActorCritic(obs_shape, action_dim, (96,) * 8, 256),
    "transformer_b_s": lambda obs_shape, action_dim: TransformerActorCritic(obs_shape, action_dim, 56, 2, 4, 128),
    "transformer_b_l": lambda obs_shape, action_dim: TransformerActorCritic(obs_shape, action_dim, 96, 5, 8, 256),
    # the older, non-"_b_" entries of the same registry (resnet.py:96-112, cnn.py:82-109)
    "resnet_s": lambda obs_shape, action_dim: ResNetActorCritic(obs_shape, action_dim, 64, 4, 256),
    "resnet_l": lambda obs_shape, action_dim: ResNetActorCritic(obs_shape, action_dim, 128, 8, 256),
    "cnn_s": lambda obs_shape, action_dim: CnnActorCritic(obs_shape, action_dim, (64,) * 4, 256),
    "cnn_l": lambda obs_shape, action_dim: CnnActorCritic(obs_shape, action_dim, (192,) * 6, 256),
}


def build_architecture(name: str, obs_shape, action_dim) -> nn.Module:
    if name not in ARCHITECTURES:
        raise ValueError(f"Unknown architecture: {name}. Known architectures: {', '.join(sorted(ARCHITECTURES))}")
    return ARCHITECTURES[name](obs_shape, action_dim)

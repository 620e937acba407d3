"""tcgen05 forward of the reference's wider convolutional networks, fed from bitboards.

``NativeConvNet`` takes a torch module with the reference's residual layout (``conv_in`` / ``res_blocks`` /
``policy_head`` / ``value_head``; src/alg/architectures/resnet.py:24-65) or its plain CNN layout (``shared_body`` /
``actor`` / ``critic``; src/alg/architectures/cnn.py:7-58) whose 3x3 convolutions all have the same width of at most
96 channels -- "resnet_b_l" (80 x 5 blocks), "cnn_b_s" (56 x 4, zero-padded to 64), "cnn_b_l" (96 x 8), "resnet_s" /
"cnn_s" (64; configs.py:36-65, resnet.py:96-103, cnn.py:82-89) -- folds eval-mode BatchNorm into the convolutions and
runs the whole convolutional body plus the heads' 1x1 convolutions as ONE kernel (``mnk_conv_tower``,
csrc/mnk_convtower.cu) straight from the packed env state.  The heads' LayerNorm / Linear tails run on ``mnk_resnet_heads_mma``
where they have its shape (width 128, boards up to 96 cells: cnn_b_s) and through the original torch modules otherwise
(width 256: plain library GEMMs, ~1 % of the forward's FLOPs).  Inference only: NNPolicy /
opponent / evaluation; the 32-channel default network has its own kernels (``mnk_b200.resnet``).
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch
import torch.nn as nn

from . import _lib
from ._lib import MnkState, check
from .resnet import _fold, mma_head_params, operand_dtype, run_mma_heads
from .sampling import MaskedCategorical

KERNEL_WIDTHS = (64, 80, 96)           # channel counts mnk_conv_tower is compiled for (narrower towers are zero-padded)


def _layers_of(model: nn.Module):
    """[(conv, bn)] of the body, the two heads, and whether the body is residual."""
    if hasattr(model, "res_blocks") and hasattr(model, "conv_in"):
        convs = [(model.conv_in[0], model.conv_in[1])]
        for blk in model.res_blocks:
            convs += [(blk.conv1, blk.bn1), (blk.conv2, blk.bn2)]
        return convs, model.policy_head, model.value_head, True
    if hasattr(model, "shared_body"):
        mods = list(model.shared_body)
        convs = [(mods[i], mods[i + 1]) for i in range(0, len(mods), 3)]
        if any(not isinstance(c, nn.Conv2d) or not isinstance(b, nn.BatchNorm2d) for c, b in convs):
            raise ValueError("NativeConvNet: shared_body must be a stack of Conv2d + BatchNorm2d + ReLU")
        return convs, model.actor, model.critic, False
    raise ValueError("NativeConvNet: the module has neither the reference's ResNet layout nor its CNN layout")


def supports(model: nn.Module) -> bool:
    """True if `model` is a convolutional network mnk_conv_tower can run (see the module docstring)."""
    try:
        convs, _, _, _ = _layers_of(model)
    except ValueError:
        return False
    width = convs[0][0].out_channels
    return (all(c.out_channels == width and tuple(c.kernel_size) == (3, 3) and tuple(c.padding) == (1, 1) for c, _ in convs)
            and 32 < width <= KERNEL_WIDTHS[-1])


class NativeConvNet:
    def __init__(self, model: nn.Module, device="cuda"):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("mnk_b200.NativeConvNet: CUDA only (no CPU fallback)")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self._dev = dev
        self._L = _lib.lib()
        self.bn_mode = "eval"
        self.use_mma_heads = True
        self.refresh(model)

    @torch.no_grad()
    def refresh(self, model: nn.Module):
        """(Re)import weights -- call after the learner updated `model`.  Device tensors keep their addresses."""
        if not supports(model):
            raise ValueError("NativeConvNet supports towers of equal-width 3x3 convolutions with 33..96 channels "
                             "(resnet_b_l, cnn_b_s, cnn_b_l, resnet_s, cnn_s); the 32-channel resnet_b_s runs on NativeResNet")
        convs, ph, vh, residual = _layers_of(model)
        width = convs[0][0].out_channels
        C = next(w for w in KERNEL_WIDTHS if w >= width)
        dev, op = self._dev, operand_dtype()
        self.width, self.channels, self.layers, self.residual = width, C, len(convs), residual
        weights = torch.zeros((len(convs), 9, C // 8, C, 8), dtype=torch.float32, device=dev)
        bias = torch.zeros((len(convs), C), dtype=torch.float32, device=dev)
        for i, (conv, bn) in enumerate(convs):
            w, b = _fold(conv, bn)                                   # [c_out][c_in][3][3], [c_out]
            full = torch.zeros((C, C, 3, 3), dtype=torch.float32, device=dev)
            full[:w.shape[0], :w.shape[1]] = w.to(dev)
            # [tap = ky*3+kx][k-chunk][c_out][8 c_in]
            weights[i] = full.permute(2, 3, 1, 0).reshape(9, C // 8, 8, C).permute(0, 1, 3, 2)
            bias[i, :width] = b.to(dev)
        head_w = torch.zeros((3, C), dtype=torch.float32, device=dev)
        head_w[:2, :width] = ph[0].weight.detach().reshape(2, width).float().to(dev)
        head_w[2, :width] = vh[0].weight.detach().reshape(width).float().to(dev)
        head_b = torch.cat([ph[0].bias.detach().reshape(2), vh[0].bias.detach().reshape(1)]).float().to(dev)
        fresh = {"weights": weights.to(op).contiguous(), "bias": bias.contiguous(), "head_w": head_w.contiguous(),
                 "head_b": head_b.contiguous()}
        fresh.update(mma_head_params(ph, vh, dev))    # tcgen05 head tails where they fit (width 128, <= 96 cells: cnn_b_s)
        old = getattr(self, "_params", None)
        if old is not None and all(old[k].shape == v.shape and old[k].dtype == v.dtype for k, v in fresh.items()):
            for k, v in fresh.items():
                old[k].copy_(v)
        else:
            self._params = fresh
            self._err = torch.zeros(1, dtype=torch.int32, device=dev)
        self.policy_tail = nn.Sequential(*list(ph)[2:]).to(dev).eval()        # LN, ReLU, Linear, LN, ReLU, Linear
        self.value_tail = nn.Sequential(*list(vh)[2:]).to(dev).eval()          # ... + Tanh
        self.version = getattr(self, "version", 0) + 1

    def pointer_signature(self):
        return tuple(t.data_ptr() for t in self._params.values()) + (self._err.data_ptr(),)

    @torch.no_grad()
    def features(self, state: MnkState, num_envs: int, cells: int, swap: Optional[torch.Tensor]):
        pf = torch.empty((num_envs, 2 * cells), dtype=torch.float32, device=self._dev)
        vf = torch.empty((num_envs, cells), dtype=torch.float32, device=self._dev)
        P = self._params
        with torch.cuda.device(self._dev):
            check(self._L.mnk_conv_tower(ctypes.byref(state), None if swap is None else swap.data_ptr(), self.channels,
                                         self.layers, int(self.residual), P["weights"].data_ptr(), P["bias"].data_ptr(),
                                         P["head_w"].data_ptr(), P["head_b"].data_ptr(), pf.data_ptr(), vf.data_ptr(),
                                         self._err.data_ptr(), torch.cuda.current_stream(self._dev).cuda_stream),
                  "mnk_conv_tower")
        return pf, vf

    @torch.no_grad()
    def tails(self, pf: torch.Tensor, vf: torch.Tensor, want_value: bool = True):
        """Head tails: mnk_resnet_heads_mma for heads of width 128 on boards up to 96 cells, else the torch modules."""
        if self.use_mma_heads and "hm_w2" in self._params:
            return run_mma_heads(self._L, self._params, pf, vf, want_value, self._err, self._dev)
        return self.policy_tail(pf), (self.value_tail(vf) if want_value else None)

    @torch.no_grad()
    def forward_env(self, env, swap: Optional[torch.Tensor] = None, want_value: bool = True):
        """Raw policy logits f32[N, m*n] and value f32[N, 1] for the CURRENT state of `env`, read from its bitboards."""
        env._fold_mirrors()
        pf, vf = self.features(env._st, env.num_envs, env.m * env.n, swap)
        return self.tails(pf, vf, want_value)

    @torch.no_grad()
    def forward(self, obs: torch.Tensor, action_mask: Optional[torch.Tensor] = None, want_value: bool = True):
        """Module-compatible forward(obs f32[B,2,m,n], mask) -> (MaskedCategorical, value[B,1])."""
        if obs.dim() == 3:
            obs = obs.unsqueeze(0)
        b, _, m, n = obs.shape
        words = self._L.mnk_state_words(m, n)
        bits = torch.empty((2, words, b), dtype=torch.int64, device=self._dev)
        meta = torch.zeros(b, dtype=torch.int32, device=self._dev)
        st = MnkState(m, n, 1, words, b, bits.data_ptr(), meta.data_ptr())
        obs = obs.to(self._dev, torch.float32).contiguous()
        with torch.cuda.device(self._dev):
            check(self._L.mnk_pack_boards(ctypes.byref(st), obs.data_ptr(), torch.cuda.current_stream(self._dev).cuda_stream),
                  "mnk_pack_boards")
        pf, vf = self.features(st, b, m * n, None)
        logits, value = self.tails(pf, vf, want_value)
        if action_mask is not None and action_mask.dim() == 1:
            action_mask = action_mask.unsqueeze(0)
        return MaskedCategorical(logits, action_mask), value

    __call__ = forward

    def check_error(self):
        if int(self._err.item()) != 0:
            raise RuntimeError("mnk_conv_tower: internal barrier wait timed out")


def native_network(model: nn.Module, device="cuda", **kwargs):
    """The tcgen05 forward for `model`: NativeResNet for the 32-channel default network, NativeConvNet for the wider
    convolutional ones, NativeTransformer for transformer_b_s / transformer_b_l; raises for anything else."""
    from .resnet import NativeResNet
    from . import transformer
    if supports(model):
        return NativeConvNet(model, device=device)
    if transformer.supports(model):
        return transformer.NativeTransformer(model, device=device)
    return NativeResNet(model, device=device, **kwargs)

"""tcgen05 forward of the reference's transformer networks, fed from bitboards.

``NativeTransformer`` takes a torch module with the reference's layout (``cell_embed``, ``pos_embed``,
``transformer.layers[i]`` = pre-norm ``nn.TransformerEncoderLayer`` with ReLU, ``policy_head`` / ``value_head`` opening
with a ``Conv1d``; src/alg/architectures/transformer.py:7-92) of one of the two registered shapes -- "transformer_b_s"
(embed 56, 4 heads) and "transformer_b_l" (embed 96, 8 heads; configs.py:7-25) -- pads head_dim to 16 and embed_dim to a
multiple of 16, lays every weight matrix out as UMMA B-operand tiles in the order the kernel consumes them, and runs the
embedding, all encoder layers and the heads' 1x1 convolutions as ONE kernel (``mnk_transformer_body``,
csrc/mnk_transformer.cu).  The heads' LayerNorm / Linear tails run on ``mnk_resnet_heads_mma`` where they have its shape
(width 128, boards up to 96 cells: transformer_b_s) and through the original torch modules otherwise.  Inference only.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch
import torch.nn as nn

from . import _lib
from ._lib import MnkState, check
from .resnet import _arrange_linear, mma_head_params, operand_dtype, run_mma_heads
from .sampling import MaskedCategorical

SHAPES = ((56, 4), (96, 8))            # (embed_dim, heads) mnk_transformer_body is compiled for
MAX_TOKENS = 128


def supports(model: nn.Module) -> bool:
    layers = getattr(getattr(model, "transformer", None), "layers", None)
    if layers is None or not hasattr(model, "cell_embed") or not hasattr(model, "pos_embed"):
        return False
    l0 = layers[0]
    shape = (l0.self_attn.embed_dim, l0.self_attn.num_heads)
    plain = all(getattr(l, "norm_first", False) and l.linear1.out_features == 4 * shape[0] and
                l.self_attn._qkv_same_embed_dim and l.self_attn.in_proj_bias is not None and
                getattr(l.activation, "__name__", "") == "relu" for l in layers)
    return plain and shape in SHAPES and getattr(model.transformer, "norm", None) is None and model.pos_embed.shape[1] <= MAX_TOKENS


class NativeTransformer:
    def __init__(self, model: nn.Module, device="cuda"):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("mnk_b200.NativeTransformer: CUDA only (no CPU fallback)")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self._dev = dev
        self._L = _lib.lib()
        self.bn_mode = "eval"
        self.use_mma_heads = True
        self.refresh(model)

    @torch.no_grad()
    def refresh(self, model: nn.Module):
        if not supports(model):
            raise ValueError("NativeTransformer supports the reference's pre-norm ReLU encoder with (embed_dim, heads) = (56, 4) "
                             "or (96, 8) on boards of at most 128 cells (transformer_b_s / transformer_b_l)")
        dev, op = self._dev, operand_dtype()
        layers = list(model.transformer.layers)
        D, NH = layers[0].self_attn.embed_dim, layers[0].self_attn.num_heads
        DH, DP, QP, F = D // NH, (D + 15) // 16 * 16, 16 * NH, 4 * D
        nsplit = 2 if 3 * QP > 256 else 1
        self.embed_dim, self.heads, self.layers, self.tokens = D, NH, len(layers), model.pos_embed.shape[1]
        f = lambda t: t.detach().float().to(dev)
        pad = lambda v, n: torch.cat([f(v).reshape(-1), torch.zeros(n - v.numel(), device=dev)])
        head_cols = (torch.arange(NH, device=dev)[:, None] * 16 + torch.arange(DH, device=dev)[None, :]).reshape(-1)   # h*16 + d
        tiles, params = [], []
        for l in layers:
            w_in, b_in = f(l.self_attn.in_proj_weight), f(l.self_attn.in_proj_bias)         # [3D, D], [3D]
            wqkv = torch.zeros((3 * QP, DP), device=dev)
            bqkv = torch.zeros(3 * QP, device=dev)
            for which in range(3):
                wqkv[which * QP + head_cols, :D] = w_in[which * D:(which + 1) * D]
                bqkv[which * QP + head_cols] = b_in[which * D:(which + 1) * D]
            nq = 3 * QP // nsplit
            tiles += [_arrange_linear(wqkv[j * nq:(j + 1) * nq], DP, nq) for j in range(nsplit)]
            wo = torch.zeros((DP, QP), device=dev)                                          # [n = D_out][k = h*16 + d]
            wo[:D, head_cols] = f(l.self_attn.out_proj.weight)
            tiles.append(_arrange_linear(wo, QP, DP))
            w1 = torch.zeros((F, DP), device=dev)
            w1[:, :D] = f(l.linear1.weight)
            nf = F // nsplit
            tiles += [_arrange_linear(w1[j * nf:(j + 1) * nf], DP, nf) for j in range(nsplit)]
            w2 = torch.zeros((DP, F), device=dev)
            w2[:D] = f(l.linear2.weight)
            tiles += [_arrange_linear(w2[:, j * nf:(j + 1) * nf].contiguous(), nf, DP) for j in range(nsplit)]
            params.append(torch.cat([pad(l.norm1.weight, DP), pad(l.norm1.bias, DP), bqkv, pad(l.self_attn.out_proj.bias, DP),
                                     pad(l.norm2.weight, DP), pad(l.norm2.bias, DP), f(l.linear1.bias), pad(l.linear2.bias, DP)]))
        weights = torch.cat([t.reshape(-1) for t in tiles]).contiguous()
        want_w = int(self._L.mnk_transformer_layer_weight_bytes(D, NH)) * len(layers)
        want_p = int(self._L.mnk_transformer_layer_params(D, NH))
        assert weights.numel() * weights.element_size() == want_w and params[0].numel() == want_p, "layout mismatch with the kernel"
        ph, vh = model.policy_head, model.value_head
        ce = f(model.cell_embed.weight).reshape(D, -1)
        embed = torch.zeros((3, DP), device=dev)
        embed[0, :D], embed[1, :D], embed[2, :D] = ce[:, 0], ce[:, 1], f(model.cell_embed.bias)
        pos = torch.zeros((self.tokens, DP), device=dev)
        pos[:, :D] = f(model.pos_embed)[0]
        head_w = torch.zeros((3, DP), device=dev)
        head_w[:2, :D] = f(ph[0].weight).reshape(2, D)
        head_w[2, :D] = f(vh[0].weight).reshape(D)
        fresh = {"weights": weights, "layer_params": torch.stack(params).contiguous(), "embed": embed.contiguous(),
                 "pos": pos.contiguous(), "head_w": head_w.contiguous(),
                 "head_b": torch.cat([f(ph[0].bias).reshape(2), f(vh[0].bias).reshape(1)]).contiguous()}
        fresh.update(mma_head_params(ph, vh, dev))
        old = getattr(self, "_params", None)
        if old is not None and all(old[k].shape == v.shape and old[k].dtype == v.dtype for k, v in fresh.items()):
            for k, v in fresh.items():
                old[k].copy_(v)
        else:
            self._params = fresh
            self._err = torch.zeros(1, dtype=torch.int32, device=dev)
        self.policy_tail = nn.Sequential(*list(ph)[2:]).to(dev).eval()
        self.value_tail = nn.Sequential(*list(vh)[2:]).to(dev).eval()
        self.version = getattr(self, "version", 0) + 1

    def pointer_signature(self):
        return tuple(t.data_ptr() for t in self._params.values()) + (self._err.data_ptr(),)

    @torch.no_grad()
    def features(self, state: MnkState, num_envs: int, cells: int, swap: Optional[torch.Tensor]):
        if cells != self.tokens:
            raise ValueError(f"NativeTransformer: the network was built for {self.tokens} cells, the env has {cells}")
        pf = torch.empty((num_envs, 2 * cells), dtype=torch.float32, device=self._dev)
        vf = torch.empty((num_envs, cells), dtype=torch.float32, device=self._dev)
        P = self._params
        with torch.cuda.device(self._dev):
            check(self._L.mnk_transformer_body(ctypes.byref(state), None if swap is None else swap.data_ptr(), self.embed_dim,
                                               self.heads, self.layers, P["weights"].data_ptr(), P["layer_params"].data_ptr(),
                                               P["embed"].data_ptr(), P["pos"].data_ptr(), P["head_w"].data_ptr(),
                                               P["head_b"].data_ptr(), pf.data_ptr(), vf.data_ptr(), self._err.data_ptr(),
                                               torch.cuda.current_stream(self._dev).cuda_stream), "mnk_transformer_body")
        return pf, vf

    @torch.no_grad()
    def tails(self, pf: torch.Tensor, vf: torch.Tensor, want_value: bool = True):
        if self.use_mma_heads and "hm_w2" in self._params:
            return run_mma_heads(self._L, self._params, pf, vf, want_value, self._err, self._dev)
        return self.policy_tail(pf), (self.value_tail(vf) if want_value else None)

    @torch.no_grad()
    def forward_env(self, env, swap: Optional[torch.Tensor] = None, want_value: bool = True):
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
            raise RuntimeError("mnk_transformer_body: internal barrier wait timed out")

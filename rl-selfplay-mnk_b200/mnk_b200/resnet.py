"""tcgen05 forward of the reference's default network, fed from bitboards.

``NativeResNet`` takes a torch module with the reference's ResNet layout (``conv_in``,
``res_blocks[i].conv1/bn1/conv2/bn2``, ``policy_head``, ``value_head``; src/alg/architectures/
resnet.py:24-65) with 32 channels, folds eval-mode BatchNorm into the convolutions, lays the bf16
weights out for the UMMA B operand and runs the whole convolutional part -- 98% of the forward's
FLOPs -- as ONE kernel (``mnk_resnet_tower``, csrc/mnk_resnet.cu) straight from the packed env
state.  The heads' LayerNorm / Linear stack (plain small GEMMs, 2% of the FLOPs) runs through the
original torch modules on the features the kernel writes.  Inference only (NNPolicy / opponent /
rollout forward); the learner keeps the torch module.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import _lib
from ._lib import MnkBnTrain, MnkHeadsWeights, MnkState, check
from .policy import Policy
from .sampling import MaskedCategorical, fresh_seed, masked_sample


def _fold(conv: nn.Conv2d, bn: nn.BatchNorm2d) -> Tuple[torch.Tensor, torch.Tensor]:
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    w = conv.weight * scale[:, None, None, None]
    b = (conv.bias - bn.running_mean) * scale + bn.bias
    return w.float(), b.float()


def operand_dtype() -> torch.dtype:
    """Element type of the tower kernels' tensor-core operands as compiled into the library (fp16 by default; bf16
    with -DMNK_ACT_BF16): the conv weights are laid out in it."""
    return torch.float16 if _lib.lib().mnk_resnet_operand_dtype() == 0 else torch.bfloat16


def _arrange(w: torch.Tensor) -> torch.Tensor:
    """[c_out=32][c_in<=32][3][3] fp32 -> operand type [tap][k-chunk=4][c_out][8 c_in] (zero-padded input channels)."""
    c_out, c_in = w.shape[0], w.shape[1]
    full = torch.zeros((c_out, 32, 3, 3), dtype=torch.float32, device=w.device)
    full[:, :c_in] = w
    t = full.permute(2, 3, 1, 0).reshape(9, 4, 8, c_out)          # [tap][kc][j][c_out]
    return t.permute(0, 1, 3, 2).contiguous().to(operand_dtype())   # [tap][kc][c_out][j]


def _arrange_rows(w: torch.Tensor) -> torch.Tensor:
    """Same parameters for the board-row kernel (mnk_resnet_tower_rows): operand type [kx][k-chunk=4][ky*32 + c_out][8 c_in]."""
    c_out, c_in = w.shape[0], w.shape[1]
    full = torch.zeros((c_out, 32, 3, 3), dtype=torch.float32, device=w.device)
    full[:, :c_in] = w
    t = full.permute(3, 1, 2, 0).reshape(3, 4, 8, 3, c_out)        # [kx][kc][j][ky][c_out]
    return t.permute(0, 1, 3, 4, 2).reshape(3, 4, 3 * c_out, 8).contiguous().to(operand_dtype())


def _arrange_linear(w: torch.Tensor, kpad: int, npad: int) -> torch.Tensor:
    """Linear.weight [n_out][k_in] fp32 -> UMMA B operand, K-major: operand type [kpad/8][npad][8] (zero padded)."""
    n, k = w.shape
    full = torch.zeros((npad, kpad), dtype=torch.float32, device=w.device)
    full[:n, :k] = w
    return full.view(npad, kpad // 8, 8).permute(1, 0, 2).contiguous().to(operand_dtype())


MMA_HEADS_MAX_CELLS = 96               # mnk_resnet_heads_mma (shared-memory bound); larger boards use the fp32 heads kernel


def mma_head_params(ph: nn.Sequential, vh: nn.Sequential, dev) -> dict:
    """Operands of mnk_resnet_heads_mma for two heads of the reference's layout (Conv2d 1x1, Flatten, LayerNorm, ReLU,
    Linear(.., 128), LayerNorm, ReLU, Linear): the three Linear weights as UMMA B operands + one parameter vector.
    Empty when the kernel does not cover the heads (hidden width != 128 or more than 96 cells)."""
    cells = ph[7].out_features
    if cells > MMA_HEADS_MAX_CELLS or ph[4].out_features != 128 or vh[4].out_features != 128:
        return {}
    f = lambda t: t.detach().float().to(dev).contiguous()
    r16 = lambda v: (v + 15) // 16 * 16
    return {
        "hm_w1p": _arrange_linear(f(ph[4].weight), r16(2 * cells), 128),
        "hm_w1v": _arrange_linear(f(vh[4].weight), r16(cells), 128),
        "hm_w2": _arrange_linear(f(ph[7].weight), 128, r16(cells)),
        "hm_params": torch.cat([f(t).reshape(-1) for t in (
            ph[2].weight, ph[2].bias, vh[2].weight, vh[2].bias, ph[4].bias, vh[4].bias, ph[5].weight, ph[5].bias,
            vh[5].weight, vh[5].bias, vh[7].weight, ph[7].bias, vh[7].bias)]).contiguous(),
    }


def run_mma_heads(L, P: dict, pf: torch.Tensor, vf: torch.Tensor, want_value: bool, err: torch.Tensor, dev):
    rows, cells = vf.shape
    logits = torch.empty((rows, cells), dtype=torch.float32, device=dev)
    values = torch.empty((rows, 1), dtype=torch.float32, device=dev) if want_value else None
    with torch.cuda.device(dev):
        check(L.mnk_resnet_heads_mma(pf.data_ptr(), vf.data_ptr(), rows, cells, P["hm_w1p"].data_ptr(), P["hm_w1v"].data_ptr(),
                                     P["hm_w2"].data_ptr(), P["hm_params"].data_ptr(), logits.data_ptr(),
                                     values.data_ptr() if want_value else None, err.data_ptr(),
                                     torch.cuda.current_stream(dev).cuda_stream), "mnk_resnet_heads_mma")
    return logits, values
ROWS_KERNEL_BOARD_ROWS = (3, 10)      # mnk_resnet_tower_rows: boards with 3 <= m <= 10 rows (shared-memory bound)


class NativeResNet:
    """bn_mode="eval": BatchNorm folded with the running statistics (NNPolicy / the frozen opponent / evaluation).
    bn_mode="train": BatchNorm with the statistics of the CURRENT batch and running-statistics updates on the device
    copies -- what the reference's rollout forward does (src/alg/ppo.py:97 never leaves train mode); one tcgen05
    launch per layer (mnk_resnet_tower_train).  `export_running_stats(model)` writes the statistics back."""

    def __init__(self, model: nn.Module, device="cuda", torch_heads: bool = False, bn_mode: str = "eval"):
        if bn_mode not in ("eval", "train"):
            raise ValueError("bn_mode must be 'eval' or 'train'")
        self.bn_mode = bn_mode
        self.train_forwards = 0             # train-mode forwards since the last export_running_stats
        self._scratch = None
        self.torch_heads = torch_heads      # run the head tails through the original torch modules (debug / comparison)
        self.use_mma_heads = True           # tcgen05 head GEMMs where the board fits (<= 96 cells); False = fp32 heads kernel
        self.use_rows_kernel = True         # board-row tower kernel where the board fits it (m <= 10); False = tap kernel
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("mnk_b200.NativeResNet: CUDA only (no CPU fallback)")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self._dev = dev
        self._L = _lib.lib()
        self.refresh(model)

    @torch.no_grad()
    def refresh(self, model: nn.Module):
        """(Re)import weights -- call after the learner updated `model`."""
        convs = [(model.conv_in[0], model.conv_in[1])]
        for blk in model.res_blocks:
            convs += [(blk.conv1, blk.bn1), (blk.conv2, blk.bn2)]
        if any(c.out_channels != 32 or tuple(c.kernel_size) != (3, 3) for c, _ in convs):
            raise ValueError("NativeResNet supports the 32-channel 3x3 tower (resnet_b_s)")
        self.blocks = len(model.res_blocks)
        folded = [_fold(c, b) for c, b in convs]
        dev = self._dev
        ph, vh = model.policy_head, model.value_head
        pc, vc = ph[0], vh[0]
        if ph[4].out_features != 128 or vh[4].out_features != 128:
            raise ValueError("NativeResNet supports head_hidden_dim = 128 (resnet_b_s)")
        f = lambda t: t.detach().float().to(dev).contiguous()
        fresh = {
            "weights": torch.stack([_arrange(w.to(dev)) for w, _ in folded]).contiguous(),            # bf16 [L][9][4][32][8]
            "weights_rows": torch.stack([_arrange_rows(w.to(dev)) for w, _ in folded]).contiguous(),  # bf16 [L][3][4][96][8]
            "bias": torch.stack([b.to(dev) for _, b in folded]).contiguous(),                         # f32 [L][32]
            "head_w": torch.cat([pc.weight.reshape(2, 32), vc.weight.reshape(1, 32)]).float().to(dev).contiguous(),
            "head_b": torch.cat([pc.bias.reshape(2), vc.bias.reshape(1)]).float().to(dev).contiguous(),
            # fused heads kernel (mnk_resnet_heads): LN / Linear parameters, Linear weights transposed to [in][out]
            "p_ln1_w": f(ph[2].weight), "p_ln1_b": f(ph[2].bias), "p_w1t": f(ph[4].weight.t()), "p_b1": f(ph[4].bias),
            "p_ln2_w": f(ph[5].weight), "p_ln2_b": f(ph[5].bias), "p_w2t": f(ph[7].weight.t()), "p_b2": f(ph[7].bias),
            "v_ln1_w": f(vh[2].weight), "v_ln1_b": f(vh[2].bias), "v_w1t": f(vh[4].weight.t()), "v_b1": f(vh[4].bias),
            "v_ln2_w": f(vh[5].weight), "v_ln2_b": f(vh[5].bias), "v_w2": f(vh[7].weight.reshape(-1)), "v_b2": f(vh[7].bias),
        }
        fresh.update(mma_head_params(ph, vh, dev))    # tcgen05 heads where the board fits
        if self.bn_mode == "train":          # unfolded conv weights + BatchNorm parameters / running statistics, [L][32]
            bns = [b for _, b in convs]
            if any(b.momentum != bns[0].momentum or b.eps != bns[0].eps or not b.track_running_stats for b in bns):
                raise ValueError("NativeResNet(bn_mode='train') needs one momentum / eps and tracked running statistics")
            self._bn_momentum, self._bn_eps = float(bns[0].momentum), float(bns[0].eps)
            fresh.update({
                "raw_rows": torch.stack([_arrange_rows(c.weight.detach().float().to(dev)) for c, _ in convs]).contiguous(),
                "bn_gamma": torch.stack([f(b.weight) for b in bns]), "bn_beta": torch.stack([f(b.bias) for b in bns]),
                "conv_bias": torch.stack([f(c.bias) for c, _ in convs]),
                "running_mean": torch.stack([f(b.running_mean) for b in bns]),
                "running_var": torch.stack([f(b.running_var) for b in bns]),
                "batch_stats": torch.zeros((len(bns), 64), dtype=torch.float32, device=dev),
            })
        # Device tensors live at STABLE addresses: a refresh copies into the existing storage, so raw pointers baked
        # into a captured CUDA graph (RolloutCollector.collect(graph=True)) keep reading current weights.  A change of
        # shape (another board size / block count) re-allocates and shows up in pointer_signature().
        old = getattr(self, "_params", None)
        if old is not None and all(old[k].shape == v.shape and old[k].dtype == v.dtype for k, v in fresh.items()):
            for k, v in fresh.items():
                old[k].copy_(v)
        else:
            self._params = fresh
            self._err = torch.zeros(1, dtype=torch.int32, device=dev)
        P = self._params
        self.weights, self.weights_rows, self.bias = P["weights"], P["weights_rows"], P["bias"]
        self.head_w, self.head_b = P["head_w"], P["head_b"]
        self._head_tensors = {n: P[n] for n in MnkHeadsWeights.NAMES}
        self._heads = MnkHeadsWeights(*[P[n].data_ptr() for n in MnkHeadsWeights.NAMES])
        self.policy_tail = nn.Sequential(*list(ph)[2:]).to(dev).eval()        # LN, ReLU, Linear, LN, ReLU, Linear
        self.value_tail = nn.Sequential(*list(vh)[2:]).to(dev).eval()          # ... + Tanh
        self.version = getattr(self, "version", 0) + 1
        if self.bn_mode == "train":
            self._bn = MnkBnTrain(P["bn_gamma"].data_ptr(), P["bn_beta"].data_ptr(), P["conv_bias"].data_ptr(),
                                  P["running_mean"].data_ptr(), P["running_var"].data_ptr(), P["batch_stats"].data_ptr(),
                                  self._bn_momentum, self._bn_eps)
            self.train_forwards = 0

    @torch.no_grad()
    def export_running_stats(self, model: nn.Module):
        """Write the running statistics the train-mode forwards accumulated on the device back into `model`'s
        BatchNorm buffers (and advance num_batches_tracked by the number of forwards)."""
        if self.bn_mode != "train":
            return
        bns = [model.conv_in[1]] + [b for blk in model.res_blocks for b in (blk.bn1, blk.bn2)]
        for i, b in enumerate(bns):
            b.running_mean.copy_(self._params["running_mean"][i])
            b.running_var.copy_(self._params["running_var"][i])
            if b.num_batches_tracked is not None:
                b.num_batches_tracked += self.train_forwards
        self.train_forwards = 0

    def running_state(self):
        """Copies of the device-side running statistics (+ the forward count): restore_running_state() undoes forwards
        that must not count, e.g. the allocation warm-up before a CUDA-graph capture."""
        if self.bn_mode != "train":
            return None
        return self._params["running_mean"].clone(), self._params["running_var"].clone(), self.train_forwards

    def restore_running_state(self, state):
        if state is not None:
            self._params["running_mean"].copy_(state[0])
            self._params["running_var"].copy_(state[1])
            self.train_forwards = state[2]

    def _train_scratch(self, m: int, n: int, num_envs: int) -> torch.Tensor:
        need = int(self._L.mnk_resnet_tower_train_scratch_bytes(m, n, num_envs, self.blocks))
        check(need if need < 0 else 0, "mnk_resnet_tower_train_scratch_bytes")
        if self._scratch is None or self._scratch.numel() < need:
            self._scratch = torch.zeros(need, dtype=torch.uint8, device=self._dev)    # (zeroed: sticky post-mortem words)
        return self._scratch

    def pointer_signature(self):
        """Addresses of every device tensor a captured launch reads (see RolloutCollector._collect_graphed)."""
        scratch = (self._scratch.data_ptr(),) if self._scratch is not None else ()
        return tuple(t.data_ptr() for t in self._params.values()) + (self._err.data_ptr(),) + scratch

    @torch.no_grad()
    def tails(self, pf: torch.Tensor, vf: torch.Tensor, want_value: bool = True):
        """logits f32[N, A], value f32[N, 1] (None when `want_value` is False) from the tower's head features."""
        if self.torch_heads:
            return self.policy_tail(pf), (self.value_tail(vf) if want_value else None)
        if self.use_mma_heads and "hm_w2" in self._params:
            return run_mma_heads(self._L, self._params, pf, vf, want_value, self._err, self._dev)
        rows, cells = vf.shape
        logits = torch.empty((rows, cells), dtype=torch.float32, device=self._dev)
        values = torch.empty((rows, 1), dtype=torch.float32, device=self._dev) if want_value else None
        with torch.cuda.device(self._dev):
            check(self._L.mnk_resnet_heads(pf.data_ptr(), vf.data_ptr(), rows, cells, ctypes.byref(self._heads),
                                            logits.data_ptr(), values.data_ptr() if want_value else None,
                                            torch.cuda.current_stream(self._dev).cuda_stream), "mnk_resnet_heads")
        return logits, values

    @torch.no_grad()
    def features(self, state: MnkState, num_envs: int, cells: int, swap: Optional[torch.Tensor]):
        pf = torch.empty((num_envs, 2 * cells), dtype=torch.float32, device=self._dev)
        vf = torch.empty((num_envs, cells), dtype=torch.float32, device=self._dev)
        if self.bn_mode == "train":
            scratch = self._train_scratch(state.m, state.n, num_envs)
            with torch.cuda.device(self._dev):
                check(self._L.mnk_resnet_tower_train(
                    ctypes.byref(state), None if swap is None else swap.data_ptr(), self._params["raw_rows"].data_ptr(),
                    ctypes.byref(self._bn), self.head_w.data_ptr(), self.head_b.data_ptr(), self.blocks, scratch.data_ptr(),
                    scratch.numel(), pf.data_ptr(), vf.data_ptr(), self._err.data_ptr(),
                    torch.cuda.current_stream(self._dev).cuda_stream), "mnk_resnet_tower_train")
            self.train_forwards += 1
            return pf, vf
        rows = self.use_rows_kernel and ROWS_KERNEL_BOARD_ROWS[0] <= state.m <= ROWS_KERNEL_BOARD_ROWS[1]
        entry, name, weights = ((self._L.mnk_resnet_tower_rows, "mnk_resnet_tower_rows", self.weights_rows) if rows else
                                (self._L.mnk_resnet_tower, "mnk_resnet_tower", self.weights))
        with torch.cuda.device(self._dev):
            check(entry(ctypes.byref(state), None if swap is None else swap.data_ptr(), weights.data_ptr(),
                        self.bias.data_ptr(), self.head_w.data_ptr(), self.head_b.data_ptr(), self.blocks, pf.data_ptr(),
                        vf.data_ptr(), self._err.data_ptr(), torch.cuda.current_stream(self._dev).cuda_stream), name)
        return pf, vf

    @torch.no_grad()
    def forward_env(self, env, swap: Optional[torch.Tensor] = None, want_value: bool = True):
        """Raw policy logits f32[N, m*n] and value f32[N, 1] for the CURRENT state of `env`, read from its
        bitboards; `swap` u8[N] != 0 exchanges the planes (the canonical view of a white mover/agent).
        `want_value=False` skips the value head (value is None)."""
        env._fold_mirrors()
        pf, vf = self.features(env._st, env.num_envs, env.m * env.n, swap)
        return self.tails(pf, vf, want_value)

    @torch.no_grad()
    def forward(self, obs: torch.Tensor, action_mask: Optional[torch.Tensor] = None, want_value: bool = True):
        """Module-compatible forward(obs f32[B,2,m,n], mask) -> (MaskedCategorical, value[B,1]): the
        observation is packed to bitboards first (mnk_pack_boards), then the same kernel runs."""
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
        """Raises if any launch hit an internal barrier timeout (one device->host read)."""
        if int(self._err.item()) != 0:
            raise RuntimeError("mnk_resnet_tower: internal barrier wait timed out")


class NativeNNPolicy(Policy):
    """NNPolicy (src/selfplay/policy.py:32-54) on the tcgen05 forward: NativeResNet for the 32-channel default network,
    NativeConvNet (mnk_b200.convnet) for the wider residual / plain convolutional architectures."""

    def __init__(self, model: nn.Module, device="cuda", seed: Optional[int] = None):
        from .convnet import native_network
        model.eval()
        self.net = native_network(model, device=device)
        self.seed = fresh_seed() if seed is None else seed
        self._calls = 0
        self.counter_base: Optional[torch.Tensor] = None     # see TorchSelfPlayWrapper.counter_base

    def act_from_env(self, env, counter: int, deterministic: bool = False) -> torch.Tensor:
        """Action for the side to move of every env, straight from the bitboards (the wrapper's dense
        opponent call): canonical view = planes swapped where the mover is white."""
        swap = (env._meta & 1).to(torch.uint8)
        logits, _ = self.net.forward_env(env, swap, want_value=False)
        return masked_sample(logits, env.legal_mask(), seed=self.seed, counter=counter, row_offset=env.env_offset,
                             deterministic=deterministic, want_log_prob=False, counter_base=self.counter_base)[0]

    def act(self, obs, deterministic: bool = False) -> torch.Tensor:
        dist, _ = self.net.forward(obs["observation"], obs["action_mask"], want_value=False)
        self._calls += 1
        return masked_sample(dist._raw, dist._mask, seed=self.seed, counter=self._calls, deterministic=deterministic,
                             want_log_prob=False, counter_base=self.counter_base)[0]

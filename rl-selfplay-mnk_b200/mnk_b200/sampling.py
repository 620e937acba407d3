"""Masked categorical sampling on the GPU (mnk_masked_sample, csrc/mnk_sample.cu).

``MaskedCategorical`` is the Categorical-like object the reference's networks return
(src/alg/architectures/resnet.py:84-94): ``sample()``, ``log_prob()``, ``entropy()`` and
``logits`` (normalised).  sample / log_prob / entropy are single warp-per-row kernel launches; only
``logits`` / ``probs`` (full [B, A] tensors, used by NNPolicy's argmax path and by tests)
fall back to three torch ops.
"""
from __future__ import annotations

import itertools
from typing import Optional, Tuple

import torch

from . import _lib

_auto_counter = itertools.count(1)
_seed_counter = itertools.count(1)


def fresh_seed() -> int:
    """A distinct 64-bit Philox key per call (splitmix64 of a process-wide counter): the default seed of every
    sampler-owning object (RandomPolicy, NNPolicy, NativeNNPolicy, RolloutCollector, MaskedCategorical).  The draws
    are keyed by (seed, global row, counter), and the counters are small per-object call counts -- two objects
    sharing one default seed would hand consecutive plies of a game (agent's draw at step t, opponent's at t) the
    same noise.  Pass an explicit seed for reproducible streams; ranks of a sharded run get the same sequence of
    default seeds and differ by their global row offsets."""
    z = (next(_seed_counter) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return z ^ (z >> 31)


def masked_sample(logits: torch.Tensor, mask: Optional[torch.Tensor], seed: int = 0, counter: Optional[int] = None,
                  row_offset: int = 0, deterministic: bool = False, given: Optional[torch.Tensor] = None,
                  want_log_prob: bool = True, want_entropy: bool = False, counter_base: Optional[torch.Tensor] = None
                  ) -> Tuple[torch.Tensor, Optional[torch.Tensor], Optional[torch.Tensor]]:
    """actions i64[B], log_prob f32[B] | None, entropy f32[B] | None for Categorical(masked logits).
    `counter_base`: optional device int64[1] added to `counter` on the device (CUDA-graph replays)."""
    if not logits.is_cuda:
        raise RuntimeError("mnk_b200.masked_sample: CUDA tensors only (no CPU fallback)")
    if logits.dim() != 2:
        raise ValueError("logits must be [rows, actions]")
    lg = logits if (logits.dtype == torch.float32 and logits.stride(1) == 1) else logits.float().contiguous()
    rows, acts = lg.shape
    mk = None
    if mask is not None:
        mk = mask if mask.dtype in (torch.bool, torch.uint8) else (mask != 0)
        mk = mk.contiguous()
        if mk.shape != (rows, acts):
            raise ValueError(f"mask shape {tuple(mk.shape)} != logits shape {(rows, acts)}")
    dev = lg.device
    if given is not None:
        given = given.to(device=dev, dtype=torch.long).contiguous()
        actions = given
    else:
        actions = torch.empty(rows, dtype=torch.long, device=dev)
    logp = torch.empty(rows, dtype=torch.float32, device=dev) if want_log_prob else None
    ent = torch.empty(rows, dtype=torch.float32, device=dev) if want_entropy else None
    if counter is None:
        counter = next(_auto_counter)
    with torch.cuda.device(dev):
        rc = _lib.lib().mnk_masked_sample(
            lg.data_ptr(), lg.stride(0), None if mk is None else mk.data_ptr(), acts, rows, seed & (2**64 - 1),
            counter & (2**64 - 1), None if counter_base is None else counter_base.data_ptr(), row_offset, int(deterministic), None if given is None else given.data_ptr(),
            actions.data_ptr(), None if logp is None else logp.data_ptr(), None if ent is None else ent.data_ptr(),
            torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "mnk_masked_sample")
    return actions, logp, ent


class MaskedCategorical:
    """Categorical(logits=where(mask, logits, -inf)) with all-masked rows made uniform."""

    def __init__(self, logits: torch.Tensor, action_mask: Optional[torch.Tensor] = None, seed: Optional[int] = None):
        self._raw = logits
        self._mask = action_mask
        self._seed = fresh_seed() if seed is None else seed
        self._normalised: Optional[torch.Tensor] = None

    def sample(self) -> torch.Tensor:
        return masked_sample(self._raw, self._mask, seed=self._seed, want_log_prob=False)[0]

    def sample_with_log_prob(self) -> Tuple[torch.Tensor, torch.Tensor]:
        a, lp, _ = masked_sample(self._raw, self._mask, seed=self._seed)
        return a, lp

    def mode(self) -> torch.Tensor:
        return masked_sample(self._raw, self._mask, deterministic=True, want_log_prob=False)[0]

    def _needs_autograd(self) -> bool:
        return self._raw.requires_grad and torch.is_grad_enabled()

    def log_prob(self, actions: torch.Tensor) -> torch.Tensor:
        if self._needs_autograd() or not self._raw.is_cuda:     # learner side: differentiable torch ops
            return self.logits.gather(1, actions.long().unsqueeze(1)).squeeze(1)
        return masked_sample(self._raw, self._mask, given=actions)[1]

    def entropy(self) -> torch.Tensor:
        if self._needs_autograd() or not self._raw.is_cuda:
            # torch.distributions.Categorical.entropy, operation for operation (so that the learner's gradients equal the
            # reference's bit for bit): clamp keeps -inf out of the product and of its gradient, probs = softmax(logits)
            lg = self.logits
            p_log_p = torch.clamp(lg, min=torch.finfo(lg.dtype).min) * torch.softmax(lg, dim=-1)
            return -p_log_p.sum(-1)
        given = torch.zeros(self._raw.shape[0], dtype=torch.long, device=self._raw.device)
        return masked_sample(self._raw, self._mask, given=given, want_log_prob=False, want_entropy=True)[2]

    @property
    def logits(self) -> torch.Tensor:
        """Normalised masked logits, as torch.distributions.Categorical.logits (resnet.py:84-94)."""
        if self._normalised is None:
            lg = self._raw.float()
            if self._mask is not None:
                lg = torch.where(self._mask.bool(), lg, -torch.inf)
                dead = lg.max(dim=1, keepdim=True)[0] == -torch.inf
                lg = torch.where(dead, torch.zeros_like(lg), lg)
            self._normalised = lg - lg.logsumexp(dim=1, keepdim=True)
        return self._normalised

    @property
    def probs(self) -> torch.Tensor:
        return self.logits.exp()

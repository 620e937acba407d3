"""TorchSelfPlayWrapper on fused sm_100a kernels.

Drop-in for the reference's ``src/selfplay/torch_self_play_wrapper.py:6-115``: same attributes
(``env device num_envs opponent_policy agent_side pending_resets``), ``set_opponent``,
``reset(seed=None, options=None) -> (obs, {})``, ``step(actions) -> (obs, rewards, terminated,
truncated, {})`` and ``get_agent_obs()``, with identical tensors for identical side / action
streams (tests/test_wrapper_gpu.py replays traces recorded from the reference).

One step is two launches around the opponent's ``act`` call -- or ONE launch when the opponent is
this package's RandomPolicy -- and never synchronises with the host.  Differences from the
reference, all confined to what no caller can rely on:
  * the opponent is called ONCE per step on a dense batch of ALL envs (rows where it is not the
    opponent's turn are ignored) instead of on gathered subsets of varying size (:83-94); row-wise
    policies -- every policy in the reference -- give the same actions;
  * sides after a reset come from a counter-based Philox stream keyed by (seed, global env id,
    episode number) instead of ``torch.randint`` on the global generator (:26, :43-45), so they do not
    depend on how envs are sharded over GPUs.  ``next_sides`` injects explicit sides (tests).
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import torch

from . import _lib
from ._lib import MnkSelfplay, check
from .env import TorchVectorMnkEnv, _ptr
from .policy import RandomPolicy
from .sampling import fresh_seed


class TorchSelfPlayWrapper:
    def __init__(self, env: TorchVectorMnkEnv, seed: Optional[int] = None):
        self.env = env
        self.device = env.device
        self.num_envs = env.num_envs
        self.opponent_policy = None
        dev = env._dev
        self._dev = dev
        self._L = _lib.lib()
        self.seed = fresh_seed() if seed is None else int(seed)   # keys the side draws and the fused random opponent
        self._side = torch.zeros(self.num_envs, dtype=torch.uint8, device=dev)
        self.pending_resets = torch.zeros(self.num_envs, dtype=torch.bool, device=dev)    # reference :14
        self._episodes = torch.zeros(self.num_envs, dtype=torch.int32, device=dev)
        self._sp = MnkSelfplay(self._side.data_ptr(), self.pending_resets.data_ptr(), self._episodes.data_ptr(),
                               self.seed & (2**64 - 1), env.env_offset, None)
        self._counter_base: Optional[torch.Tensor] = None
        self._spp = ctypes.byref(self._sp)
        self._opp_active = torch.zeros(self.num_envs, dtype=torch.uint8, device=dev)
        self._steps = 0
        self.next_sides: Optional[torch.Tensor] = None   # i64[N]: sides for envs reset by the NEXT step (else Philox)
        self._side_mirror: Optional[torch.Tensor] = None

    # ------------------------------------------------------------------ reference attributes
    @property
    def agent_side(self) -> torch.Tensor:
        """i64[N], 0 = black, 1 = white (reference :13).  A live, writable mirror of the u8 sides the kernels
        use (like env.boards): created on first access, folded back before every wrapper operation (so in-place
        writes such as ``wrapper.agent_side[:] = 1`` take effect) and refreshed afterwards."""
        if self._side_mirror is None:
            self._side_mirror = self._side.long()
        return self._side_mirror

    @agent_side.setter
    def agent_side(self, value):
        self._side.copy_(torch.as_tensor(value, device=self._dev).to(torch.uint8).expand(self.num_envs))
        if self._side_mirror is not None:
            self._side_mirror.copy_(self._side)

    def _fold_side(self):
        if self._side_mirror is not None:
            self._side.copy_(self._side_mirror)

    def _refresh_side(self):
        if self._side_mirror is not None:
            self._side_mirror.copy_(self._side)

    def set_opponent(self, policy):
        self.opponent_policy = policy
        if hasattr(policy, "counter_base"):
            policy.counter_base = self._counter_base

    @property
    def counter_base(self) -> Optional[torch.Tensor]:
        """Device int64[1] added on the device to the step counters of this wrapper's (and its native
        opponent's) random draws: a CUDA graph that captured steps bakes the counters, bumping this tensor
        between replays makes every replay draw fresh numbers (RolloutCollector.collect(graph=True))."""
        return self._counter_base

    @counter_base.setter
    def counter_base(self, t: Optional[torch.Tensor]):
        self._counter_base = t
        self._sp.counter_base = None if t is None else t.data_ptr()
        if hasattr(self.opponent_policy, "counter_base"):
            self.opponent_policy.counter_base = t

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return torch.cuda.current_stream(self._dev).cuda_stream

    def _new_out(self):
        n = self.num_envs
        return (torch.empty(n, dtype=torch.float32, device=self._dev), torch.empty(n, dtype=torch.bool, device=self._dev))

    def _run(self, actions: Optional[torch.Tensor], forced: Optional[torch.Tensor], reset_all: bool,
             materialise: bool = True):
        env = self.env
        env._fold_mirrors()
        self._fold_side()
        flags = _lib.SP_RESET_ALL if reset_all else 0
        a = None
        if actions is not None:
            a = torch.as_tensor(actions, device=self._dev)
            if a.dtype == torch.int32:
                flags |= _lib.SP_ACTIONS_I32
            else:
                a = a.to(torch.long)
            a = a.contiguous()
            if a.numel() != self.num_envs:
                raise ValueError(f"step: expected {self.num_envs} actions, got {a.numel()}")
        if forced is not None:
            forced = torch.as_tensor(forced, device=self._dev).to(torch.long).expand(self.num_envs).contiguous()
        rewards, terminated = self._new_out()
        obs, mask = env._new_obs() if materialise else (None, None)
        self._steps += 1
        opp = self.opponent_policy
        with torch.cuda.device(self._dev):
            if isinstance(opp, RandomPolicy):
                check(self._L.mnk_selfplay_step_random(env._stp, self._spp, _ptr(a), _ptr(forced), self._steps,
                                                        _ptr(rewards), _ptr(terminated), _ptr(obs), _ptr(mask), flags,
                                                        self._stream()), "mnk_selfplay_step_random")
            else:
                if opp is None:
                    raise RuntimeError("TorchSelfPlayWrapper: set_opponent() has not been called")
                # native policies read the bitboards: no f32 opponent view
                from_bits = callable(getattr(opp, "act_from_env", None)) and getattr(opp, "reads_bitboards", True)
                opp_obs, opp_mask = (None, None) if from_bits else env._new_obs()
                check(self._L.mnk_selfplay_agent(env._stp, self._spp, _ptr(a), _ptr(forced), _ptr(rewards),
                                                  _ptr(terminated), _ptr(self._opp_active), _ptr(opp_obs), _ptr(opp_mask),
                                                  flags, self._stream()), "mnk_selfplay_agent")
                with torch.no_grad():
                    if from_bits:
                        opp_actions = opp.act_from_env(env, self._steps)
                    else:                  # reference :91-94: one positional argument
                        opp_actions = opp.act({"observation": opp_obs, "action_mask": opp_mask})
                oa = torch.as_tensor(opp_actions, device=self._dev)
                oflags = 0
                if oa.dtype == torch.int32:
                    oflags = _lib.SP_ACTIONS_I32
                else:
                    oa = oa.to(torch.long)
                oa = oa.contiguous()
                if oa.numel() != self.num_envs:
                    raise ValueError(f"opponent returned {oa.numel()} actions for a dense batch of {self.num_envs}")
                check(self._L.mnk_selfplay_opponent(env._stp, self._spp, _ptr(oa), _ptr(self._opp_active), _ptr(rewards),
                                                     _ptr(terminated), _ptr(obs), _ptr(mask), oflags, self._stream()),
                      "mnk_selfplay_opponent")
        env._refresh_mirrors()
        self._refresh_side()
        return {"observation": obs, "action_mask": mask}, rewards, terminated

    # ------------------------------------------------------------------ reference methods
    def reset(self, seed=None, options=None, materialise: bool = True):
        """reference :19-30 (``seed`` is ignored there too)."""
        forced = None
        if options and "agent_side" in options:
            forced = options["agent_side"]
        elif self.next_sides is not None:
            forced = self.next_sides
        obs, _, _ = self._run(None, forced, reset_all=True, materialise=materialise)
        self.pending_resets.zero_()     # no-op by construction; mirrors :21
        return obs, {}

    def step(self, actions: torch.Tensor, materialise: bool = True):
        """reference :32-67.  materialise=False skips writing the f32 observation / bool mask (callers that
        read the bitboards, e.g. the tcgen05 forward): obs entries are then None."""
        obs, rewards, terminated = self._run(actions, self.next_sides, reset_all=False, materialise=materialise)
        return obs, rewards, terminated, torch.zeros_like(terminated), {}

    def get_agent_obs(self) -> Dict[str, torch.Tensor]:
        """reference :99-115"""
        self.env._fold_mirrors()
        self._fold_side()
        return self.env._observe_packed(swap=self._side, fix_all_masked=True)

    _get_canonical_obs = get_agent_obs

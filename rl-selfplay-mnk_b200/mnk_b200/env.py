"""TorchVectorMnkEnv on packed bitboards and sm_100a kernels.

Drop-in for the reference's ``src/env/torch_vector_mnk_env.py:7-119``: same constructor,
attributes (``m n k num_envs device max_moves env_indices boards current_player
move_counts``), methods and return tensors (``observation`` f32[N,2,m,n], ``action_mask``
bool[N,m*n], ``rewards`` f32[N], ``dones`` bool[N]).  The state itself lives in HBM as two
guard-strided bitboards per env plus one u32 of (move_count, player) -- see
include/mnk_b200.h -- and every operation is one launch of libmnk_b200.so on the current
torch CUDA stream with no host synchronisation.

``boards``, ``current_player`` and ``move_counts`` are materialised lazily as *live mirrors*
the first time they are read: from then on every operation first folds the mirror back
into the bitboards (so writes such as ``env.boards[0, 0, 0, 0] = 1`` in the reference's
tests take effect) and refreshes it afterwards (so a held reference stays current).  The
training hot path never touches them and pays nothing.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from ._lib import MnkState, check


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


class TorchVectorMnkEnv:
    def __init__(self, m: int, n: int, k: int, num_envs: int, device: str = "cuda", strict: Optional[bool] = None,
                 env_offset: int = 0):
        assert m >= k and n >= k, f"Board ({m}x{n}) is too small for k={k}"   # reference :9
        self.m, self.n, self.k = int(m), int(n), int(k)
        self.num_envs = int(num_envs)
        self.device = device
        self._dev = torch.device(device)
        if self._dev.type != "cuda":
            raise RuntimeError("mnk_b200 runs on sm_100a CUDA devices only (no CPU fallback); "
                               f"got device={device!r}")
        if self._dev.index is None:
            self._dev = torch.device("cuda", torch.cuda.current_device())
        self._L = _lib.lib()
        words = self._L.mnk_state_words(self.m, self.n)
        check(words if words < 0 else 0, "TorchVectorMnkEnv")
        # strict: raise the reference's (dead-code) ValueErrors on illegal / out-of-range moves, :86-104.  Default off --
        # the reference applies such moves silently -- or MNK_B200_STRICT=1 in the environment.
        self.strict = bool(int(os.environ.get("MNK_B200_STRICT", "0"))) if strict is None else bool(strict)
        self._illegal: Optional[torch.Tensor] = None     # i32[2] strict-mode counters written by the step kernels
        self.env_offset = int(env_offset)     # global id of local env 0 (sharded runs)
        self.max_moves = self.m * self.n
        self.env_indices = torch.arange(self.num_envs, device=self._dev)
        # packed state: bits i64[2, words, N] (bit patterns of u64), meta i32[N]
        self._bits = torch.zeros((2, words, self.num_envs), dtype=torch.int64, device=self._dev)
        self._meta = torch.zeros(self.num_envs, dtype=torch.int32, device=self._dev)
        self._st = MnkState(self.m, self.n, self.k, words, self.num_envs, self._bits.data_ptr(), self._meta.data_ptr())
        self._stp = ctypes.byref(self._st)
        self._boards_mirror: Optional[torch.Tensor] = None
        self._player_mirror: Optional[torch.Tensor] = None
        self._count_mirror: Optional[torch.Tensor] = None
        self._dev_actions = self._dev_rd = None     # staging buffers of step_host (non zero-copy transport)
        self._host_views = {}

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return torch.cuda.current_stream(self._dev).cuda_stream

    def _call(self, fn, *args):
        if torch.cuda.current_device() != self._dev.index:
            with torch.cuda.device(self._dev):
                check(fn(self._stp, *args, self._stream()), fn.__name__)
        else:
            check(fn(self._stp, *args, self._stream()), fn.__name__)

    def _fold_mirrors(self):
        """Mirror tensors the caller may have written -> packed state."""
        if self._boards_mirror is not None:
            self._call(self._L.mnk_pack_boards, _ptr(self._boards_mirror))
        if self._player_mirror is not None or self._count_mirror is not None:
            self._call(self._L.mnk_import_meta, _ptr(self._player_mirror), _ptr(self._count_mirror))

    def _refresh_mirrors(self):
        if self._boards_mirror is not None:
            self._call(self._L.mnk_unpack_boards, _ptr(self._boards_mirror))
        if self._player_mirror is not None or self._count_mirror is not None:
            self._call(self._L.mnk_export_meta, _ptr(self._player_mirror), _ptr(self._count_mirror))

    # ------------------------------------------------------------------ reference attributes
    @property
    def boards(self) -> torch.Tensor:
        """f32[N,2,m,n] one-hot planes (reference :17); a live, writable mirror."""
        if self._boards_mirror is None:
            self._boards_mirror = torch.empty((self.num_envs, 2, self.m, self.n), dtype=torch.float32, device=self._dev)
            self._call(self._L.mnk_unpack_boards, _ptr(self._boards_mirror))
        return self._boards_mirror

    @boards.setter
    def boards(self, value: torch.Tensor):
        self.boards.copy_(value)

    @property
    def current_player(self) -> torch.Tensor:
        """i64[N] side to move (reference :18); a live, writable mirror."""
        if self._player_mirror is None:
            self._player_mirror = torch.empty(self.num_envs, dtype=torch.long, device=self._dev)
            self._call(self._L.mnk_export_meta, _ptr(self._player_mirror), None)
        return self._player_mirror

    @current_player.setter
    def current_player(self, value: torch.Tensor):
        self.current_player.copy_(value)

    @property
    def move_counts(self) -> torch.Tensor:
        """i64[N] plies played (reference :19); a live, writable mirror."""
        if self._count_mirror is None:
            self._count_mirror = torch.empty(self.num_envs, dtype=torch.long, device=self._dev)
            self._call(self._L.mnk_export_meta, None, _ptr(self._count_mirror))
        return self._count_mirror

    @move_counts.setter
    def move_counts(self, value: torch.Tensor):
        self.move_counts.copy_(value)

    def release_mirrors(self):
        """Drop the lazily created boards / current_player / move_counts mirrors (after folding any
        writes back), returning the env to the zero-overhead packed mode."""
        self._fold_mirrors()
        self._boards_mirror = self._player_mirror = self._count_mirror = None

    # ------------------------------------------------------------------ reference methods
    def _new_obs(self) -> Tuple[torch.Tensor, torch.Tensor]:
        obs = torch.empty((self.num_envs, 2, self.m, self.n), dtype=torch.float32, device=self._dev)
        mask = torch.empty((self.num_envs, self.m * self.n), dtype=torch.bool, device=self._dev)
        return obs, mask

    def reset(self, env_indices: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """reference :34-44"""
        self._fold_mirrors()
        if env_indices is None:
            self._call(self._L.mnk_reset, None, 0)
        else:
            idx = torch.as_tensor(env_indices, device=self._dev).to(torch.long).contiguous()
            if idx.numel() > 0:     # an empty tensor has a NULL data_ptr, which the C ABI reads as "all envs"
                self._call(self._L.mnk_reset, _ptr(idx), idx.numel())
        self._refresh_mirrors()
        return self._observe_packed()

    def _observe_packed(self, swap: Optional[torch.Tensor] = None, fix_all_masked: bool = False):
        obs, mask = self._new_obs()
        self._call(self._L.mnk_observe, _ptr(obs), _ptr(mask), _ptr(swap), int(fix_all_masked))
        return {"observation": obs, "action_mask": mask}

    def legal_mask(self, fix_all_masked: bool = False) -> torch.Tensor:
        """bool[N, m*n] legal-cell mask only (no f32 observation): 1/9 of observe()'s bytes."""
        self._fold_mirrors()
        mask = torch.empty((self.num_envs, self.m * self.n), dtype=torch.bool, device=self._dev)
        self._call(self._L.mnk_observe, None, _ptr(mask), None, int(fix_all_masked))
        return mask

    def observe(self) -> Dict[str, torch.Tensor]:
        """reference :46-53 -- fresh tensors every call (callers mutate them)."""
        self._fold_mirrors()
        return self._observe_packed()

    def step(self, actions: torch.Tensor):
        """reference :55-58"""
        return self._step(actions, None)

    def step_subset(self, actions: torch.Tensor, active_indices: torch.Tensor):
        """reference :60-84 -- rewards / dones are full-size [N], zero for unlisted envs."""
        return self._step(actions, active_indices)

    def _prep_actions(self, actions: torch.Tensor) -> Tuple[torch.Tensor, int]:
        a = torch.as_tensor(actions, device=self._dev)
        if a.dtype == torch.int32:
            return a.contiguous(), _lib.STEP_ACTIONS_I32
        return a.to(torch.long).contiguous(), 0

    def _step(self, actions, active_indices, autoreset: bool = False, materialise: bool = True,
              out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
        self._fold_mirrors()
        a, flags = self._prep_actions(actions)
        idx = None
        if active_indices is not None:
            idx = torch.as_tensor(active_indices, device=self._dev).to(torch.long).contiguous()
            if a.numel() != idx.numel():
                raise ValueError(f"step_subset: {a.numel()} actions for {idx.numel()} indices")
        elif a.numel() != self.num_envs:
            raise ValueError(f"step: expected {self.num_envs} actions, got {a.numel()}")
        illegal = None
        if self.strict:
            if self._illegal is None:
                self._illegal = torch.zeros(2, dtype=torch.int32, device=self._dev)
            illegal = self._illegal
            illegal.zero_()
        if autoreset:
            flags |= _lib.STEP_AUTORESET
        rewards = torch.empty(self.num_envs, dtype=torch.float32, device=self._dev)
        dones = torch.empty(self.num_envs, dtype=torch.bool, device=self._dev)
        obs = mask = None
        if out is not None:
            obs, mask = out
        elif materialise:
            obs, mask = self._new_obs()
        if idx is not None and idx.numel() == 0:
            # the reference raises on an empty subset (view(0, -1) in _check_wins); here it is a no-op
            rewards.zero_(), dones.zero_()
            if obs is not None or mask is not None:
                self._call(self._L.mnk_observe, _ptr(obs), _ptr(mask), None, 0)
            return {"observation": obs, "action_mask": mask}, rewards, dones
        self._call(self._L.mnk_step, _ptr(a), _ptr(idx), a.numel(), _ptr(rewards), _ptr(dones), _ptr(obs), _ptr(mask),
                   _ptr(illegal), flags)
        self._refresh_mirrors()
        if illegal is not None:
            self._raise_if_illegal(a, idx)
        return {"observation": obs, "action_mask": mask}, rewards, dones

    def _raise_if_illegal(self, a: torch.Tensor, idx: Optional[torch.Tensor]):
        """Opt-in legality check with the reference's (dead-code) messages, :86-104.  The step kernel counts
        moves onto occupied / out-of-range cells into a device flag (include/mnk_b200.h, `illegal`); strict mode
        costs ONE device->host read of that flag per step.  The offending moves have been applied exactly as in
        the default mode (occupied cell: stone written over; out of range: no stone) when the error is raised."""
        count, key = self._illegal.tolist()
        if count == 0:
            return
        env_id = 0x7FFFFFFF - key                           # smallest offending env index
        pos = env_id if idx is None else int(torch.nonzero(idx == env_id)[0])
        val = int(a[pos])
        cells = self.m * self.n
        if val < 0 or val >= cells:
            raise ValueError(f"Action out of bounds! Env received {val}, expected [0, {cells - 1}]")
        raise ValueError(f"Illegal Move: Env {env_id} tried to play in occupied cell.")

    # ------------------------------------------------------------------ extensions (not in the reference)
    def step_autoreset(self, actions: torch.Tensor, materialise: bool = True, out=None):
        """env.step(actions) followed by env.reset(dones.nonzero()) as ONE launch: the returned
        observation is the reference's step() observation (terminal boards included); finished envs
        are empty boards from the next call on."""
        return self._step(actions, None, autoreset=True, materialise=materialise, out=out)

    def random_legal_actions(self, seed: int, counter: int, deterministic: bool = False,
                             out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """RandomPolicy.act (src/selfplay/policy.py:13-29) straight from the bitboards."""
        self._fold_mirrors()
        if out is None:
            out = torch.empty(self.num_envs, dtype=torch.long, device=self._dev)
        self._call(self._L.mnk_random_legal, seed & (2**64 - 1), counter, self.env_offset, int(deterministic), _ptr(out))
        return out

    def step_host(self, host_actions: torch.Tensor, host_out: torch.Tensor, autoreset: bool = False,
                  out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, zero_copy: bool = False, sync: bool = True):
        """End-to-end step for callers holding HOST buffers: pinned int64 actions in, pinned
        rewards/dones bytes out (5*N bytes: f32 rewards then u8 dones), one H2D + launch + D2H +
        stream sync inside libmnk_b200 (mnk_step_host).  Observation / mask stay on the device.
        zero_copy=True: both host buffers must be pinned (torch .pin_memory()); the kernel then reads the
        actions and writes rewards / dones over PCIe itself -- one launch + one sync, no staging copies.
        sync=False enqueues without waiting: synchronise the stream the call ran on before reading the returned
        host views (a host loop can keep two env groups in flight on two streams)."""
        self._fold_mirrors()
        n = self.num_envs
        flags = _lib.STEP_AUTORESET if autoreset else 0
        if zero_copy:
            if not (host_actions.is_pinned() and host_out.is_pinned()):
                raise ValueError("step_host(zero_copy=True) needs pinned host tensors")
            flags |= _lib.STEP_ZEROCOPY
            dev_actions = dev_rd = None
        else:
            if self._dev_actions is None:
                self._dev_actions = torch.empty(n, dtype=torch.long, device=self._dev)
                self._dev_rd = torch.empty(5 * n, dtype=torch.uint8, device=self._dev)
            dev_actions, dev_rd = self._dev_actions.data_ptr(), self._dev_rd.data_ptr()
        if host_actions.dtype == torch.int32:
            flags |= _lib.STEP_ACTIONS_I32
        if not sync:
            flags |= _lib.STEP_NOSYNC
        obs, mask = out if out is not None else self._new_obs()
        self._call(self._L.mnk_step_host, host_actions.data_ptr(), dev_actions, dev_rd, host_out.data_ptr(), _ptr(obs),
                   _ptr(mask), flags)
        self._refresh_mirrors()
        views = self._host_views.get(host_out.data_ptr())
        if views is None:       # typed views of the caller's byte buffer, built once per buffer
            views = (host_out[:4 * n].view(torch.float32), host_out[4 * n:5 * n].view(torch.bool))
            self._host_views = {host_out.data_ptr(): views}
        return {"observation": obs, "action_mask": mask}, views[0], views[1]

    def step_host_loop(self, host_actions: torch.Tensor, host_out: torch.Tensor, slab_steps: int = 4, autoreset: bool = False,
                       ring: Optional[Tuple[list, list]] = None, host_obs: Optional[torch.Tensor] = None,
                       host_mask: Optional[torch.Tensor] = None, buffers: int = 3):
        """K dense steps in ONE call of the C ABI (mnk_step_host_loop): `host_actions` pinned int64|int32 [K, N],
        `host_out` pinned uint8 [K, 5*N] (per step f32 rewards then bool dones).  The library pipelines slabs of
        `slab_steps` steps -- H2D of the next slab | the step kernels | D2H of the previous slab -- with one copy per
        slab and direction and one host wait per slab.  `ring` = (obs_list, mask_list) of device tensors the steps
        materialise their observation / mask into (step t -> slot t % len); None = packed mode.  `host_obs` /
        `host_mask` (pinned f32 [K, N, 2, m, n] / bool [K, N, m*n]) additionally bring every step's views to the
        host.  Returns (rewards f32 [K, N], dones bool [K, N]) as views of `host_out`."""
        self._fold_mirrors()
        n = self.num_envs
        if host_actions.dim() != 2 or host_actions.shape[1] != n or not host_actions.is_contiguous():
            raise ValueError(f"step_host_loop: host_actions must be a contiguous [K, {n}] tensor")
        steps = host_actions.shape[0]
        if host_out.dtype != torch.uint8 or host_out.numel() < steps * 5 * n or not host_out.is_contiguous():
            raise ValueError("step_host_loop: host_out must be a contiguous uint8 tensor of at least K * 5 * N bytes")
        if not (host_actions.is_pinned() and host_out.is_pinned()):
            raise ValueError("step_host_loop needs pinned host tensors")
        flags = _lib.STEP_AUTORESET if autoreset else 0
        if host_actions.dtype == torch.int32:
            flags |= _lib.STEP_ACTIONS_I32
        elif host_actions.dtype != torch.long:
            raise ValueError("step_host_loop: actions must be int64 or int32")
        slab_steps = max(1, min(int(slab_steps), steps)) if steps else 1
        buffers = max(2, min(int(buffers), 4))
        if ring is not None and (host_obs is not None or host_mask is not None):
            buffers = max(2, min(buffers, len(ring[0]) // slab_steps))       # the views of every slab in flight stay in the ring
        key = (slab_steps, host_actions.dtype, buffers)
        if getattr(self, "_loop_key", None) != key:      # device slabs (`buffers` deep) and the stream / event pipe
            self._loop_dev_actions = torch.empty((buffers, slab_steps, n), dtype=host_actions.dtype, device=self._dev)
            self._loop_dev_rd = torch.empty((buffers, slab_steps, 5 * n), dtype=torch.uint8, device=self._dev)
            self._loop_key = key
        if getattr(self, "_loop_pipe", None) is None:
            pipe = ctypes.c_void_p()
            with torch.cuda.device(self._dev):
                check(self._L.mnk_host_pipe_create(ctypes.byref(pipe)), "mnk_host_pipe_create")
            self._loop_pipe = pipe
        job = _lib.MnkHostLoop()
        job.host_actions, job.host_rd = host_actions.data_ptr(), host_out.data_ptr()
        job.host_obs = None if host_obs is None else host_obs.data_ptr()
        job.host_mask = None if host_mask is None else host_mask.data_ptr()
        for t_ in (host_obs, host_mask):
            if t_ is not None and not (t_.is_pinned() and t_.is_contiguous()):
                raise ValueError("step_host_loop: host_obs / host_mask must be pinned and contiguous")
        job.dev_actions, job.dev_rd = self._loop_dev_actions.data_ptr(), self._loop_dev_rd.data_ptr()
        if ring is not None:
            obs_list, mask_list = ring
            cnt = len(obs_list)
            arr_o = (ctypes.c_void_p * cnt)(*[o.data_ptr() for o in obs_list])
            arr_m = (ctypes.c_void_p * cnt)(*[m_.data_ptr() for m_ in mask_list])
            job.obs_ring, job.mask_ring, job.ring = arr_o, arr_m, cnt
        else:
            job.ring = 0
        job.steps, job.slab_steps, job.buffers = steps, slab_steps, buffers
        self._call(self._L.mnk_step_host_loop, ctypes.byref(job), self._loop_pipe, flags)
        self._refresh_mirrors()
        out = host_out.reshape(-1)[: steps * 5 * n].view(steps, 5 * n)
        return out[:, : 4 * n].view(torch.float32), out[:, 4 * n:].view(torch.bool)

    def __del__(self):
        pipe = getattr(self, "_loop_pipe", None)
        if pipe is not None:
            try:
                self._L.mnk_host_pipe_destroy(pipe)
            except Exception:
                pass

    def state_checksum(self) -> int:
        """Order-sensitive 64-bit digest of the packed state (used by bench.py / tests)."""
        self._fold_mirrors()
        w = torch.arange(1, self._bits.numel() + 1, device=self._dev, dtype=torch.long)
        h = (self._bits.reshape(-1) * w).sum() + (self._meta.long() * (w[: self.num_envs] * 31 + 7)).sum()
        return int(h.item()) & (2**64 - 1)

"""CPU ORACLE (numpy + the C restatement in mnk_oracle.c) for the MNK hot path.

TEST INFRASTRUCTURE ONLY.  Restates, on the CPU and in a representation that shares
nothing with the CUDA path (one byte per cell, python control flow), what the
reference does on the path named in BASELINE.json:

  * OracleEnv       <- src/env/torch_vector_mnk_env.py:7-119
  * OracleWrapper   <- src/selfplay/torch_self_play_wrapper.py:6-115
  * masked_log_softmax / masked_argmax / first_legal <- src/alg/architectures/resnet.py:84-94,
                       src/selfplay/policy.py:13-29
  * philox4x32 / random_legal_actions / side_draw: the counter-based RNG contract of the
    CUDA path (NEW, no reference counterpart: the reference draws from torch's global
    generator, which no kernel can reproduce; SURVEY.md section 7 "RNG").

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  Parity status: PINNED against tests/golden/*.npz (generated
from the unmodified reference by oracle/gen_golden.py).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Callable, Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

PLAYER_BLACK = 0  # reference: src/env/constants.py:1-2
PLAYER_WHITE = 1


def build(force: bool = False) -> str:
    """Compile mnk_oracle.c -> liboracle.so with gcc (no GPU, no torch)."""
    src = os.path.join(_HERE, "mnk_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-fvisibility=hidden", "-o", _LIB_PATH, src]
        subprocess.run(cmd, check=True)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        p, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
        L.orc_reset.argtypes = [p, p, p, i32, i32, i64, p, i64]
        L.orc_reset.restype = None
        L.orc_observe.argtypes = [p, i32, i32, i64, p, p]
        L.orc_observe.restype = None
        L.orc_plane_has_line.argtypes = [p, i32, i32, i32]
        L.orc_plane_has_line.restype = i32
        L.orc_step_subset.argtypes = [p, p, p, i32, i32, i32, i64, p, p, i64, p, p]
        L.orc_step_subset.restype = None
        L.orc_nth_legal.argtypes = [p, i32, i32, i64, i64]
        L.orc_nth_legal.restype = i64
        L.orc_count_legal.argtypes = [p, i32, i32, i64]
        L.orc_count_legal.restype = i64
        _lib = L
    return _lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class OracleEnv:
    """Restatement of TorchVectorMnkEnv (src/env/torch_vector_mnk_env.py:7-84)."""

    def __init__(self, m: int, n: int, k: int, num_envs: int):
        assert m >= k and n >= k, f"Board ({m}x{n}) is too small for k={k}"  # :9
        self.m, self.n, self.k, self.num_envs = m, n, k, num_envs
        self.boards = np.zeros((num_envs, 2, m, n), dtype=np.uint8)   # f32 0/1 in the reference (:17)
        self.current_player = np.zeros(num_envs, dtype=np.int64)     # :18
        self.move_counts = np.zeros(num_envs, dtype=np.int64)        # :19
        self.max_moves = m * n                                       # :20
        self.env_indices = np.arange(num_envs, dtype=np.int64)       # :22

    # :34-44
    def reset(self, env_indices: Optional[np.ndarray] = None) -> Dict[str, np.ndarray]:
        idx = None if env_indices is None else np.ascontiguousarray(env_indices, dtype=np.int64)
        lib().orc_reset(_ptr(self.boards), _ptr(self.current_player), _ptr(self.move_counts),
                        self.m, self.n, self.num_envs, _ptr(idx), 0 if idx is None else idx.size)
        return self.observe()

    # :46-53
    def observe(self) -> Dict[str, np.ndarray]:
        obs = np.empty((self.num_envs, 2, self.m, self.n), dtype=np.float32)
        mask = np.empty((self.num_envs, self.m * self.n), dtype=np.uint8)
        lib().orc_observe(_ptr(self.boards), self.m, self.n, self.num_envs, _ptr(obs), _ptr(mask))
        return {"observation": obs, "action_mask": mask.astype(bool)}

    # :55-58
    def step(self, actions: np.ndarray):
        return self.step_subset(actions, self.env_indices)

    # :60-84
    def step_subset(self, actions: np.ndarray, active_indices: np.ndarray):
        a = np.ascontiguousarray(actions, dtype=np.int64)
        idx = np.ascontiguousarray(active_indices, dtype=np.int64)
        assert a.shape == idx.shape
        rewards = np.empty(self.num_envs, dtype=np.float32)
        dones = np.empty(self.num_envs, dtype=np.uint8)
        lib().orc_step_subset(_ptr(self.boards), _ptr(self.current_player), _ptr(self.move_counts),
                              self.m, self.n, self.k, self.num_envs, _ptr(a), _ptr(idx), idx.size,
                              _ptr(rewards), _ptr(dones))
        return self.observe(), rewards, dones.astype(bool)


PolicyFn = Callable[[Dict[str, np.ndarray]], np.ndarray]


class OracleWrapper:
    """Restatement of TorchSelfPlayWrapper (src/selfplay/torch_self_play_wrapper.py:6-115).

    `opponent` is a callable obs_dict -> int64[B] (Policy.act with one positional
    argument, :92-94).  `side_fn(env_indices) -> int64[len]` replaces the reference's
    torch.randint(0, 2, ...) draws (:26, :43-45) so that tests can inject sides."""

    def __init__(self, env: OracleEnv, side_fn: Optional[Callable[[np.ndarray], np.ndarray]] = None):
        self.env = env
        self.num_envs = env.num_envs
        self.opponent: Optional[PolicyFn] = None
        self.agent_side = np.zeros(self.num_envs, dtype=np.int64)       # :13
        self.pending_resets = np.zeros(self.num_envs, dtype=bool)      # :14
        self.side_fn = side_fn or (lambda idx: np.random.randint(0, 2, size=len(idx)).astype(np.int64))

    def set_opponent(self, policy: PolicyFn):                           # :16-17
        self.opponent = policy

    def reset(self, seed=None, options=None):                           # :19-30
        self.env.reset()
        self.pending_resets[:] = False
        if options and "agent_side" in options:
            self.agent_side[:] = np.asarray(options["agent_side"], dtype=np.int64)
        else:
            self.agent_side = np.asarray(self.side_fn(np.arange(self.num_envs)), dtype=np.int64)
        self._opponent_move_if_needed(np.arange(self.num_envs))
        return self._get_canonical_obs(), {}

    def step(self, actions: np.ndarray):                                # :32-67
        actions = np.asarray(actions, dtype=np.int64)
        reset_mask = self.pending_resets.copy()
        play_mask = ~reset_mask
        rewards = np.zeros(self.num_envs, dtype=np.float32)
        terminated = np.zeros(self.num_envs, dtype=bool)

        if reset_mask.any():                                            # :39-46
            reset_idxs = np.nonzero(reset_mask)[0]
            self.env.reset(reset_idxs)
            self.agent_side[reset_idxs] = np.asarray(self.side_fn(reset_idxs), dtype=np.int64)
            self._opponent_move_if_needed(reset_idxs)

        if play_mask.any():                                             # :48-63
            play_idxs = np.nonzero(play_mask)[0]
            _, r_ag, t_ag = self.env.step_subset(actions[play_idxs], play_idxs)
            rewards[play_idxs] = r_ag[play_idxs]
            terminated[play_idxs] = t_ag[play_idxs]
            still = play_idxs[~terminated[play_idxs]]
            if len(still) > 0:
                opp_r, opp_t = self._opponent_move_if_needed(still)
                if opp_r is not None:
                    rewards[still] -= opp_r[still]
                    terminated[still] = opp_t[still]

        self.pending_resets = terminated.copy()                         # :65
        return self._get_canonical_obs(), rewards, terminated, np.zeros_like(terminated), {}

    def _opponent_move_if_needed(self, env_idxs: np.ndarray):           # :69-97
        if len(env_idxs) == 0:
            return None, None
        opp_turn = self.env.current_player[env_idxs] != self.agent_side[env_idxs]
        if not opp_turn.any():
            return None, None
        active = env_idxs[opp_turn]
        self.last_active = active          # test hook: which envs the opponent is answering
        full = self.env.observe()
        obs_subset = full["observation"][active].copy()
        mask_subset = full["action_mask"][active].copy()
        opp_is_white = self.env.current_player[active] == PLAYER_WHITE
        if opp_is_white.any():
            obs_subset[opp_is_white] = obs_subset[opp_is_white][:, ::-1]   # flip channels (:89)
        opp_actions = np.asarray(self.opponent({"observation": obs_subset, "action_mask": mask_subset}),
                                 dtype=np.int64)
        _, r, t = self.env.step_subset(opp_actions, active)
        return r, t

    def _get_canonical_obs(self):                                       # :99-112
        raw = self.env.observe()
        obs = raw["observation"].copy()
        mask = raw["action_mask"]
        white = self.agent_side == PLAYER_WHITE
        if white.any():
            obs[white] = obs[white][:, ::-1]
        invalid = mask.sum(axis=1) == 0
        if invalid.any():
            mask[invalid, 0] = True
        return {"observation": obs, "action_mask": mask}

    get_agent_obs = _get_canonical_obs                                  # :114-115


# ----------------------------------------------------------------------------
# masking / sampling arithmetic (resnet.py:84-94, policy.py:13-29, ppo.py:99-100)
# ----------------------------------------------------------------------------

def masked_log_softmax(logits: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """Categorical(logits=where(mask, logits, -inf)).logits; all-masked rows -> uniform."""
    lg = np.where(mask.astype(bool), logits.astype(np.float64), -np.inf)
    all_masked = np.max(lg, axis=1, keepdims=True) == -np.inf
    lg = np.where(all_masked, 0.0, lg)
    mx = lg.max(axis=1, keepdims=True)
    lse = mx + np.log(np.exp(lg - mx).sum(axis=1, keepdims=True))
    return lg - lse


def first_legal(mask: np.ndarray) -> np.ndarray:
    """RandomPolicy.act(deterministic=True): argmax of the (possibly all-zero) mask (policy.py:26-27)."""
    return np.argmax(mask.astype(np.float32), axis=1).astype(np.int64)


# ----------------------------------------------------------------------------
# counter-based RNG contract of the CUDA path (Philox4x32-10)
# ----------------------------------------------------------------------------
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
STREAM_ACTION, STREAM_SIDE, STREAM_SAMPLE, STREAM_OPPONENT = 0, 1, 2, 3


def philox4x32(c0, c1, c2, c3, k0: int, k1: int, rounds: int = 10):
    """Vectorised Philox4x32-10.  c* are uint32 arrays (broadcastable), k* python ints."""
    c0, c1, c2, c3 = [np.asarray(x, dtype=np.uint64) & np.uint64(0xFFFFFFFF)
                      for x in np.broadcast_arrays(c0, c1, c2, c3)]
    k0 &= 0xFFFFFFFF
    k1 &= 0xFFFFFFFF
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(rounds):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)) & mask, lo1, (hi0 ^ c3 ^ np.uint64(k1)) & mask, lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return [x.astype(np.uint32) for x in (c0, c1, c2, c3)]


def _draw_u32(seed: int, global_env_ids: np.ndarray, counter, stream: int) -> np.ndarray:
    gid = np.asarray(global_env_ids, dtype=np.uint64)
    ctr = np.asarray(counter, dtype=np.uint64)
    # 64-bit counter: low word = Philox counter word, high word folded into the key (csrc/mnk_device.cuh::mnk_philox)
    hi = int(ctr.max() >> np.uint64(32)) if ctr.size else 0
    assert ctr.ndim == 0 or hi == 0, "per-row counters must fit 32 bits"
    out = philox4x32(gid & np.uint64(0xFFFFFFFF), gid >> np.uint64(32), ctr & np.uint64(0xFFFFFFFF),
                     np.uint64(stream), seed & 0xFFFFFFFF, ((seed >> 32) ^ hi) & 0xFFFFFFFF)
    return out[0]


def random_legal_actions(mask: np.ndarray, seed: int, counter: int, env_offset: int = 0, env_ids=None,
                         stream: int = STREAM_ACTION) -> np.ndarray:
    """Contract of mnk_random_legal / the fused random opponent: j = mulhi32(philox, #legal); pick the
    j-th legal cell in ascending cell order; rows with no legal cell draw uniformly from all cells
    (the reference's RandomPolicy adds 1e-8 to every entry of such rows, policy.py:21-24).
    Row i belongs to global env  env_offset + (env_ids[i] if given else i)."""
    mask = mask.astype(bool)
    n_env, cells = mask.shape
    ids = np.arange(n_env) if env_ids is None else np.asarray(env_ids)
    x = _draw_u32(seed, env_offset + ids, counter, stream).astype(np.uint64)
    cnt = mask.sum(axis=1).astype(np.uint64)
    eff = np.where(cnt == 0, np.uint64(cells), cnt)
    j = ((x * eff) >> np.uint64(32)).astype(np.int64)
    order = np.cumsum(mask, axis=1) - 1            # rank of each legal cell
    hit = mask & (order == j[:, None])
    picked = np.argmax(hit, axis=1).astype(np.int64)
    return np.where(cnt == 0, j, picked)


def side_draw(seed: int, global_env_ids: np.ndarray, episode_counter) -> np.ndarray:
    """Contract of the wrapper's on-device side assignment: lowest Philox bit."""
    return (_draw_u32(seed, global_env_ids, episode_counter, STREAM_SIDE) & np.uint32(1)).astype(np.int64)


def gae(rewards, values, dones, last_values, gamma, lam):
    """RolloutBuffer.compute_advantages_and_returns (src/alg/rollout_buffer.py:60-80) in float32,
    same operation order as the reference's tensor expressions."""
    f = np.float32
    rewards, values = rewards.astype(f), values.astype(f)
    steps, n = rewards.shape
    adv = np.zeros((steps, n), dtype=f)
    last_gae = np.zeros(n, dtype=f)
    # gamma * gae_lambda is a product of two Python floats (double) before it meets an fp32 tensor (:76)
    gl = f(float(gamma) * float(lam))
    gamma = f(gamma)
    for t in reversed(range(steps)):
        next_values = last_values.astype(f) if t == steps - 1 else values[t + 1]
        nnt = f(1.0) - dones[t].astype(f)
        delta = rewards[t] + gamma * next_values * nnt - values[t]
        last_gae = delta + gl * nnt * last_gae
        adv[t] = last_gae
    return adv, adv + values

"""Rollout into PPO buffers on packed bitboards.

``RolloutBuffer`` keeps the reference's interface (src/alg/rollout_buffer.py:5-113: ``add``,
``compute_advantages_and_returns``, ``get_data_loader``, ``reset``, the per-step arrays) but
stores each observation as the agent's two canonical planes in bitboard form (32 B per env-step at
9x9 instead of 648 B of f32 + 81 B of mask); f32 observations and bool masks are re-materialised
only for the rows of a minibatch (``mnk_rollout_gather``).  BASELINE cfg3 (512 steps x 262,144
envs) is ~102 GB in the reference layout and ~9.7 GB here.

``RolloutCollector`` is the rollout section of ``PPOAgent.learn`` (src/alg/ppo.py:78-133): same
per-step order (forward, sample, log_prob, env step, buffer add, episode accounting), the
observation carried across calls, ``fps = n_steps * num_envs / rollout_time`` -- with episode
statistics accumulated on the device and read back once per rollout instead of ``.tolist()`` every
step, and reduced over ranks with a single NCCL all-reduce when sharded.
"""
from __future__ import annotations

import ctypes
import time
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from ._lib import MnkState, check
from .sampling import MaskedCategorical, fresh_seed, masked_sample

_STATIC_K = {(3, 3): 3, (9, 9): 5, (13, 13): 5, (15, 15): 5, (19, 19): 5}


def _ptr(t):
    return None if t is None else t.data_ptr()


class RolloutBuffer:
    def __init__(self, n_steps, num_envs, obs_shape, action_dim, device="cuda", k: Optional[int] = None):
        self.n_steps, self.num_envs = int(n_steps), int(num_envs)
        self.obs_shape = tuple(obs_shape)
        self.action_dim = int(action_dim)
        self.device = device
        self._dev = torch.device(device)
        if self._dev.type != "cuda":
            raise RuntimeError("mnk_b200.RolloutBuffer: CUDA only (no CPU fallback)")
        _, self.m, self.n = self.obs_shape
        assert self.m * self.n == self.action_dim
        self.k = int(k) if k is not None else _STATIC_K.get((self.m, self.n), 1)
        self._L = _lib.lib()
        self.words = self._L.mnk_state_words(self.m, self.n)
        check(self.words if self.words < 0 else 0, "RolloutBuffer")
        self._dummy_meta = torch.zeros(self.num_envs, dtype=torch.int32, device=self._dev)
        self._allocated = False
        self.reset()

    def _stream(self):
        return torch.cuda.current_stream(self._dev).cuda_stream

    def reset(self):
        """reference :13-45 re-allocates every array each iteration; here they are allocated once and
        re-zeroed (returns / advantages are fully overwritten by compute_advantages_and_returns)."""
        T, N, dev = self.n_steps, self.num_envs, self._dev
        if not self._allocated:
            self.packed_obs = torch.zeros((T, 2, self.words, N), dtype=torch.int64, device=dev)
            self.actions = torch.zeros((T, N), dtype=torch.long, device=dev)
            self.log_probs = torch.zeros((T, N), dtype=torch.float32, device=dev)
            self.rewards = torch.zeros((T, N), dtype=torch.float32, device=dev)
            self.values = torch.zeros((T, N), dtype=torch.float32, device=dev)
            self.returns = torch.zeros((T, N), dtype=torch.float32, device=dev)
            self.advantages = torch.zeros((T, N), dtype=torch.float32, device=dev)
            self.dones = torch.zeros((T, N), dtype=torch.bool, device=dev)
            self._allocated = True
        else:
            for a in (self.packed_obs, self.actions, self.log_probs, self.rewards, self.values, self.returns,
                      self.advantages, self.dones):
                a.zero_()
        self.ptr = 0

    def _slot_state(self, t: int) -> MnkState:
        return MnkState(self.m, self.n, self.k, self.words, self.num_envs, self.packed_obs[t].data_ptr(),
                        self._dummy_meta.data_ptr())

    def _add_scalars(self, action, reward, value, log_prob, done):
        t = self.ptr
        self.actions[t].copy_(action)
        self.rewards[t].copy_(reward)
        self.values[t].copy_(value.view(-1))
        self.log_probs[t].copy_(log_prob)
        self.dones[t].copy_(done)
        self.ptr += 1

    def add(self, obs, action, reward, value, log_prob, done, action_mask=None):
        """reference :47-58.  `obs` is the f32 canonical observation; it is packed on the way in.
        `action_mask` is accepted for signature compatibility; masks are derived from the planes."""
        if self.ptr >= self.n_steps:
            raise IndexError("Buffer was full.")
        st = self._slot_state(self.ptr)
        obs = obs.contiguous()
        with torch.cuda.device(self._dev):
            check(self._L.mnk_pack_boards(ctypes.byref(st), obs.data_ptr(), self._stream()), "mnk_pack_boards")
        self._add_scalars(action, reward, value, log_prob, done)

    def store_obs_from(self, wrapper):
        """Write the agent's canonical planes of wrapper.env's CURRENT state into the next slot straight
        from the bitboards (no f32 round trip).  Call before the step, then add_transition()."""
        if self.ptr >= self.n_steps:
            raise IndexError("Buffer was full.")
        env = wrapper.env
        env._fold_mirrors()
        with torch.cuda.device(self._dev):
            check(self._L.mnk_rollout_store_obs(env._stp, wrapper._side.data_ptr(), self.packed_obs[self.ptr].data_ptr(),
                                                 self._stream()), "mnk_rollout_store_obs")

    def add_transition(self, action, reward, value, log_prob, done):
        self._add_scalars(action, reward, value, log_prob, done)

    def compute_advantages_and_returns(self, last_values, gamma=0.99, gae_lambda=0.95):
        """reference :60-80 as one kernel (thread per env, reverse scan), bit-identical in fp32."""
        steps = self.ptr
        last_values = last_values.reshape(self.num_envs).float().contiguous()
        with torch.cuda.device(self._dev):
            check(self._L.mnk_gae(_ptr(self.rewards), _ptr(self.values), _ptr(self.dones), _ptr(last_values), steps,
                                   self.num_envs, gamma, gae_lambda, _ptr(self.advantages), _ptr(self.returns),
                                   self._stream()), "mnk_gae")

    def gather_obs(self, flat_index: Optional[torch.Tensor], count: Optional[int] = None):
        """f32[B,2,m,n] observation and bool[B,m*n] mask of samples `flat_index` of the [steps*N] rollout."""
        if flat_index is not None:
            flat_index = flat_index.to(device=self._dev, dtype=torch.long).contiguous()
            count = flat_index.numel()
        obs = torch.empty((count, *self.obs_shape), dtype=torch.float32, device=self._dev)
        mask = torch.empty((count, self.action_dim), dtype=torch.bool, device=self._dev)
        with torch.cuda.device(self._dev):
            check(self._L.mnk_rollout_gather(self.m, self.n, self.k, _ptr(self.packed_obs), self.num_envs,
                                              _ptr(flat_index), count, _ptr(obs), _ptr(mask), self._stream()),
                  "mnk_rollout_gather")
        return obs, mask

    @property
    def observations(self) -> torch.Tensor:
        """The reference's f32[T,N,2,m,n] array, materialised on demand (tests / small runs)."""
        obs, _ = self.gather_obs(None, self.n_steps * self.num_envs)
        return obs.view(self.n_steps, self.num_envs, *self.obs_shape)

    @property
    def action_masks(self) -> torch.Tensor:
        _, mask = self.gather_obs(None, self.n_steps * self.num_envs)
        return mask.view(self.n_steps, self.num_envs, self.action_dim)

    def get_data_loader(self, batch_size, normalize_advantages=True):
        """reference :82-113: same tuple order (obs, actions, log_probs, returns, advantages, masks, values)."""
        steps = self.ptr
        num_samples = steps * self.num_envs
        b_actions = self.actions[:steps].view(num_samples)
        b_log_probs = self.log_probs[:steps].view(num_samples)
        b_returns = self.returns[:steps].view(num_samples)
        b_advantages = self.advantages[:steps].view(num_samples)
        b_values = self.values[:steps].view(num_samples)
        if normalize_advantages:
            b_advantages = (b_advantages - b_advantages.mean()) / (b_advantages.std() + 1e-8)
        indices = torch.randperm(num_samples, device=self._dev)
        for start in range(0, num_samples, batch_size):
            batch_idx = indices[start:start + batch_size]
            obs, masks = self.gather_obs(batch_idx)
            yield (obs, b_actions[batch_idx], b_log_probs[batch_idx], b_returns[batch_idx], b_advantages[batch_idx],
                   masks, b_values[batch_idx])


@dataclass
class RolloutStats:
    mean_reward: float
    mean_length: float
    episodes: float
    wins: float
    losses: float
    draws: float
    fps: float
    rollout_time: float
    agent_steps: int


class RolloutCollector:
    """The rollout half of PPOAgent.learn (src/alg/ppo.py:78-133)."""

    def __init__(self, num_envs: int, device="cuda", seed: Optional[int] = None, row_offset: int = 0, process_group=None,
                 world_size: int = 1):
        self.num_envs = num_envs
        self._dev = torch.device(device)
        self.seed, self.row_offset = (fresh_seed() if seed is None else seed), row_offset
        self.group, self.world_size = process_group, world_size
        self._L = _lib.lib()
        self._last_obs = None
        self._ep_reward = torch.zeros(num_envs, dtype=torch.float32, device=self._dev)
        self._ep_len = torch.zeros(num_envs, dtype=torch.float32, device=self._dev)
        self._calls = 0

    def _stats_kernel(self, rewards, dones, totals):
        with torch.cuda.device(self._dev):                 # ppo.py:110-120 without host reads
            check(self._L.mnk_episode_stats(_ptr(rewards), _ptr(dones), self.num_envs, _ptr(self._ep_reward),
                                             _ptr(self._ep_len), _ptr(totals),
                                             torch.cuda.current_stream(self._dev).cuda_stream), "mnk_episode_stats")

    def _native_step(self, network, vec_env, buffer, totals, counter_base=None):
        """One agent step on the bitboard-fed path: tcgen05 forward, sampler, packed store, fused wrapper step."""
        env = vec_env.env
        with torch.no_grad():
            logits, values = network.forward_env(env, swap=vec_env._side)
            self._calls += 1
            actions, log_probs, _ = masked_sample(logits, env.legal_mask(fix_all_masked=True), seed=self.seed,
                                                  counter=self._calls, row_offset=self.row_offset, counter_base=counter_base)
        buffer.store_obs_from(vec_env)
        _, rewards, terminateds, truncateds, _ = vec_env.step(actions, materialise=False)
        dones = terminateds | truncateds
        buffer.add_transition(actions, rewards, values, log_probs, dones)
        self._stats_kernel(rewards, dones, totals)

    def _collect_graphed(self, network, vec_env, buffer, steps):
        """The whole rollout as ONE CUDA graph, captured on first use and replayed afterwards.  Every launch of
        the path takes its stream from torch and never synchronises, so the python loop can be captured as is;
        the Philox counters baked into the captured launches are offset by a device-resident base that is
        bumped before each replay, so every replay draws fresh numbers.  Removes the per-launch host overhead
        (decisive at small batch: the reference's default is 384 envs)."""
        # The graph bakes raw device pointers (weights, state, buffer slots).  The key holds STRONG references to the
        # captured objects (an id() can be recycled once its object dies) plus their pointer signatures: NativeResNet
        # keeps its weight tensors at stable addresses across refresh() and reports any re-allocation through
        # `pointer_signature()`, so a replay can never read freed or stale weight storage.
        def make_key():
            opp_ = vec_env.opponent_policy
            sig = tuple(o.pointer_signature() if hasattr(o, "pointer_signature") else None
                        for o in (network, getattr(opp_, "net", None)))
            return (network, vec_env, buffer, steps, opp_, sig, buffer.packed_obs.data_ptr(), vec_env.env._bits.data_ptr())

        key = make_key()
        old = getattr(self, "_graph_key", None)
        if old is None or len(old) != len(key) or any(a is not b and a != b for a, b in zip(old, key)):
            with torch.no_grad():                          # one-time work that must not happen under capture
                state = network.running_state() if hasattr(network, "running_state") else None
                network.forward_env(vec_env.env, swap=vec_env._side)       # (scratch allocation, shared-memory opt-in)
                if state is not None:                      # a train-mode BatchNorm forward: it must not count
                    network.restore_running_state(state)
                opp = vec_env.opponent_policy
                if getattr(opp, "net", None) is not None:
                    opp.net.forward_env(vec_env.env)
            key = make_key()                               # (the warm-up may have allocated the train-mode scratch arrays)
            self._graph_base = torch.zeros(1, dtype=torch.int64, device=self._dev)
            self._graph_totals = torch.zeros(6, dtype=torch.float64, device=self._dev)
            vec_env.counter_base = self._graph_base
            buffer.ptr = 0
            torch.cuda.synchronize(self._dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                for _ in range(steps):
                    self._native_step(network, vec_env, buffer, self._graph_totals, counter_base=self._graph_base)
            self._graph, self._graph_key = graph, key
            self._graph_replays = 0
        else:
            self._graph_base += 1 << 24                    # fresh counter range for this replay
            if hasattr(network, "train_forwards"):         # (the capture pass counted the first replay's forwards)
                network.train_forwards += steps
        buffer.ptr = 0
        self._graph_totals.zero_()
        self._graph.replay()
        buffer.ptr = steps
        self._graph_replays += 1
        return self._graph_totals

    def collect(self, network, vec_env, buffer: RolloutBuffer, n_steps: Optional[int] = None, graph: bool = False) -> RolloutStats:
        start = time.time()
        steps = buffer.n_steps if n_steps is None else n_steps
        packed_fast_path = hasattr(vec_env, "_side") and hasattr(buffer, "store_obs_from")
        native = packed_fast_path and hasattr(network, "forward_env")     # tcgen05 forward fed from bitboards
        if self._last_obs is None:                        # ppo.py:81-84: reset once, then carry obs across calls
            self._last_obs, _ = vec_env.reset(materialise=not native) if native else vec_env.reset()
        elif self._last_obs["observation"] is None and not native:       # previous rollout ran on the bitboard-fed path
            self._last_obs = vec_env.get_agent_obs()
        obs = self._last_obs
        totals = torch.zeros(6, dtype=torch.float64, device=self._dev)
        if native and graph:
            totals = self._collect_graphed(network, vec_env, buffer, steps)
            steps_to_run = 0
        else:
            steps_to_run = steps
        for _ in range(steps_to_run):
            if native:
                self._native_step(network, vec_env, buffer, totals)
                continue
            observation, action_mask = obs["observation"], obs["action_mask"]
            with torch.no_grad():                          # ppo.py:97-100
                dist, values = network(observation, action_mask)
                self._calls += 1
                if isinstance(dist, MaskedCategorical):
                    actions, log_probs, _ = masked_sample(dist._raw, dist._mask, seed=self.seed, counter=self._calls,
                                                          row_offset=self.row_offset)
                else:   # a torch Categorical: its logits are masked (-inf) and normalised already
                    actions, log_probs, _ = masked_sample(dist.logits, None, seed=self.seed, counter=self._calls,
                                                          row_offset=self.row_offset)
            if packed_fast_path:
                buffer.store_obs_from(vec_env)             # planes straight from the bitboards, before the step
            next_obs, rewards, terminateds, truncateds, _ = vec_env.step(actions)      # ppo.py:102
            dones = terminateds | truncateds
            if packed_fast_path:
                buffer.add_transition(actions, rewards, values, log_probs, dones)
            else:
                buffer.add(observation, actions, rewards, values, log_probs, dones, action_mask)   # ppo.py:106-108
            self._stats_kernel(rewards, dones, totals)
            obs = next_obs
        self._last_obs = obs if not native else {"observation": None, "action_mask": None}
        if self.world_size > 1:                            # the only collective of the rollout path
            import torch.distributed as dist_mod
            totals = totals.clone()                        # (the graph's accumulator stays rank-local)
            dist_mod.all_reduce(totals, op=dist_mod.ReduceOp.SUM, group=self.group)
        # one device->host read per rollout; it also carries the tcgen05 kernels' "internal barrier wait timed out"
        # flags, so a poisoned forward cannot silently fill the buffer with garbage features
        flags = [o._err for o in (network, getattr(vec_env.opponent_policy, "net", None)) if hasattr(o, "_err")]
        tot = torch.cat([totals] + [f.double() for f in flags]).tolist()
        if any(v != 0.0 for v in tot[6:]):
            who = ["agent network", "opponent network"]
            codes = ", ".join(f"{w} flag {int(v):#x}" for w, v in zip(who, tot[6:]))
            raise RuntimeError("mnk_b200: a tcgen05 launch of this rollout hit an internal barrier timeout; its features (and "
                               f"everything sampled from them) are invalid [{codes}; 0x1 tower, 0x2 heads, 0x1LL weights / "
                               "0x2LL commit watcher / 0x4LL operand TMA of train-mode layer LL]")
        elapsed = time.time() - start
        agent_steps = steps * self.num_envs * self.world_size
        episodes = tot[0]
        return RolloutStats(
            mean_reward=tot[1] / episodes if episodes else 0.0, mean_length=tot[2] / episodes if episodes else 0.0,
            episodes=episodes, wins=tot[3], losses=tot[4], draws=tot[5],
            fps=agent_steps / elapsed if elapsed > 0 else 0.0, rollout_time=elapsed, agent_steps=agent_steps)

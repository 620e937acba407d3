"""PPO learner around the B200 rollout path (the row after the hot path).

``PPOAgent.learn(vec_env) -> TrainingMetrics`` keeps the reference's contract (src/alg/ppo.py:27-166):
one call = one rollout of ``n_steps`` agent steps over all envs (``RolloutCollector``: observation
carried across calls, on-device episode statistics), GAE (``mnk_gae``), then ``ppo_epochs`` passes of
clipped-surrogate minibatch updates (:168-262) whose minibatches are re-materialised from the packed
buffer (``mnk_rollout_gather``).  The rollout's forward runs on the tcgen05 path with train-mode BatchNorm
(``NativeResNet(bn_mode="train")``, as the reference's rollout does); the update itself is stock PyTorch autograd (bf16 autocast, gradient
clipping at 0.5, the caller's optimiser) -- the learner is outside the hot path.

Data parallel: with ``world_size > 1`` every rank collects from its own env shard and the ranks stay in
lock-step by averaging gradients (one flat all-reduce per optimiser step over NCCL) and by normalising
advantages with the GLOBAL mean / std, so N ranks x B samples behave like one process with N*B.
"""
from __future__ import annotations

import time
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn.functional as F

from .dist import average_gradients, broadcast_module, common_minibatches, global_mean_std
from .rollout import RolloutBuffer, RolloutCollector


@dataclass
class TrainingMetrics:          # same fields as src/alg/ppo.py:11-24
    mean_reward: float
    mean_length: float
    actor_loss: float
    critic_loss: float
    entropy_loss: float
    grad_norm: float
    clip_fraction: float
    explained_variance: float
    approx_kl: float
    fps: float
    rollout_time: float
    learn_time: float


class PPOAgent:
    def __init__(self, obs_shape, action_dim, network, n_steps: int, optimizer, gamma=0.99, gae_lambda=0.95,
                 clip_range=0.2, ppo_epochs=4, batch_size=64, value_coef=0.5, entropy_coef=0.01, num_envs=1,
                 device="cuda", lr_scheduler=None, entropy_scheduler=None, k: Optional[int] = None,
                 autocast_dtype: Optional[torch.dtype] = torch.bfloat16, seed: int = 0, env_offset: int = 0,
                 world_size: int = 1, process_group=None, max_grad_norm: float = 0.5, native_rollout: bool = True,
                 graph_rollout: bool = False):
        self.device = torch.device(device)
        self.network = network.to(self.device)
        self.optimizer = optimizer
        self.gamma, self.gae_lambda, self.clip_range = gamma, gae_lambda, clip_range
        self.ppo_epochs, self.batch_size = ppo_epochs, batch_size
        self.value_coef, self.entropy_coef = value_coef, entropy_coef
        self.num_envs, self.n_steps = num_envs, n_steps
        self.lr_scheduler, self.entropy_scheduler = lr_scheduler, entropy_scheduler
        self.autocast_dtype = autocast_dtype
        self.world_size, self.group = world_size, process_group
        self.max_grad_norm = max_grad_norm
        self.buffer = RolloutBuffer(n_steps, num_envs, obs_shape, action_dim, device=device, k=k)
        self.collector = RolloutCollector(num_envs, device=device, seed=seed, row_offset=env_offset,
                                          process_group=process_group, world_size=world_size)
        if world_size > 1:          # replicas start from rank 0's parameters / buffers / optimiser state
            broadcast_module(self.network, self.optimizer, group=process_group)
        # Rollout forward on the tcgen05 path with the semantics of the reference's rollout, which never leaves train mode
        # (src/alg/ppo.py:97): TRAIN-mode BatchNorm for the default architecture (resnet_b_s layout, boards up to 13 rows);
        # the transformers have neither BatchNorm nor dropout, so their native forward is already exact.  Any other module
        # keeps the generic path (stock PyTorch forward on f32 observations).
        self.native, self.graph_rollout = None, graph_rollout
        if native_rollout:
            from . import transformer
            from .resnet import NativeResNet
            try:
                if transformer.supports(self.network):
                    self.native = transformer.NativeTransformer(self.network, device=self.device)
                elif obs_shape[1] <= 13:
                    self.native = NativeResNet(self.network, device=self.device, bn_mode="train")
            except (ValueError, AttributeError):
                self.native = None

    # ------------------------------------------------------------------ reference :78-166
    def learn(self, vec_env) -> TrainingMetrics:
        native = self.native is not None and hasattr(vec_env, "_side")
        if native:
            stats = self.collector.collect(self.native, vec_env, self.buffer, graph=self.graph_rollout)       # :93-124
            _, last_values = self.native.forward_env(vec_env.env, swap=vec_env._side)    # :131-135 (train mode there too)
            if hasattr(self.native, "export_running_stats"):
                self.native.export_running_stats(self.network)  # the update continues from the rollout's statistics
        else:
            stats = self.collector.collect(self.network, vec_env, self.buffer)
            obs = self.collector._last_obs
            with torch.no_grad():                                                    # :131-135 bootstrap value
                _, last_values = self.network(obs["observation"], obs["action_mask"])
        self.buffer.compute_advantages_and_returns(last_values.reshape(self.num_envs), self.gamma, self.gae_lambda)
        learn_start = time.time()
        metrics = self.update_networks()
        learn_time = time.time() - learn_start
        if native:
            self.native.refresh(self.network)                   # new weights (and the update's running statistics)
        if self.lr_scheduler:
            self.lr_scheduler.step()
        if self.entropy_scheduler:
            self.entropy_scheduler.step()
            self.entropy_coef = self.entropy_scheduler.get_last_coef()
        self.buffer.reset()
        return TrainingMetrics(stats.mean_reward, stats.mean_length, *metrics, fps=stats.fps,
                               rollout_time=stats.rollout_time, learn_time=learn_time)

    # ------------------------------------------------------------------ reference :168-262
    def update_networks(self):
        dev = self.device
        totals = torch.zeros(7, device=dev)
        updates = 0
        buf = self.buffer
        # global advantage statistics (== the single-process normalisation of rollout_buffer.py:96-99)
        if self.world_size > 1:
            adv_mean, adv_std = global_mean_std(buf.advantages[:buf.ptr], group=self.group)
        else:       # the reference's own arithmetic (fp32 mean / unbiased std of the flattened rollout)
            flat = buf.advantages[:buf.ptr].view(-1)
            adv_mean, adv_std = flat.mean(), flat.std()
        # every rank must join the same number of gradient all-reduces: with uneven shards the ranks agree on the
        # smallest minibatch count (all-reduce MIN) and the longer loaders drop their last batches
        n_batches = common_minibatches(-(-buf.ptr * buf.num_envs // self.batch_size), self.world_size, self.group, dev)
        for _ in range(self.ppo_epochs):
            for b, (obs, actions, old_log_probs, returns, advantages, masks, _old_values) in enumerate(buf.get_data_loader(
                    self.batch_size, normalize_advantages=False)):
                if b >= n_batches:
                    break
                advantages = (advantages - adv_mean) / (adv_std + 1e-8)
                self.optimizer.zero_grad(set_to_none=True)
                with torch.autocast("cuda", dtype=self.autocast_dtype, enabled=self.autocast_dtype is not None):
                    dist, values = self.network(obs, masks)
                    values = values.squeeze(-1)
                    new_log_probs = dist.log_prob(actions)
                    entropy = dist.entropy().mean()
                    ratio = torch.exp(new_log_probs - old_log_probs)
                    surrogate = torch.min(ratio * advantages,
                                          torch.clamp(ratio, 1.0 - self.clip_range, 1.0 + self.clip_range) * advantages)
                    actor_loss = -surrogate.mean()
                    critic_loss = F.mse_loss(values.float(), returns)
                    loss = actor_loss.float() + self.value_coef * critic_loss.float() - self.entropy_coef * entropy.float()
                loss.backward()
                average_gradients(self.network.parameters(), self.world_size, self.group)
                grad_norm = torch.nn.utils.clip_grad_norm_(self.network.parameters(), self.max_grad_norm)
                self.optimizer.step()
                with torch.no_grad():
                    updates += 1
                    log_ratio = new_log_probs - old_log_probs
                    returns_var = returns.var()
                    explained = torch.where(returns_var > 1e-8, 1 - F.mse_loss(values.float(), returns) / returns_var.clamp(min=1e-8),
                                            torch.zeros((), device=dev))
                    totals += torch.stack([actor_loss.detach().float(), critic_loss.detach().float(), (-entropy).detach().float(),
                                           grad_norm.float(), (torch.abs(ratio - 1.0) > self.clip_range).float().mean(),
                                           explained.float(), ((torch.exp(log_ratio) - 1) - log_ratio).mean().float()])
        return tuple((totals / max(updates, 1)).tolist())      # one host read per iteration

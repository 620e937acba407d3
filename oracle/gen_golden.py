"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (the reference lives at /root/reference there and
does not travel to the GPU box):

    python oracle/gen_golden.py            # writes tests/golden/

The fixtures are small (packed bits) and committed; tests compare the oracle
(oracle/mnk_oracle.{c,py}, oracle/torch_port.py) and the CUDA path against them.
Nothing here is imported by the product.

What is recorded
  env_trace_<m>x<n>x<k>.npz   step / step_subset / reset(idx) / reset() sequences with random
                              legal moves, occasional illegal and post-terminal moves
  env_poke_<m>x<n>x<k>.npz    single steps from randomly poked positions (boards, player and
                              move counter written directly, as the reference tests do)
  wrapper_trace_<...>.npz     TorchSelfPlayWrapper reset/step sequences with injected sides
                              (torch.randint patched) and a deterministic opponent
"""
from __future__ import annotations

import os
import sys
from unittest import mock

import numpy as np
import torch

REF = os.environ.get("MNK_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(REF, "src"))
sys.path.insert(0, REF)

from env.torch_vector_mnk_env import TorchVectorMnkEnv  # noqa: E402  (the reference)
from selfplay.torch_self_play_wrapper import TorchSelfPlayWrapper  # noqa: E402
from selfplay.policy import RandomPolicy  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

OP_STEP_SUBSET, OP_RESET_IDX, OP_RESET_ALL, OP_STEP = 0, 1, 2, 3


def pack(a) -> np.ndarray:
    return np.packbits(np.asarray(a).astype(bool).reshape(a.shape[0], -1), axis=1)


def gen_env_trace(m, n, k, num_envs, steps, seed):
    rng = np.random.default_rng(seed)
    env = TorchVectorMnkEnv(m, n, k, num_envs, device="cpu")
    obs = env.reset()
    cells = m * n
    rec = {key: [] for key in ("op", "active", "actions", "boards", "mask", "player", "count", "rewards", "dones")}
    done_state = np.zeros(num_envs, dtype=bool)
    for t in range(steps):
        mask = obs["action_mask"].numpy()
        roll = rng.random()
        active = np.ones(num_envs, dtype=bool)
        actions = np.full(num_envs, -1, dtype=np.int64)
        rewards = np.zeros(num_envs, dtype=np.float32)
        dones = np.zeros(num_envs, dtype=bool)
        if roll < 0.02:
            op = OP_RESET_ALL
            obs = env.reset()
            done_state[:] = False
        elif roll < 0.22 and done_state.any():
            op = OP_RESET_IDX
            # reset most (not all) finished envs; the rest keep being stepped post-terminal
            active = done_state & (rng.random(num_envs) < 0.8)
            idx = torch.from_numpy(np.nonzero(active)[0])
            obs = env.reset(idx)
            done_state[active] = False
        else:
            if roll < 0.55:
                op = OP_STEP_SUBSET
                active = rng.random(num_envs) < 0.6
                active[rng.integers(0, num_envs)] = True   # the reference raises on an empty subset (view(0,-1))
            else:
                op = OP_STEP
            for e in np.nonzero(active)[0]:
                legal = np.nonzero(mask[e])[0]
                if len(legal) == 0 or rng.random() < 0.03:
                    actions[e] = rng.integers(0, cells)          # possibly occupied: silently applied
                else:
                    actions[e] = rng.choice(legal)
            idx = np.nonzero(active)[0]
            a = torch.from_numpy(actions[idx])
            if op == OP_STEP:
                obs, r, d = env.step(a)
            else:
                obs, r, d = env.step_subset(a, torch.from_numpy(idx))
            rewards, dones = r.numpy().copy(), d.numpy().copy()
            done_state |= dones
        rec["op"].append(op)
        rec["active"].append(active.copy())
        rec["actions"].append(actions)
        rec["boards"].append(pack(env.boards.numpy()))
        rec["mask"].append(pack(obs["action_mask"].numpy()))
        rec["player"].append(env.current_player.numpy().copy())
        rec["count"].append(env.move_counts.numpy().copy())
        rec["rewards"].append(rewards)
        rec["dones"].append(dones)
        assert torch.equal(obs["observation"], env.boards)
    out = {key: np.stack(v) for key, v in rec.items()}
    out["geom"] = np.array([m, n, k, num_envs], dtype=np.int64)
    return out


def gen_env_poke(m, n, k, cases, seed):
    """One env.step from a position written straight into env.boards / current_player /
    move_counts.  Dense random positions hit every line direction, overlines and the
    row-wrap non-wins; `count` near m*n exercises the draw rule."""
    rng = np.random.default_rng(seed)
    cells = m * n
    env = TorchVectorMnkEnv(m, n, k, cases, device="cpu")
    env.reset()
    dens = rng.choice([0.15, 0.35, 0.5, 0.7], size=cases)
    occ = rng.random((cases, cells)) < dens[:, None]
    owner = rng.random((cases, cells)) < 0.5
    both = rng.random((cases, cells)) < 0.01          # illegal-move leftovers: both planes set
    init = np.zeros((cases, 2, cells), dtype=np.float32)
    init[:, 0] = (occ & owner) | (occ & both)
    init[:, 1] = (occ & ~owner) | (occ & both)
    player = rng.integers(0, 2, size=cases)
    actions = rng.integers(0, cells, size=cases)
    # plant structure in ~60% of the cases so that sparse geometries (k = 8) also see lines:
    #   kind 0: k-1 .. k+1 stones of the mover along a direction, the action completing the gap
    #   kind 1: k stones contiguous in the FLATTENED index straddling a row end (must not win)
    for c in range(cases):
        u = rng.random()
        p = player[c]
        if u < 0.45:
            dr, dc = [(0, 1), (1, 0), (1, 1), (1, -1)][rng.integers(0, 4)]
            length = int(rng.integers(k - 1, k + 2))
            length = min(length, n if dr == 0 else m if dc == 0 else min(m, n))
            r0 = rng.integers(0, m - (length - 1) * dr)
            c0 = rng.integers(0, n - (length - 1)) if dc >= 0 else rng.integers(length - 1, n)
            line = [(r0 + t * dr) * n + (c0 + t * dc) for t in range(length)]
            gap = line[rng.integers(0, length)]
            for cell in line:
                init[c, p, cell] = 1.0
                init[c, 1 - p, cell] = 0.0
            init[c, p, gap] = 0.0
            if rng.random() < 0.8:
                actions[c] = gap
        elif u < 0.6 and m > 1:
            r0 = rng.integers(0, m - 1)
            start = r0 * n + n - int(rng.integers(1, k))
            for cell in range(start, min(start + k, cells)):
                init[c, p, cell] = 1.0
                init[c, 1 - p, cell] = 0.0
            actions[c] = min(start + k - 1, cells - 1)
    count = np.where(rng.random(cases) < 0.3, cells - 1, rng.integers(0, cells + 3, size=cases))
    env.boards[:] = torch.from_numpy(init.reshape(cases, 2, m, n))
    env.current_player[:] = torch.from_numpy(player)
    env.move_counts[:] = torch.from_numpy(count)
    obs, r, d = env.step(torch.from_numpy(actions))
    return {
        "geom": np.array([m, n, k, cases], dtype=np.int64),
        "init_boards": pack(init), "init_player": player.astype(np.int64), "init_count": count.astype(np.int64),
        "actions": actions.astype(np.int64),
        "boards": pack(env.boards.numpy()), "mask": pack(obs["action_mask"].numpy()),
        "player": env.current_player.numpy().copy(), "count": env.move_counts.numpy().copy(),
        "rewards": r.numpy().copy(), "dones": d.numpy().copy(),
    }


class HashPolicy:
    """Deterministic, row-wise opponent used on BOTH sides of the wrapper parity tests.
    score = sum(obs[c, cell] * (c * cells + cell + 1)); picks the (score mod #legal)-th legal
    cell (ascending), or score mod cells on a full board.  Takes obs_dict only, like the
    reference tests' ScriptedPolicy."""

    def act(self, obs_dict):
        obs = obs_dict["observation"]
        mask = obs_dict["action_mask"]
        b = obs.shape[0]
        cells = mask.shape[1]
        w = torch.arange(1, 2 * cells + 1, dtype=torch.int64)
        score = (obs.reshape(b, -1).to(torch.int64) * w).sum(dim=1)
        cnt = mask.sum(dim=1)
        j = score % torch.clamp(cnt, min=1)
        rank = torch.cumsum(mask.to(torch.int64), dim=1) - 1
        hit = mask & (rank == j[:, None])
        picked = torch.argmax(hit.to(torch.int64), dim=1)
        return torch.where(cnt == 0, score % cells, picked)


def gen_wrapper_trace(m, n, k, num_envs, steps, seed, use_options_reset):
    rng = np.random.default_rng(seed)
    cells = m * n
    sides = rng.integers(0, 2, size=(steps + 1, num_envs)).astype(np.int64)
    env = TorchVectorMnkEnv(m, n, k, num_envs, device="cpu")
    wrapper = TorchSelfPlayWrapper(env)
    wrapper.set_opponent(HashPolicy())
    pending_rows = {"row": None}

    def fake_randint(low, high, size, **kw):
        assert (low, high) == (0, 2)
        row = pending_rows["row"]
        assert row is not None and len(row) == size[0]
        return torch.from_numpy(row.copy())

    rec = {key: [] for key in ("actions", "obs", "mask", "rewards", "terminated", "agent_side", "pending",
                               "boards", "player", "count")}
    with mock.patch.object(torch, "randint", fake_randint):
        if use_options_reset:
            obs, _ = wrapper.reset(options={"agent_side": torch.from_numpy(sides[0])})
        else:
            pending_rows["row"] = sides[0]
            obs, _ = wrapper.reset()
        obs0, mask0 = pack(obs["observation"].numpy()), pack(obs["action_mask"].numpy())
        side0 = wrapper.agent_side.numpy().copy()
        for t in range(steps):
            mask = obs["action_mask"].numpy()
            actions = np.zeros(num_envs, dtype=np.int64)
            for e in range(num_envs):
                legal = np.nonzero(mask[e])[0]
                actions[e] = rng.integers(0, cells) if rng.random() < 0.02 else rng.choice(legal)
            reset_idx = np.nonzero(wrapper.pending_resets.numpy())[0]
            pending_rows["row"] = sides[t + 1][reset_idx]
            obs, r, term, trunc, info = wrapper.step(torch.from_numpy(actions))
            assert not trunc.any() and info == {}
            rec["actions"].append(actions)
            rec["obs"].append(pack(obs["observation"].numpy()))
            rec["mask"].append(pack(obs["action_mask"].numpy()))
            rec["rewards"].append(r.numpy().copy())
            rec["terminated"].append(term.numpy().copy())
            rec["agent_side"].append(wrapper.agent_side.numpy().copy())
            rec["pending"].append(wrapper.pending_resets.numpy().copy())
            rec["boards"].append(pack(env.boards.numpy()))
            rec["player"].append(env.current_player.numpy().copy())
            rec["count"].append(env.move_counts.numpy().copy())
    out = {key: np.stack(v) for key, v in rec.items()}
    out.update(geom=np.array([m, n, k, num_envs], dtype=np.int64), sides=sides,
               options_reset=np.array(int(use_options_reset)), obs0=obs0, mask0=mask0, side0=side0)
    return out


def gen_random_policy(seed):
    """RandomPolicy.act(deterministic=True) (policy.py:26-27) incl. an all-masked row."""
    rng = np.random.default_rng(seed)
    mask = rng.random((64, 81)) < 0.4
    mask[3] = False
    mask[7] = True
    act = RandomPolicy(81).act({"action_mask": torch.from_numpy(mask)}, deterministic=True)
    return {"mask": pack(mask), "cells": np.array(81), "first_legal": act.numpy()}


def gen_rollout_buffer(seed):
    """RolloutBuffer.add / compute_advantages_and_returns (src/alg/rollout_buffer.py:47-80) on random data."""
    import importlib
    rb = importlib.import_module("alg.rollout_buffer")
    rng = np.random.default_rng(seed)
    steps, ne, m, n = 37, 19, 3, 3
    buf = rb.RolloutBuffer(steps, ne, (2, m, n), m * n, device="cpu")
    obs = (rng.random((steps, ne, 2, m, n)) < 0.3).astype(np.float32)
    obs[:, :, 1] *= 1 - obs[:, :, 0]
    masks = ~(obs[:, :, 0].astype(bool) | obs[:, :, 1].astype(bool)).reshape(steps, ne, m * n)
    actions = rng.integers(0, m * n, size=(steps, ne))
    rewards = rng.choice([-1.0, 0.0, 0.0, 0.0, 1.0], size=(steps, ne)).astype(np.float32)
    values = rng.normal(size=(steps, ne)).astype(np.float32)
    logp = -rng.random((steps, ne)).astype(np.float32) * 3
    dones = rng.random((steps, ne)) < 0.15
    last = rng.normal(size=ne).astype(np.float32)
    for t in range(steps):
        buf.add(*[torch.from_numpy(x[t]) for x in (obs, actions, rewards)], torch.from_numpy(values[t]).view(-1, 1),
                torch.from_numpy(logp[t]), torch.from_numpy(dones[t]), torch.from_numpy(masks[t]))
    buf.compute_advantages_and_returns(torch.from_numpy(last), 0.99, 0.95)
    return {"geom": np.array([steps, ne, m, n]), "obs": pack(obs.reshape(steps * ne, -1)), "masks": pack(masks.reshape(steps * ne, -1)),
            "actions": actions, "rewards": rewards, "values": values, "log_probs": logp, "dones": dones, "last_values": last,
            "gamma": np.float32(0.99), "lam": np.float32(0.95), "advantages": buf.advantages.numpy().copy(),
            "returns": buf.returns.numpy().copy(), "stored_obs": pack(buf.observations.numpy().reshape(steps * ne, -1)),
            "stored_masks": pack(buf.action_masks.numpy().reshape(steps * ne, -1))}


def gen_resnet(seed, m, n, k, batch):
    """Eval-mode forward of the reference's default network "resnet_b_s" (configs.py:28-35,
    resnet.py:73-95) with randomised weights AND BatchNorm statistics, on positions from random play."""
    import importlib
    cfg = importlib.import_module("alg.architectures.configs")
    torch.manual_seed(seed)
    rng = np.random.default_rng(seed)
    net = cfg.ResNetSActorCritic((2, m, n), m * n)
    with torch.no_grad():
        for mod in net.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.normal_(0, 0.3)
                mod.running_var.uniform_(0.5, 1.5)
                mod.weight.uniform_(0.7, 1.3)
                mod.bias.normal_(0, 0.2)
            elif isinstance(mod, (torch.nn.Conv2d, torch.nn.Linear)):
                mod.bias.normal_(0, 0.1)
            elif isinstance(mod, torch.nn.LayerNorm):
                mod.weight.uniform_(0.8, 1.2)
                mod.bias.normal_(0, 0.1)
        net.policy_head[7].weight.mul_(60.0)        # the reference initialises this layer with gain 0.01: scale the
    net.eval()                                      # logits up so that the parity check is not vacuous
    env = TorchVectorMnkEnv(m, n, k, batch, device="cpu")
    obs = env.reset()
    depth = rng.integers(0, m * n - 1, size=batch)
    for t in range(m * n - 1):
        a = RandomPolicy(m * n).act(obs)
        idx = torch.from_numpy(np.nonzero(depth > t)[0])
        if len(idx) == 0:
            break
        obs, _, _ = env.step_subset(a[idx], idx)
    x = obs["observation"].clone()
    flip = torch.from_numpy(rng.random(batch) < 0.5)
    x[flip] = torch.flip(x[flip], dims=(1,))
    mask = obs["action_mask"].clone()
    mask[0] = False                                 # an all-masked row
    with torch.no_grad():
        dist, value = net(x, mask)
        body = net.forward_body(x)
    out = {f"param/{key}": v.numpy() for key, v in net.state_dict().items()}
    out.update(geom=np.array([m, n, k, batch]), obs=pack(x.numpy()), mask=pack(mask.numpy()), logits=dist.logits.numpy(),
               value=value.numpy(), body_absmax=np.float32(body.abs().max().item()))
    return out


def gen_widenet(arch, seed, m, n, k, batch):
    """Eval-mode forward of the reference's wider convolutional networks (configs.py:36-65: "resnet_b_l" = 80 channels x
    5 residual blocks, "cnn_b_s" = [56] * 4, "cnn_b_l" = [96] * 8) with randomised weights and BatchNorm statistics.
    The conv / linear weights are rounded to fp16-representable values first so that the fixture can store them as fp16
    (half the size) and still hold EXACTLY the parameters the recorded outputs were computed with."""
    import importlib
    cfg = importlib.import_module("alg.architectures.configs")
    cls = {"resnet_b_l": cfg.ResNetLActorCritic, "cnn_b_s": cfg.CnnSActorCritic, "cnn_b_l": cfg.CnnLActorCritic,
           "transformer_b_s": cfg.TransformerSActorCritic, "transformer_b_l": cfg.TransformerLActorCritic}[arch]
    torch.manual_seed(seed)
    rng = np.random.default_rng(seed)
    net = cls((2, m, n), m * n)
    actor = net.policy_head if hasattr(net, "policy_head") else net.actor
    with torch.no_grad():
        for mod in net.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.normal_(0, 0.3)
                mod.running_var.uniform_(0.5, 1.5)
                mod.weight.uniform_(0.7, 1.3)
                mod.bias.normal_(0, 0.2)
            elif isinstance(mod, (torch.nn.Conv2d, torch.nn.Linear)):
                mod.bias.normal_(0, 0.1)
            elif isinstance(mod, torch.nn.LayerNorm):
                mod.weight.uniform_(0.8, 1.2)
                mod.bias.normal_(0, 0.1)
            elif isinstance(mod, torch.nn.MultiheadAttention):
                mod.in_proj_bias.normal_(0, 0.1)
        if hasattr(net, "pos_embed"):                # (initialised with std 0.02: make the embedding matter)
            net.pos_embed.normal_(0, 0.5)
            net.cell_embed.weight.normal_(0, 0.5)
        actor[7].weight.mul_(60.0)                   # (gain 0.01 at initialisation: scale the logits up, as in gen_resnet)
        for name, prm in net.named_parameters():     # every matrix-shaped ".weight" is stored as fp16 below
            if name.endswith(".weight") and prm.dim() >= 2:
                prm.copy_(prm.half().float())
    net.eval()
    env = TorchVectorMnkEnv(m, n, k, batch, device="cpu")
    obs = env.reset()
    depth = rng.integers(0, m * n - 1, size=batch)
    for t in range(m * n - 1):
        a = RandomPolicy(m * n).act(obs)
        idx = torch.from_numpy(np.nonzero(depth > t)[0])
        if len(idx) == 0:
            break
        obs, _, _ = env.step_subset(a[idx], idx)
    x = obs["observation"].clone()
    flip = torch.from_numpy(rng.random(batch) < 0.5)
    x[flip] = torch.flip(x[flip], dims=(1,))
    mask = obs["action_mask"].clone()
    mask[0] = False                                 # an all-masked row
    with torch.no_grad():
        dist, value = net(x, mask)
    out = {}
    for key, v in net.state_dict().items():
        half = key.endswith(".weight") and v.dim() >= 2
        out[f"param/{key}"] = v.numpy().astype(np.float16) if half else v.numpy()
    out.update(geom=np.array([m, n, k, batch]), obs=pack(x.numpy()), mask=pack(mask.numpy()), logits=dist.logits.numpy(),
               value=value.numpy(), arch=np.array(arch))
    return out


WIDENETS = [("resnet_b_l", 31, 9, 9, 5, 40), ("cnn_b_s", 32, 9, 9, 5, 40), ("cnn_b_l", 33, 7, 7, 4, 24)]
TRANSFORMERS = [("transformer_b_s", 41, 9, 9, 5, 40), ("transformer_b_l", 42, 7, 7, 4, 24)]


def gen_resnet_train(seed, m, n, k, batch):
    """TRAIN-mode forward of the same network, as PPOAgent.learn's rollout runs it (src/alg/ppo.py:97: the module is
    never switched to eval there): BatchNorm uses the statistics of this batch and updates its running buffers.
    Records the parameters BEFORE the call, the outputs, and the running statistics AFTER it."""
    import importlib
    cfg = importlib.import_module("alg.architectures.configs")
    base = gen_resnet(seed, m, n, k, batch)         # same randomised weights / positions as the eval fixture recipe
    net = cfg.ResNetSActorCritic((2, m, n), m * n)
    net.load_state_dict({key[len("param/"):]: torch.from_numpy(v) for key, v in base.items() if key.startswith("param/")})
    net.train()
    x = torch.from_numpy(np.unpackbits(base["obs"], axis=1)[:, :2 * m * n].reshape(batch, 2, m, n).astype(np.float32))
    mask = torch.from_numpy(np.unpackbits(base["mask"], axis=1)[:, :m * n].astype(bool))
    with torch.no_grad():
        dist, value = net(x, mask)
    out = {key: v for key, v in base.items() if key.startswith("param/") or key in ("geom", "obs", "mask")}
    out.update(logits=dist.logits.numpy(), value=value.numpy())
    for key, v in net.state_dict().items():
        if "running_" in key or "num_batches" in key:
            out[f"after/{key}"] = v.numpy()
    return out


def main():
    if "--transformer-only" in sys.argv:
        for arch, seed, m, n, k, batch in TRANSFORMERS:
            np.savez_compressed(os.path.join(OUT, f"widenet_{arch}_{m}x{n}.npz"), **gen_widenet(arch, seed, m, n, k, batch))
        return
    if "--widenet-only" in sys.argv:
        for arch, seed, m, n, k, batch in WIDENETS:
            np.savez_compressed(os.path.join(OUT, f"widenet_{arch}_{m}x{n}.npz"), **gen_widenet(arch, seed, m, n, k, batch))
        return
    if "--resnet-train-only" in sys.argv:
        np.savez_compressed(os.path.join(OUT, "resnet_train_b_s_9x9.npz"), **gen_resnet_train(21, 9, 9, 5, 96))
        np.savez_compressed(os.path.join(OUT, "resnet_train_b_s_7x7.npz"), **gen_resnet_train(22, 7, 7, 4, 40))
        return
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    env_cfgs = [  # m, n, k, envs, steps
        (3, 3, 3, 16, 120), (9, 9, 5, 8, 400), (13, 13, 5, 4, 500), (19, 19, 5, 2, 700),
        (7, 11, 4, 6, 300), (5, 4, 3, 8, 150), (4, 6, 4, 8, 150), (8, 8, 8, 4, 300),
    ]
    for i, (m, n, k, ne, st) in enumerate(env_cfgs):
        np.savez_compressed(os.path.join(OUT, f"env_trace_{m}x{n}x{k}.npz"), **gen_env_trace(m, n, k, ne, st, 100 + i))
        np.savez_compressed(os.path.join(OUT, f"env_poke_{m}x{n}x{k}.npz"), **gen_env_poke(m, n, k, 256, 200 + i))
    wr_cfgs = [(3, 3, 3, 16, 120, True), (3, 3, 3, 16, 120, False), (9, 9, 5, 8, 260, False),
               (5, 4, 3, 8, 160, False), (13, 13, 5, 4, 260, True), (19, 19, 5, 2, 300, False)]
    for i, (m, n, k, ne, st, opt) in enumerate(wr_cfgs):
        tag = "opt" if opt else "rnd"
        np.savez_compressed(os.path.join(OUT, f"wrapper_trace_{m}x{n}x{k}_{tag}.npz"),
                            **gen_wrapper_trace(m, n, k, ne, st, 300 + i, opt))
    np.savez_compressed(os.path.join(OUT, "random_policy_first_legal.npz"), **gen_random_policy(7))
    np.savez_compressed(os.path.join(OUT, "rollout_buffer_gae.npz"), **gen_rollout_buffer(9))
    np.savez_compressed(os.path.join(OUT, "resnet_b_s_9x9.npz"), **gen_resnet(11, 9, 9, 5, 96))
    np.savez_compressed(os.path.join(OUT, "resnet_b_s_13x13.npz"), **gen_resnet(12, 13, 13, 5, 24))
    np.savez_compressed(os.path.join(OUT, "resnet_train_b_s_9x9.npz"), **gen_resnet_train(21, 9, 9, 5, 96))
    np.savez_compressed(os.path.join(OUT, "resnet_train_b_s_7x7.npz"), **gen_resnet_train(22, 7, 7, 4, 40))
    for arch, seed, m, n, k, batch in WIDENETS + TRANSFORMERS:
        np.savez_compressed(os.path.join(OUT, f"widenet_{arch}_{m}x{n}.npz"), **gen_widenet(arch, seed, m, n, k, batch))
    total = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print(f"wrote {len(os.listdir(OUT))} fixtures, {total / 1024:.1f} KiB -> {os.path.normpath(OUT)}")


if __name__ == "__main__":
    main()

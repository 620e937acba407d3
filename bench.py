#!/usr/bin/env python
"""bench.py -- benchmarks of the MNK hot path (BASELINE.json configs).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg2|cfg3|cfg4|cfg5]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

The ONE JSON line rank 0 prints is the headline workload, BASELINE.json configs[1] (cfg2: Gomoku 9x9x5,
65,536 envs per GPU, seeded uniformly-random legal actions; one "step" = TorchVectorMnkEnv.step over the whole batch --
stone placement, win/draw check, rewards, dones, player toggle, API-exact f32 observation + bool mask materialised --
fused with the reset of finished envs).  With no --workload the same line carries a `secondary` object holding the
lines of the other BASELINE configs, each measured the same way in the same process:
  cfg3  9x9x5 self-play rollout (agent + opponent network forward, sampling, wrapper, packed PPO buffer), agent-steps/s
  cfg4  13x13x5, 1,048,576 envs sharded over the N GPUs (strong scaling)
  cfg5  19x19x5, 524,288 envs per GPU
Envs are independent, so ranks hold disjoint shards; there is no data-path collective (one NCCL all-reduce of
end-of-run statistics, max-over-ranks of the timings).

Measurement protocol (env workloads)
  * The games are advanced 64 plies to a stationary mix of depths, the packed state is snapshotted and the W+K action
    batches are pre-generated ON DEVICE by playing the games once (untimed); the state is restored.
  * `value`: the K steps are K kernel nodes of one CUDA graph (inputs resident in HBM).  The graph is replayed R times
    (--replays, default 20; the state is restored between replays, outside the timed span), each replay bracketed by
    CUDA events, max over ranks per replay; `value` uses the MEDIAN replay and `timing` reports min / max / first.
    Outputs rotate through a ring of buffers larger than L2.  The state digest after the last replay must equal the
    digest reached during trace generation (the timed region demonstrably did the work).
  * `roofline`: algorithmic bytes per launch (814 B/env-step API-exact at 9x9) / the median launch
    duration, against MEASURED_PEAKS.json.  `frac_lower_bound` charges the WHOLE R-replay sequence (restores included)
    and subtracts one L2 capacity of possibly-unwritten dirty lines: a bound on the DRAM-honest fraction.
  * `e2e`: the same K steps through the public host-buffer API (TorchVectorMnkEnv.step_host_loop -> mnk_step_host_loop):
    pinned int64 actions in, f32 rewards + bool dones out, slab-pipelined copies; plus the one-step-per-call API
    (step_host) and a variant that also brings every observation + mask to the host (PCIe-bound).
  * `cpu_baseline` (rank 0, N=1): the UNMODIFIED reference (oracle/_ref, staged by __graft_entry__.build(); kind
    "reference") -- or, where it is absent, oracle/torch_port.py (kind "port") -- replaying the head of the SAME action
    trace from the SAME positions on all host cores, cross-checked against the GPU state.
  * `--impl reference`: that CPU arm alone.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "rl-selfplay-mnk_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

UNIT = "env-steps/s"
MIX_PLIES = 64
L2_BYTES = 126 * 1024 * 1024
SEED = 20261018
FLOP_PER_FORWARD = {(9, 9): 12.136e6, (13, 13): 25.3e6, (19, 19): 54.1e6}     # resnet_b_s, SURVEY 8a15


class Workload:
    """Board geometry, per-GPU env count and algorithmic bytes per env-step (8 B action + 2 x packed state
    + f32 observation + bool mask + reward/done)."""
    TABLE = {
        "cfg2": (9, 9, 5, lambda n: 65536, "cfg2: gomoku 9x9x5, 65536 envs/GPU", "weak"),
        "cfg4": (13, 13, 5, lambda n: 1048576 // n, "cfg4: 13x13x5, 1,048,576 envs sharded over the GPUs", "strong"),
        "cfg5": (19, 19, 5, lambda n: 4194304 // 8, "cfg5: 19x19x5, 524,288 envs/GPU (4,194,304 over 8 GPUs)", "weak"),
    }

    def __init__(self, name: str, gpus: int, envs_override=None):
        self.name = name
        self.m, self.n, self.k, envs_fn, self.desc, self.scaling = self.TABLE[name]
        self.cells = self.m * self.n
        self.envs = int(envs_override or envs_fn(max(gpus, 1)))
        words = (self.m * (self.n + 1) + 63) // 64
        state_bytes = 2 * 8 * words + 4
        self.alg_bytes = 8 + 2 * state_bytes + 8 * self.cells + self.cells + 5      # 814 at 9x9
        self.packed_bytes = 8 + 2 * state_bytes + 5                                 # 85 at 9x9
        self.metric = f"env steps/sec ({self.m}x{self.n}x{self.k}, win-check)"

    def config(self, world: int):
        """Workload description shared verbatim by the B200 arm and the reference arm."""
        per_set = self.envs * 9 * self.cells
        ring = max(2, -(-3 * L2_BYTES // per_set))
        return {
            "workload": (f"{self.desc} ({self.envs} envs on this GPU), seeded random legal actions, env.step (placement + "
                         "k-in-a-row win/draw check + rewards/dones + f32 obs & bool mask materialised) + auto-reset"),
            "envs_per_gpu": self.envs, "global_envs": self.envs * world,
            "l2": f"obs/mask outputs rotate through a ring of {ring} buffer sets ({ring * per_set / 2**20:.0f} MiB > 126 MiB L2); "
                  "every step reads a distinct action batch",
            "parallelism": f"env-shard x{world}, no per-step collective",
        }


# ------------------------------------------------------------------------------------------------
# process context, clocks sampler (NVML), peaks
# ------------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (the B200 arm has no CPU fallback; use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        from mnk_b200.dist import pin_to_gpu_numa
        self.numa = pin_to_gpu_numa(self.local_rank)             # host threads (and the pinned buffers they first touch)
        if self.world > 1:                                       # next to this GPU's PCIe root
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, values, op="max"):
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            ops = {"max": self.dist.ReduceOp.MAX, "min": self.dist.ReduceOp.MIN, "sum": self.dist.ReduceOp.SUM}
            self.dist.all_reduce(t, op=ops[op])
        return [float(x) for x in t.tolist()]

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


class ClockSampler:
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if bits & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.002)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def measured_peak(key="hbm_gbs"):
    fallback = {"hbm_gbs": (6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"),
                "bf16_tflops_sustained": (1400.0, "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)")}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)[key]), f"MEASURED_PEAKS.json {key}"
    except Exception:
        return fallback[key]


def profiled_traffic(wl: Workload):
    """DRAM bytes per launch of step_dense_kernel from the committed steady-state ncu capture (profiles/), cfg2 only."""
    if (wl.m, wl.n, wl.envs) != (9, 9, 65536):
        return None, None
    for name in ("r02_step_dense_steady_traffic.json", "step_dense_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                d = json.load(f)
            return d.get("dram_bytes_per_launch"), f"profiles/{name}"
        except Exception:
            continue
    return None, None


# ------------------------------------------------------------------------------------------------
# CPU arm: the unmodified reference (oracle/_ref) or, where absent, the op-for-op port (oracle/torch_port.py)
# ------------------------------------------------------------------------------------------------
class ReferenceEnvArm:
    """src/env/torch_vector_mnk_env.py + src/selfplay/policy.py::RandomPolicy of the UNMODIFIED reference tree."""
    kind = "reference"

    def __init__(self, wl: Workload, envs: int, device="cpu"):
        from oracle import ref_tree
        env_mod, pol_mod = ref_tree.load("env.torch_vector_mnk_env", "selfplay.policy")
        self.env = env_mod.TorchVectorMnkEnv(wl.m, wl.n, wl.k, envs, device=device)
        self.policy = pol_mod.RandomPolicy(wl.cells)
        self.what = f"the unmodified reference ({os.path.relpath(ref_tree.root(), ROOT)}/src/env/torch_vector_mnk_env.py)"

    def reset(self):
        return self.env.reset()

    def set_state(self, planes, to_move, plies):
        self.env.boards.copy_(planes), self.env.current_player.copy_(to_move), self.env.move_counts.copy_(plies)
        return self.env.observe()

    def draw(self, obs):
        return self.policy.act(obs)

    def step_and_reset(self, actions):
        import torch
        obs, r, d = self.env.step(actions)
        idx = torch.nonzero(d).squeeze(1)
        if idx.numel():
            obs = self.env.reset(idx)
        return obs

    def state(self):
        return self.env.boards, self.env.move_counts


class PortEnvArm:
    kind = "port"

    def __init__(self, wl: Workload, envs: int, device="cpu"):
        from oracle import torch_port as tp
        self.tp = tp
        self.s = tp.port_make(wl.m, wl.n, wl.k, envs, device=device)
        self.what = "oracle/torch_port.py (the reference's torch op sequence: index_put, 3x conv2d, threshold, any, clone)"

    def reset(self):
        return self.tp.port_reset(self.s)

    def set_state(self, planes, to_move, plies):
        self.s.planes.copy_(planes), self.s.to_move.copy_(to_move), self.s.plies.copy_(plies)
        return self.tp.port_observe(self.s)

    def draw(self, obs):
        return self.tp.port_uniform_legal(obs["action_mask"])

    def step_and_reset(self, actions):
        import torch
        obs, r, d = self.tp.port_step(self.s, actions)
        idx = torch.nonzero(d).squeeze(1)
        if idx.numel():
            obs = self.tp.port_reset(self.s, idx)
        return obs

    def state(self):
        return self.s.planes, self.s.plies


def make_cpu_arm(wl: Workload, envs: int, device="cpu", force_port=False):
    from oracle import ref_tree
    if ref_tree.available() and not force_port:
        return ReferenceEnvArm(wl, envs, device)
    return PortEnvArm(wl, envs, device)


def cpu_env_run(arm, actions, warmup, steps, budget_s, check_state=None, init_state=None, device="cpu"):
    """Replays `actions[t]` ([T, envs] int64 legal actions, or None => draw them with the reference's RandomPolicy
    arithmetic outside the timed sections).  Times only env.step + env.reset(done_idx), like the harness of SURVEY 8d."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    obs = arm.reset()
    if init_state is not None:          # start from the same mid-game positions as the B200 arm
        obs = arm.set_state(*init_state)
    on_gpu = str(device) != "cpu"
    sync = torch.cuda.synchronize if on_gpu else (lambda: None)
    timed, done_steps = 0.0, 0
    t_begin = time.perf_counter()
    for t in range(warmup + steps):
        a = actions[t] if actions is not None else arm.draw(obs)
        sync()
        t0 = time.perf_counter()
        obs = arm.step_and_reset(a)
        sync()
        dt = time.perf_counter() - t0
        if t >= warmup:
            timed += dt
            done_steps += 1
        if time.perf_counter() - t_begin > budget_s and done_steps >= 3:
            break
    ok = check_state(arm, warmup + done_steps) if check_state is not None else None
    envs = arm.state()[1].numel()
    return {"steps": done_steps, "seconds": timed, "value": envs * done_steps / timed, "threads": torch.get_num_threads(),
            "parity": ok, "envs": envs}


def cpu_rollout_run(wl: Workload, envs: int, n_steps: int, budget_s: float):
    """cfg3 on the host cores: the UNMODIFIED reference's PPOAgent.learn (src/alg/ppo.py:78-133) on its own env /
    wrapper / network / NNPolicy opponent / RolloutBuffer, with the update half stubbed out; the figure is the `fps` the
    reference computes itself (:126-129) over its rollout section."""
    import copy
    import torch
    from oracle import ref_tree
    if not ref_tree.available():
        return None
    torch.set_num_threads(os.cpu_count() or 1)
    ppo, hw, cfg, env_mod, wrap_mod, pol_mod = ref_tree.load(
        "alg.ppo", "utils.hardware", "alg.architectures.configs", "env.torch_vector_mnk_env",
        "selfplay.torch_self_play_wrapper", "selfplay.policy")
    torch.manual_seed(0)
    net = cfg.ResNetSActorCritic((2, wl.m, wl.n), wl.cells)
    agent = ppo.PPOAgent((2, wl.m, wl.n), wl.cells, net, hw_config=hw.HardwareConfig("cpu", torch.float32, False, None),
                         n_steps=n_steps, optimizer=torch.optim.SGD(net.parameters(), lr=0.0), num_envs=envs)
    agent.update_networks = lambda: (0.0,) * 7          # the rollout section only
    wrapper = wrap_mod.TorchSelfPlayWrapper(env_mod.TorchVectorMnkEnv(wl.m, wl.n, wl.k, envs, device="cpu"))
    wrapper.set_opponent(pol_mod.NNPolicy(copy.deepcopy(net)))
    t0 = time.perf_counter()
    agent.learn(wrapper)                                 # warm-up (includes the lazy reset, :81-84)
    fps, calls = [], 0
    while calls < 3 and (time.perf_counter() - t0 < budget_s or not fps):
        fps.append(agent.learn(wrapper).fps)
        calls += 1
    return {"value": statistics.median(fps), "unit": "agent-steps/s", "cores": torch.get_num_threads(), "kind": "reference",
            "sample": f"{calls} x PPOAgent.learn rollout sections of {n_steps} steps x {envs} envs (unmodified reference: env, wrapper, "
                      "resnet_b_s agent in train mode, NNPolicy(deepcopy) opponent, RolloutBuffer; update_networks stubbed), "
                      f"fps as the reference computes it (ppo.py:126-129), on {os.cpu_count()} host CPUs"}


def run_reference_arm(args, wl: Workload):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    envs = wl.envs
    arm = make_cpu_arm(wl, envs)
    # bound the whole run to a few minutes: calibrate a few steps, then sample envs if needed
    probe = cpu_env_run(arm, None, 1, 3, 60.0)
    per_step = probe["seconds"] / max(probe["steps"], 1)
    budget = 150.0
    total = args.steps + args.warmup
    if per_step * total > budget:
        envs = max(1024, int(envs * budget / (per_step * total)) // 1024 * 1024)
        arm = make_cpu_arm(wl, envs)
    res = cpu_env_run(arm, None, args.warmup, args.steps, 1e9)
    value = res["value"]
    line = {
        "impl": "reference", "metric": wl.metric, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": res["steps"],
        "warmup": args.warmup, "ms_per_step": 1e3 * res["seconds"] / res["steps"], "higher_is_better": True,
        "scaling": wl.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": wl.config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": res["threads"], "kind": arm.kind,
                         "sample": f"{res['steps']} steps x {envs} envs of the workload through {arm.what}: actions drawn by the "
                                   "reference's RandomPolicy outside the timed sections, env.step + env.reset(done) timed, on "
                                   f"{os.cpu_count()} host CPUs"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "host": {"cpu_count": os.cpu_count(), "torch_threads": res["threads"], "torch": torch.__version__, "envs_timed": envs},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# B200 arm: env workloads (cfg2 / cfg4 / cfg5)
# ------------------------------------------------------------------------------------------------
def bench_env_workload(ctx: Ctx, wl: Workload, args, primary: bool):
    torch = ctx.torch
    from mnk_b200 import TorchVectorMnkEnv, _lib
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    envs, K, W = wl.envs, args.steps, args.warmup
    total = W + K
    R = max(1, args.replays)
    env = TorchVectorMnkEnv(wl.m, wl.n, wl.k, envs, device=f"cuda:{ctx.local_rank}", env_offset=rank * envs)
    env.reset()
    for t in range(MIX_PLIES):                      # stationary mix of game depths
        env.step_autoreset(env.random_legal_actions(SEED, t), materialise=False)
    snap_bits, snap_meta = env._bits.clone(), env._meta.clone()

    # ---- pre-generate the action trace by playing the games once (untimed) --------------------------
    want_cpu = world == 1 and rank == 0 and not args.no_cpu_baseline
    cpu_steps = args.cpu_steps if primary else min(args.cpu_steps, 12)
    cpu_envs = envs if primary else min(envs, 32768)         # secondary lines: a bounded sample of the shard
    gen = max(total, W + cpu_steps) if want_cpu else total
    actions = torch.empty((gen, envs), dtype=torch.long, device=dev)
    stats = torch.zeros(3, dtype=torch.float64, device=dev)     # episodes, wins, plies
    want_digest = want_last = None
    for t in range(gen):
        env.random_legal_actions(SEED, MIX_PLIES + t, out=actions[t])
        _, r, d = env.step_autoreset(actions[t], materialise=False)
        if W <= t < total:
            stats += torch.stack([d.sum(), r.sum(), torch.tensor(float(envs), device=dev)]).double()
        if t == total - 1:
            want_digest = env.state_checksum()              # state after exactly W + K steps
            want_last = (r.cpu(), d.cpu())                  # ... and that step's rewards / dones

    def restore():
        env._bits.copy_(snap_bits)
        env._meta.copy_(snap_meta)

    # ---- output ring larger than L2 --------------------------------------------------------------
    per_set = envs * 9 * wl.cells
    ring = max(2, -(-3 * L2_BYTES // per_set))
    obs_ring = [torch.empty((envs, 2, wl.m, wl.n), dtype=torch.float32, device=dev) for _ in range(ring)]
    mask_ring = [torch.empty((envs, wl.cells), dtype=torch.bool, device=dev) for _ in range(ring)]
    rewards = torch.empty(envs, dtype=torch.float32, device=dev)
    dones = torch.empty(envs, dtype=torch.bool, device=dev)
    L = _lib.lib()
    flags = _lib.STEP_AUTORESET | (0 if args.no_pdl else _lib.STEP_PDL)

    def launch(t, stream, views=True):
        rc = L.mnk_step(env._stp, actions[t].data_ptr(), None, envs, rewards.data_ptr(), dones.data_ptr(),
                        obs_ring[t % ring].data_ptr() if views else None, mask_ring[t % ring].data_ptr() if views else None,
                        None, flags, stream)
        _lib.check(rc, "mnk_step")

    # warm-up: eager launches of the W warm-up steps
    restore()
    stream = torch.cuda.current_stream().cuda_stream
    for t in range(W):
        launch(t, stream)
    torch.cuda.synchronize()
    warm_bits, warm_meta = env._bits.clone(), env._meta.clone()

    def to_warm():
        env._bits.copy_(warm_bits)
        env._meta.copy_(warm_meta)

    side = torch.cuda.Stream()

    def capture(views):
        g = torch.cuda.CUDAGraph()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                for t in range(W, total):
                    launch(t, side.cuda_stream, views)
        torch.cuda.current_stream().wait_stream(side)
        g.replay()                                  # one untimed replay (graph upload, icache)
        torch.cuda.synchronize()
        return g

    def timed_replays(g, reps):
        """reps replays, each its own event pair (restore between, outside the pair); also one pair around everything."""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        e_all0, e_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        to_warm()
        ctx.barrier()
        e_all0.record()
        for i in range(reps):
            if i:
                to_warm()
            evs[i][0].record()
            g.replay()
            evs[i][1].record()
        e_all1.record()
        ctx.barrier()
        per = ctx.reduce([a.elapsed_time(b) for a, b in evs], "max")
        whole = ctx.reduce([e_all0.elapsed_time(e_all1)], "max")[0]
        return per, whole

    # ---- value: R replays of the K-step graph ------------------------------------------------------
    graph = capture(True)
    sampler = ClockSampler(ctx.local_rank).start()
    per_replay, whole_ms = timed_replays(graph, R)
    verified = env.state_checksum() == want_digest
    ms = statistics.median(per_replay)

    # eager launches of the same K steps (no graph), for the record
    to_warm()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    ev0.record()
    for t in range(W, total):
        launch(t, stream)
    ev1.record()
    ctx.barrier()
    ms_eager = ctx.reduce([ev0.elapsed_time(ev1)], "max")[0]
    verified = verified and env.state_checksum() == want_digest

    # ---- slab launches (for the record): the same K steps as ceil(K / 16) launches of mnk_step_slab ----------
    ms_slab = None
    if primary and actions.is_contiguous():
        import ctypes
        slab_rd = torch.empty((K, 5 * envs + 8), dtype=torch.uint8, device=dev)
        slab_calls = []
        for t0 in range(W, total, 16):
            cnt = min(16, total - t0)
            PtrArr = ctypes.c_void_p * cnt
            slab_calls.append((actions[t0].data_ptr(), slab_rd[t0 - W].data_ptr(), cnt,
                               PtrArr(*[obs_ring[t % ring].data_ptr() for t in range(t0, t0 + cnt)]),
                               PtrArr(*[mask_ring[t % ring].data_ptr() for t in range(t0, t0 + cnt)])))
        g_slab = torch.cuda.CUDAGraph()
        side.wait_stream(torch.cuda.current_stream())
        to_warm()
        with torch.cuda.stream(side):
            with torch.cuda.graph(g_slab, stream=side):
                for a_ptr, rd_ptr, cnt, po, pm in slab_calls:
                    _lib.check(L.mnk_step_slab(env._stp, a_ptr, actions.stride(0) * 8, rd_ptr, slab_rd.stride(0), cnt, po, pm, flags,
                                               side.cuda_stream), "mnk_step_slab")
        torch.cuda.current_stream().wait_stream(side)
        g_slab.replay()
        torch.cuda.synchronize()
        per_s, _ = timed_replays(g_slab, min(R, 10))
        ms_slab = statistics.median(per_s)
        verified = verified and env.state_checksum() == want_digest

    # ---- packed mode: the same K steps without materialising observation / mask -----------
    ms_packed = None
    if primary:
        graph_p = capture(False)
        per_p, _ = timed_replays(graph_p, min(R, 10))
        ms_packed = statistics.median(per_p)
        verified = verified and env.state_checksum() == want_digest

    # ---- e2e: host buffers through the public API -----------------------------------------------------
    e2e_steps = min(K, args.e2e_steps)
    host_actions = torch.empty((W + e2e_steps, envs), dtype=torch.long).pin_memory()
    host_actions.copy_(actions[: W + e2e_steps])
    host_out = torch.empty((e2e_steps, 5 * envs), dtype=torch.uint8).pin_memory()
    host_one = torch.empty(5 * envs, dtype=torch.uint8).pin_memory()

    def warm_to_host_state():
        restore()
        for t in range(W):
            launch(t, stream)

    def e2e_loop(slab, reps, host_obs=None, host_mask=None, steps=e2e_steps, host_actions=host_actions):
        """K steps per C call (mnk_step_host_loop); wall clock AND device events, the larger counts."""
        out = []
        for _ in range(reps):
            warm_to_host_state()
            ctx.barrier()
            t0 = time.perf_counter()
            ev0.record()
            r_h, d_h = env.step_host_loop(host_actions[W:W + steps], host_out, slab_steps=slab, autoreset=True,
                                          ring=(obs_ring, mask_ring), host_obs=host_obs, host_mask=host_mask)
            ev1.record()
            wall = 1e3 * (time.perf_counter() - t0)
            ctx.barrier()
            out.append(max(ev0.elapsed_time(ev1), wall))
        good = env.state_checksum() == want_digest if steps == K else True
        return ctx.reduce(out, "max"), good, (r_h, d_h)

    def e2e_single(zero_copy):
        warm_to_host_state()
        ctx.barrier()
        t0 = time.perf_counter()
        ev0.record()
        for t in range(W, W + e2e_steps):
            env.step_host(host_actions[t], host_one, autoreset=True, out=(obs_ring[t % ring], mask_ring[t % ring]),
                          zero_copy=zero_copy)
        ev1.record()
        wall = 1e3 * (time.perf_counter() - t0)
        ctx.barrier()
        good = env.state_checksum() == want_digest if e2e_steps == K else True
        return ctx.reduce([max(ev0.elapsed_time(ev1), wall)], "max")[0], good

    e2e_reps = 5 if primary else 3
    slabs = sorted({s for s in (args.e2e_slab, 2, 4, 8, 16) if s <= max(e2e_steps, 1)}) if primary else [args.e2e_slab]
    loop_ms = {}
    for s_ in slabs:
        per, ok_l, (r_h, d_h) = e2e_loop(s_, e2e_reps)
        loop_ms[s_] = statistics.median(per)
        verified = verified and ok_l
    # the loop's host results must be the device results of the same steps (last step cross-check)
    if e2e_steps == K:
        verified = verified and bool(torch.equal(r_h[-1], want_last[0]) and torch.equal(d_h[-1], want_last[1]))
    best_slab = min(loop_ms, key=loop_ms.get)
    e2e_ms = loop_ms[best_slab]
    single_zc_ms = single_copy_ms = full_ms = None
    full_steps = 0
    if primary:
        single_copy_ms, ok1 = e2e_single(False)
        single_zc_ms, ok2 = e2e_single(True)
        verified = verified and ok1 and ok2
        full_steps = min(e2e_steps, 8)
        host_obs = torch.empty((full_steps, envs, 2, wl.m, wl.n), dtype=torch.float32).pin_memory()
        host_mask = torch.empty((full_steps, envs, wl.cells), dtype=torch.bool).pin_memory()
        per, _, _ = e2e_loop(max(1, min(4, ring // 2)), 2, host_obs, host_mask, steps=full_steps)
        full_ms = min(per)
        host_actions32 = host_actions.to(torch.int32).pin_memory()       # the reference accepts int32 actions too
        per, ok32, _ = e2e_loop(best_slab, e2e_reps, host_actions=host_actions32)
        i32_ms = statistics.median(per)
        verified = verified and ok32
    clocks = sampler.stop()

    st = ctx.reduce(stats.tolist(), "sum")          # end-of-run statistics over NCCL
    verified = ctx.reduce([1.0 if verified else 0.0], "min")[0] == 1.0
    if rank != 0:
        return None
    peak, peak_src = measured_peak("hbm_gbs")
    total_envs = envs * world
    value = total_envs * K / (ms * 1e-3)
    launch_us = 1e3 * ms / K
    achieved = wl.alg_bytes * envs / (launch_us * 1e-6) / 1e9
    seq_bytes = R * K * wl.alg_bytes * envs
    lower = max(0.0, seq_bytes - L2_BYTES) / (whole_ms * 1e-3) / 1e9
    traffic, traffic_src = profiled_traffic(wl)
    cfg = wl.config(world)
    line = {
        "metric": wl.metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": wl.scaling, "vs_baseline": None,
        "dtype": "u64", "data": "synthetic", "config": cfg,
        "method": {"launch": f"one CUDA graph of {K} step_dense_kernel nodes{'' if args.no_pdl else ' with programmatic dependent launch edges'}, "
                             f"replayed {R} times (state restored between replays, outside the timed span); value = median replay",
                   "eager_launches_value": total_envs * K / (ms_eager * 1e-3)},
        "slab_launches": None if ms_slab is None else {
            "value": total_envs * K / (ms_slab * 1e-3), "unit": UNIT, "launches": -(-K // 16),
            "hbm_frac": wl.alg_bytes * envs / (1e3 * ms_slab / K * 1e-6) / 1e9 / measured_peak("hbm_gbs")[0],
            "note": "for the record, NOT the headline: the same K steps (same actions, same per-step outputs) as ceil(K / 16) launches "
                    "of mnk_step_slab -- a CTA loops over its tile's steps, so there is no per-step launch fill / drain; only "
                    "usable when the actions of several steps are known in advance (mnk_step_host_loop uses it)"},
        "timing": {"replays": R, "ms_median": ms, "ms_min": min(per_replay), "ms_max": max(per_replay), "ms_first": per_replay[0],
                   "value_at_min": total_envs * K / (min(per_replay) * 1e-3), "value_at_max": total_envs * K / (max(per_replay) * 1e-3),
                   "whole_sequence_ms": whole_ms},
        "verified_state_digest": verified,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "frac_lower_bound": lower / peak,
                     "frac_lower_bound_note": f"({R} replays x {K} launches x algorithmic bytes - one L2 capacity of possibly-dirty lines) / the "
                                              "time of the WHOLE replay sequence incl. state restores",
                     "traffic": traffic, "traffic_source": traffic_src,
                     "kernel": f"step_dense_kernel<SGeom<{wl.m},{wl.n},{wl.k}>>",
                     "alg_bytes_per_env_step": wl.alg_bytes, "launch_us": launch_us, "peak_source": peak_src},
        "e2e": {"value": total_envs * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 8 * envs,
                "d2h_bytes_per_step": 5 * envs, "steps": e2e_steps, "slab_steps": best_slab,
                "api": "TorchVectorMnkEnv.step_host_loop -> mnk_step_host_loop: all K steps in one C call, pinned int64 actions in, f32 rewards "
                       "+ bool dones out (bytes are per GPU per step), copies pipelined per slab on side streams (copy-in a slab ahead), one "
                       "kernel launch (mnk_step_slab) and one host wait per slab; "
                       "observation + mask are materialised every step and stay on the device (their consumer is the GPU network) -- see "
                       "full_d2h for a host consumer",
                "by_slab_steps": {str(s_): total_envs * e2e_steps / (v * 1e-3) for s_, v in loop_ms.items()}},
        "gpu_launches": K,
        "clocks": clocks,
        "stats": {"episodes": st[0], "wins": st[1], "plies": st[2]},
    }
    if ctx.numa is not None:
        line["e2e"]["host_affinity"] = ctx.numa
    if primary:
        line["e2e"]["single_step_api"] = {
            "note": "TorchVectorMnkEnv.step_host -> mnk_step_host: ONE step per call, stream synchronised every step",
            "staged_copies": total_envs * e2e_steps / (single_copy_ms * 1e-3),
            "zero_copy": total_envs * e2e_steps / (single_zc_ms * 1e-3)}
        line["e2e"]["int32_actions"] = {
            "value": total_envs * e2e_steps / (i32_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 4 * envs,
            "note": "same loop fed int32 actions (half the host->device bytes): the int64 loop is bound by the PCIe copies -- an "
                    "8-step slab moves 4 MiB in and 2.5 MiB out, ~100 us with both directions busy (tools/pcie_probe.py: 42 + 26 "
                    "GB/s concurrently) against 70 us of kernels"}
        line["e2e"]["full_d2h"] = {
            "value": total_envs * full_steps / (full_ms * 1e-3), "unit": UNIT, "steps": full_steps,
            "d2h_bytes_per_step": (5 + 9 * wl.cells) * envs,
            "note": "same loop, every step's f32 observation + bool mask ALSO copied to pinned host memory: PCIe-bound "
                    f"({9 * wl.cells} B/env-step)"}
        line["packed_mode"] = {"value": total_envs * K / (ms_packed * 1e-3), "unit": UNIT, "launch_us": 1e3 * ms_packed / K,
                               "alg_bytes_per_env_step": wl.packed_bytes,
                               "hbm_frac": wl.packed_bytes * envs / (1e3 * ms_packed / K * 1e-6) / 1e9 / peak,
                               "note": "same K steps, observation / mask not materialised (state + action + reward / done "
                                       "traffic only); latency-bound, reported for SURVEY 8d, not the headline"}
    if want_cpu:
        acts_cpu = actions[: W + cpu_steps, :cpu_envs].cpu()

        def check_state(arm, n_steps):
            # cross-check: the CPU arm after n_steps of the SAME trace == the CUDA env after n_steps
            restore()
            for t in range(n_steps):
                launch(t, stream)
            torch.cuda.synchronize()
            planes, plies = arm.state()
            same = torch.equal(env.boards[:cpu_envs].cpu(), planes) and torch.equal(env.move_counts[:cpu_envs].cpu(), plies)
            env.release_mirrors()
            return bool(same)

        restore()
        init = (env.boards[:cpu_envs].cpu(), env.current_player[:cpu_envs].cpu(), env.move_counts[:cpu_envs].cpu())
        env.release_mirrors()
        arm = make_cpu_arm(wl, cpu_envs)
        res = cpu_env_run(arm, acts_cpu, min(W, 3), cpu_steps - min(W, 3), args.cpu_budget if primary else 10.0, check_state, init)
        line["cpu_baseline"] = {
            "value": res["value"], "unit": UNIT, "cores": res["threads"], "kind": arm.kind,
            "sample": f"first {res['steps']} steps of the same action trace, {'all' if cpu_envs == envs else 'the first'} {cpu_envs} envs, "
                      f"{arm.what} on {os.cpu_count()} host CPUs",
            "parity_with_gpu_state": res["parity"]}
        if primary:
            # informative second baseline: the same code on this B200 (device="cuda", stock PyTorch kernels)
            gsteps = min(cpu_steps, 40)
            garm = make_cpu_arm(wl, envs, device=dev)
            gres = cpu_env_run(garm, actions, min(W, 3), gsteps - min(W, 3), 60.0, None,
                               tuple(x.to(dev) for x in init), device=dev)
            line["cpu_baseline"]["stock_torch_same_gpu"] = {
                "value": gres["value"], "unit": UNIT, "steps": gres["steps"], "kind": garm.kind,
                "note": f"{garm.what} with device='cuda' (stock PyTorch kernels on this B200), same trace, host-synchronised "
                        "per step like the reference's loop"}
    del obs_ring, mask_ring, actions
    torch.cuda.empty_cache()
    return line


# ------------------------------------------------------------------------------------------------
# B200 arm: cfg3, self-play rollout with the policy/value network
# ------------------------------------------------------------------------------------------------
def bench_rollout(ctx: Ctx, args, envs: int, agent_bn: str, extras: bool = True):
    """BASELINE cfg3 per GPU: 9x9x5, agent and opponent both resnet_b_s (same random-init weights, opponent frozen),
    tcgen05 forward fed from bitboards, Gumbel-max sampling, fused wrapper, packed PPO buffer, on-device episode
    statistics; K rollout steps timed.  Metric = the reference's fps (ppo.py:126-129): agent steps / s.
    agent_bn = "train": the agent's forward uses batch statistics and updates the running ones, as the reference's
    rollout does (ppo.py:97, the network is in train mode); "eval": frozen statistics (the opponent's mode)."""
    import copy
    torch = ctx.torch
    from mnk_b200 import (NativeNNPolicy, NativeResNet, ResNetActorCritic, RolloutBuffer, RolloutCollector,
                          TorchSelfPlayWrapper, TorchVectorMnkEnv)
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    m, n, k = 9, 9, 5
    cells = m * n
    K, W = args.steps, args.warmup
    torch.manual_seed(0)
    net = ResNetActorCritic((2, m, n), cells).to(dev)
    net.train(agent_bn == "train")
    agent = NativeResNet(net, device=dev, bn_mode=agent_bn)
    agent_mode = agent.bn_mode
    opponent = NativeNNPolicy(copy.deepcopy(net), device=dev, seed=7)
    env = TorchVectorMnkEnv(m, n, k, envs, device=f"cuda:{ctx.local_rank}", env_offset=rank * envs)
    wr = TorchSelfPlayWrapper(env, seed=SEED)
    wr.set_opponent(opponent)
    col = RolloutCollector(envs, device=dev, seed=11, row_offset=rank * envs, world_size=world)
    wr.reset(materialise=False)
    col._last_obs = {"observation": None, "action_mask": None}
    warm_buf = RolloutBuffer(max(W, MIX_PLIES // 2), envs, (2, m, n), cells, device=dev, k=k)
    col.collect(agent, wr, warm_buf)                       # warm-up: also brings games to a stationary depth mix
    use_graph = not args.no_graph
    buf = RolloutBuffer(K, envs, (2, m, n), cells, device=dev, k=k)
    if use_graph:                                          # capture the K-step rollout once; its first replay is untimed
        col.collect(agent, wr, buf, graph=True)
        buf.reset()
    sampler = ClockSampler(ctx.local_rank).start()
    reps = max(1, min(args.replays, 5))
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    walls, stats = [], None
    for i in range(reps):
        buf.reset()
        ctx.barrier()
        t0 = time.perf_counter()
        ev[i][0].record()
        stats = col.collect(agent, wr, buf, graph=use_graph)   # K steps; ends with the NCCL all-reduce + one host read
        ev[i][1].record()
        ctx.barrier()
        walls.append(1e3 * (time.perf_counter() - t0))
    clocks = sampler.stop()
    per = ctx.reduce([a.elapsed_time(b) for a, b in ev], "max")
    walls = ctx.reduce(walls, "max")
    ms, wall_ms = statistics.median(per), statistics.median(walls)
    if rank != 0:
        del buf, warm_buf, col, wr, env, agent, opponent
        torch.cuda.empty_cache()
        if extras and agent_mode == "train":          # every rank takes part in the eval-mode variant's barriers
            bench_rollout(ctx, args, envs, "eval", extras=False)
        return None
    peak, peak_src = measured_peak("bf16_tflops_sustained")
    flop_per_agent_step = 2 * FLOP_PER_FORWARD[(m, n)]      # agent + opponent forward
    total = envs * world * K
    value = total / (ms * 1e-3)
    achieved = value * flop_per_agent_step / 1e12 / world   # per GPU, against the per-GPU peak
    launches = 19 if agent_mode == "train" else 10       # kernels per agent step (the train-mode tower is 9 layer launches + 1)
    line = {
        "metric": "self-play rollout steps/sec (9x9x5, resnet_b_s agent + opponent)", "value": value, "unit": "agent-steps/s",
        "n_gpus": world, "steps": K, "warmup": warm_buf.n_steps, "ms_per_step": ms / K, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"cfg3: gomoku 9x9x5 self-play rollout, {envs} envs/GPU x {K} steps, resnet_b_s agent (BatchNorm in {agent_mode} "
                               "mode) + frozen copy as opponent (eval mode), random-init weights, tcgen05 forward from bitboards, Gumbel-max "
                               "sampling, fused wrapper, packed PPO buffer, on-device episode stats"
                               + (", the whole K-step rollout replayed as one CUDA graph" if use_graph else ", eager launches"),
                   "envs_per_gpu": envs, "global_envs": envs * world, "agent_batchnorm": agent_mode,
                   "l2": f"per step the towers stream {envs * 72 / 2**20:.1f} MiB of bitboards and {envs * 972 * 2 / 2**20:.0f} MiB of "
                         "head features; rollout buffer slots are distinct per step",
                   "parallelism": f"env-shard x{world}, one NCCL all-reduce of 6 doubles per rollout"},
        "timing": {"replays": reps, "ms_median": ms, "ms_min": min(per), "ms_max": max(per)},
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "traffic": None,
                     "kernel": ("resnet_layer_train_kernel x 9 + resnet_train_features_kernel (agent, batch-statistics BatchNorm: fp16 "
                                "pre-activations round-trip through HBM between the layer launches, ~156 KB per env and forward) + "
                                "resnet_tower_rows_kernel (opponent)" if agent_mode == "train" else
                                "resnet_tower_rows_kernel (agent + opponent forward per agent-step)"),
                     "flop_per_agent_step": flop_per_agent_step, "peak_source": peak_src, "per_gpu": True},
        "e2e": {"value": total / (wall_ms * 1e-3), "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 64.0 / K,
                "api": "RolloutCollector.collect (wall clock incl. the statistics all-reduce and host read)"},
        "gpu_launches": launches * K,
        "clocks": clocks,
        "stats": {"episodes": stats.episodes, "mean_reward": stats.mean_reward, "mean_length": stats.mean_length,
                  "wins": stats.wins, "losses": stats.losses, "draws": stats.draws},
    }
    del buf, warm_buf, col, wr, env, agent, opponent
    torch.cuda.empty_cache()
    if not extras:
        return line
    if agent_mode == "train":
        # the same rollout with the agent's BatchNorm frozen (eval mode: ONE fused tower kernel instead of one launch per
        # layer) -- NOT what the reference's rollout computes (ppo.py:97 stays in train mode); reported for the record
        ev = bench_rollout(ctx, args, envs, "eval", extras=False)
        if ev is not None:
            line["eval_mode_agent_variant"] = {
                "value": ev["value"], "unit": ev["unit"], "ms_per_step": ev["ms_per_step"], "roofline_frac": ev["roofline"]["frac"],
                "note": "agent BatchNorm in eval mode (frozen statistics, whole tower fused into one kernel); the reference's "
                        "rollout forward uses batch statistics, so the headline of this line is the train-mode figure"}
    if world == 1 and not args.no_cpu_baseline:
        wl = Workload("cfg2", 1)
        base = cpu_rollout_run(wl, args.cpu_rollout_envs, args.cpu_rollout_steps, 20.0)
        if base is not None:
            line["cpu_baseline"] = base
    return line


# ------------------------------------------------------------------------------------------------
# B200 arm: cfg1, the reference's own CPU-runnable case through the self-play wrapper
# ------------------------------------------------------------------------------------------------
def bench_wrapper_cfg1(ctx: Ctx, args):
    """BASELINE configs[0] / SURVEY 8d cfg1: 3x3x3, 1,024 envs per GPU, RandomPolicy as agent AND opponent through
    TorchSelfPlayWrapper (reset of finished games, side draw, opponent turn, zero-sum rewards, canonical obs + mask), 300
    steps after warm-up.  Here: one sampler launch (the agent's RandomPolicy.act on the returned mask) + ONE fused wrapper
    launch per step (mnk_selfplay_step_random), eager and as one CUDA graph; the CPU figure is the unmodified reference
    running the same loop on the host.  Launch-bound at this size by construction -- the line shows the wrapper path at the
    reference's default scale, not a roofline."""
    torch = ctx.torch
    from mnk_b200 import RandomPolicy, TorchSelfPlayWrapper, TorchVectorMnkEnv
    m, n, k, envs, steps = 3, 3, 3, 1024, 300
    env = TorchVectorMnkEnv(m, n, k, envs, device=f"cuda:{ctx.local_rank}", env_offset=ctx.rank * envs)
    wr = TorchSelfPlayWrapper(env, seed=SEED)
    wr.set_opponent(RandomPolicy(m * n, seed=5))
    agent = RandomPolicy(m * n, seed=6)
    obs, _ = wr.reset()
    done_total = torch.zeros((), dtype=torch.float64, device=ctx.dev)

    def one_step(o):
        o, r, term, trunc, _ = wr.step(agent.act(o))
        done_total.add_(term.sum())
        return o

    for _ in range(10):
        obs = one_step(obs)
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        obs = one_step(obs)
    e1.record()
    ctx.barrier()
    ms_eager = ctx.reduce([e0.elapsed_time(e1)], "max")[0]
    finished = float(done_total.item())
    # the same `steps` steps captured once as a CUDA graph (every launch of the path is capturable: no host synchronisation)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            o = obs
            for _ in range(steps):
                o = one_step(o)
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    ctx.barrier()
    per = []
    for _ in range(5):
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        per.append(e0.elapsed_time(e1))
    ms_graph = ctx.reduce([statistics.median(per)], "max")[0]
    line = {"metric": "self-play wrapper steps/sec (3x3x3, random vs random)", "value": envs * ctx.world * steps / (ms_eager * 1e-3),
            "cuda_graph": {"value": envs * ctx.world * steps / (ms_graph * 1e-3), "unit": "agent-steps/s", "us_per_step": 1e3 * ms_graph / steps,
                           "note": "the same loop replayed as one CUDA graph (what RolloutCollector.collect(graph=True) does): the eager "
                                   "figure is bound by ~70 us of Python per step"},
            "unit": "agent-steps/s", "n_gpus": ctx.world, "steps": steps, "warmup": 10, "ms_per_step": ms_eager / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": "cfg1: tic-tac-toe 3x3x3, 1,024 envs per GPU, RandomPolicy agent and opponent through TorchSelfPlayWrapper "
                                   "(f32 canonical observation + bool mask materialised every step), eager launches",
                       "envs_per_gpu": envs, "l2": "the whole working set (37 KB of outputs per step) is cache-resident: launch-bound"},
            "gpu_launches": 2 * steps, "episodes_finished": finished}
    if ctx.rank == 0 and ctx.world == 1 and not args.no_cpu_baseline:
        from oracle import ref_tree
        if ref_tree.available():
            env_mod, wrap_mod, pol_mod = ref_tree.load("env.torch_vector_mnk_env", "selfplay.torch_self_play_wrapper", "selfplay.policy")
            torch.set_num_threads(os.cpu_count() or 1)
            torch.manual_seed(0)
            r_wr = wrap_mod.TorchSelfPlayWrapper(env_mod.TorchVectorMnkEnv(m, n, k, envs, device="cpu"))
            r_wr.set_opponent(pol_mod.RandomPolicy(m * n))
            r_agent = pol_mod.RandomPolicy(m * n)
            o, _ = r_wr.reset()
            for _ in range(3):
                o, *_ = r_wr.step(r_agent.act(o))
            t0 = time.perf_counter()
            for _ in range(steps):
                o, *_ = r_wr.step(r_agent.act(o))
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": envs * steps / dt, "unit": "agent-steps/s", "cores": torch.get_num_threads(), "kind": "reference",
                                    "sample": f"the same loop, {steps} steps x {envs} envs, through the unmodified reference (env, wrapper, "
                                              f"RandomPolicy) with device='cpu' on {os.cpu_count()} host CPUs"}
    del wr, env
    return line if ctx.rank == 0 else None


# ------------------------------------------------------------------------------------------------
# B200 arm: the network forwards alone (every tcgen05 forward the package ships), samples/s
# ------------------------------------------------------------------------------------------------
def bench_forwards(ctx: Ctx, args, envs: int):
    """`forward_env` (tower + head tails, logits and values from the packed bitboards) of every native network at 9x9 on
    `envs` mid-game positions per GPU: the default resnet_b_s in eval and train mode, the wider convolutional architectures of
    src/alg/architectures/configs.py:36-65 on mnk_conv_tower and the transformers of :7-25 on mnk_transformer_body.  CUDA
    events around 10 forwards after 3 warm-ups, max over ranks; useful FLOPs = 2 * MACs of the convolutions, the head Linear
    layers and the encoder's matmuls (what the reference network computes, padding excluded) against the measured sustained bf16 peak.  Each forward streams new features (tens of MB) and the
    towers re-read only weights, so no L2 flush is needed between iterations."""
    torch = ctx.torch
    from mnk_b200 import NativeResNet, TorchVectorMnkEnv, build_architecture, native_network
    m, n, k = 9, 9, 5
    cells = m * n
    env = TorchVectorMnkEnv(m, n, k, envs, device=f"cuda:{ctx.local_rank}", env_offset=ctx.rank * envs)
    env.reset()
    for t in range(24):
        env.step_autoreset(env.random_legal_actions(SEED, t), materialise=False)
    peak, peak_src = measured_peak("bf16_tflops_sustained")
    small = min(envs, 4096)                      # the stock-PyTorch comparison runs on a slice (its activations are large)
    env_small = TorchVectorMnkEnv(m, n, k, small, device=f"cuda:{ctx.local_rank}")
    env_small._bits.copy_(env._bits[:, :, :small])
    env_small._meta.copy_(env._meta[:small])
    obs_small = env_small.observe()["observation"]
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    out = {}
    for name, mode in (("resnet_b_s", "eval"), ("resnet_b_s", "train"), ("resnet_b_l", "eval"), ("cnn_b_s", "eval"), ("cnn_b_l", "eval"),
                       ("transformer_b_s", "eval"), ("transformer_b_l", "eval")):
        torch.manual_seed(0)
        net = build_architecture(name, (2, m, n), cells).to(ctx.dev)
        net.train(mode == "train")
        fwd = NativeResNet(net, device=ctx.dev, bn_mode="train") if mode == "train" else native_network(net, device=ctx.dev)
        flops = sum(2 * cells * c.in_channels * c.out_channels * c.kernel_size[0] * c.kernel_size[1]
                    for c in net.modules() if isinstance(c, torch.nn.Conv2d))
        heads = [getattr(net, a) for a in ("policy_head", "value_head", "actor", "critic") if hasattr(net, a)]
        flops += sum(2 * l.in_features * l.out_features for hd in heads for l in hd if isinstance(l, torch.nn.Linear))
        if hasattr(net, "transformer"):     # per token and layer: in_proj + out_proj + feed-forward = 12 D^2 MACs, QK^T + PV = 2 T D
            Dm = net.embed_dim
            flops += net.num_layers * (2 * cells * 12 * Dm * Dm + 2 * 2 * cells * cells * Dm) + 2 * cells * 3 * Dm
        for _ in range(3):
            fwd.forward_env(env)
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            fwd.forward_env(env)
        e1.record()
        ctx.barrier()
        ms = ctx.reduce([e0.elapsed_time(e1) / reps], "max")[0]
        fwd.check_error()
        tf = envs * flops / (ms * 1e-3) / 1e12
        # the same forward through the stock module on the slice (TF32 allowed, as the reference configures: utils/hardware.py:35-37)
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = True
        with torch.no_grad():
            for _ in range(2):
                net(obs_small, None)
            e0.record()
            for _ in range(3):
                net(obs_small, None)
            e1.record()
        torch.cuda.synchronize()
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
        stock_sps = small / (e0.elapsed_time(e1) / 3 * 1e-3)
        out[f"{name}_{mode}"] = {"samples_per_s": envs * ctx.world / (ms * 1e-3), "ms_per_forward": ms, "mflop_per_sample": flops / 1e6,
                                 "useful_tflops_per_gpu": tf, "frac_of_bf16_peak": tf / peak,
                                 "kernel": type(fwd).__name__ + (" (mnk_resnet_tower_train)" if mode == "train" else ""),
                                 "stock_torch_samples_per_s_per_gpu": stock_sps, "x_stock_torch": envs / (ms * 1e-3) / stock_sps}
        del fwd, net
    del env, env_small, obs_small
    torch.cuda.empty_cache()
    return {"metric": "network forward samples/sec (9x9, logits + value from packed bitboards)", "unit": "samples/s", "n_gpus": ctx.world,
            "envs_per_gpu": envs, "dtype": "fp16 operands, fp32 accumulate", "peak_tflops": peak, "peak_source": peak_src,
            "stock_torch": f"the same module's forward (fp32, TF32 allowed; train mode for the *_train entry) on {small} of the positions",
            "forwards": out}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(Workload.TABLE) + ["cfg3"],
                    help="run ONE workload as the printed line; default: cfg2 as the line + cfg3 / cfg4 / cfg5 under `secondary`")
    ap.add_argument("--envs", type=int, default=None, help="envs per GPU (default: the workload's)")
    ap.add_argument("--replays", type=int, default=20, help="timed replays of the K-step graph (median reported)")
    ap.add_argument("--e2e-steps", type=int, default=500)
    ap.add_argument("--e2e-slab", type=int, default=4, help="steps per slab of mnk_step_host_loop")
    ap.add_argument("--cpu-steps", type=int, default=60)
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--cpu-rollout-envs", type=int, default=1024)
    ap.add_argument("--cpu-rollout-steps", type=int, default=4)
    ap.add_argument("--rollout-envs", type=int, default=32768, help="cfg3: envs per GPU (262,144 over 8 GPUs)")
    ap.add_argument("--agent-bn", default="train", choices=["train", "eval"], help="cfg3: BatchNorm mode of the agent's rollout forward")
    ap.add_argument("--no-secondary", action="store_true", help="print the cfg2 line without the cfg3 / cfg4 / cfg5 lines")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="cfg3: eager launches instead of one CUDA graph per rollout")
    ap.add_argument("--no-pdl", action="store_true", help="plain launches instead of programmatic dependent launch")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference_arm(args, Workload(args.workload if args.workload in Workload.TABLE else "cfg2", args.gpus, args.envs))
    ctx = Ctx()
    try:
        if args.workload == "cfg3":
            line = bench_rollout(ctx, args, args.envs or args.rollout_envs, args.agent_bn)
        elif args.workload is not None:
            line = bench_env_workload(ctx, Workload(args.workload, ctx.world, args.envs), args, primary=True)
        else:
            line = bench_env_workload(ctx, Workload("cfg2", ctx.world, args.envs), args, primary=True)
            if not args.no_secondary:
                secondary = {}
                for name in ("cfg1", "cfg3", "cfg4", "cfg5", "forwards"):
                    try:
                        if name == "cfg1":
                            sec = bench_wrapper_cfg1(ctx, args)
                        elif name == "cfg3":
                            sec = bench_rollout(ctx, args, args.rollout_envs, args.agent_bn)
                        elif name == "forwards":
                            sec = bench_forwards(ctx, args, args.rollout_envs)
                        else:
                            sec = bench_env_workload(ctx, Workload(name, ctx.world), args, primary=False)
                    except Exception as exc:      # a secondary line must never take the headline down with it
                        sec = {"error": f"{type(exc).__name__}: {exc}"[:400]}
                        if ctx.world > 1:
                            raise
                    secondary[name] = sec
                if line is not None:
                    line["secondary"] = secondary
        if ctx.rank == 0 and line is not None:
            print(json.dumps(line), flush=True)
    finally:
        ctx.close()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- headline benchmark of the MNK hot path (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (cfg2): Gomoku 9x9x5, 65,536 envs per GPU, seeded uniformly-random legal actions,
one "step" = TorchVectorMnkEnv.step over the whole batch (stone placement, win/draw check,
rewards, dones, player toggle, API-exact f32 observation + bool mask materialised) fused with
the reset of finished envs.  Envs are independent, so ranks hold disjoint shards (weak scaling,
no data-path collective; one NCCL all-reduce of end-of-run statistics).

Measurement protocol
  * The games are first advanced 64 plies so that envs sit at a stationary mix of depths, the
    packed state is snapshotted, and the W+K action batches are pre-generated ON DEVICE by
    replaying the games once (untimed).  The state is then restored.
  * `value`: the K timed steps are K kernel nodes of one CUDA graph (inputs resident in HBM),
    bracketed by barrier + synchronize, timed with CUDA events, max over ranks.  Outputs rotate
    through a ring of buffers larger than L2.  The final state digest must equal the digest
    reached during trace generation (the timed region demonstrably did the work).
  * `e2e`: the same K steps through the public host-buffer API (env.step_host -> mnk_step_host):
    per step one H2D copy of the pinned int64 actions, the launch, one D2H copy of
    rewards + dones and a stream synchronise.
  * `roofline`: algorithmic bytes per launch (814 B/env-step API-exact at 9x9)
    / the average launch duration inside the timed region, against MEASURED_PEAKS.json.
  * `cpu_baseline` (rank 0, N=1): oracle/torch_port.py -- the reference's env step restated with
    the same torch ops -- replaying the first steps of the SAME action trace on the host cores.
  * `--impl reference`: that CPU port alone, as the reference arm.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "rl-selfplay-mnk_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "env steps/sec (9x9x5, win-check)"
UNIT = "env-steps/s"
MIX_PLIES = 64
L2_BYTES = 126 * 1024 * 1024

# name -> (m, n, k, envs per GPU at N GPUs, BASELINE.json config it stands for)
WORKLOADS = {
    "cfg2": (9, 9, 5, lambda n: 65536, "cfg2: gomoku 9x9x5, 65536 envs/GPU", "weak"),
    "cfg4": (13, 13, 5, lambda n: 1048576 // n, "cfg4: 13x13x5, 1,048,576 envs sharded over the GPUs", "strong"),
    "cfg5": (19, 19, 5, lambda n: 4194304 // 8, "cfg5: 19x19x5, 524,288 envs/GPU (4,194,304 over 8 GPUs)", "weak"),
}
SCALING = "weak"
M = N_COLS = K_LINE = CELLS = ENVS_PER_GPU = ALG_BYTES_PER_ENV_STEP = PACKED_BYTES_PER_ENV_STEP = 0
WORKLOAD_DESC = ""


def configure(name: str, gpus: int, envs_override=None):
    """Sets the board geometry, per-GPU env count and the algorithmic bytes per env-step (SURVEY 8d:
    8 B action + 2 x packed state + f32 observation + bool mask + reward/done)."""
    global M, N_COLS, K_LINE, CELLS, ENVS_PER_GPU, ALG_BYTES_PER_ENV_STEP, PACKED_BYTES_PER_ENV_STEP, WORKLOAD_DESC, METRIC, SCALING
    M, N_COLS, K_LINE, envs_fn, WORKLOAD_DESC, SCALING = WORKLOADS[name]
    CELLS = M * N_COLS
    ENVS_PER_GPU = envs_override or envs_fn(gpus)
    words = (M * (N_COLS + 1) + 63) // 64
    state_bytes = 2 * 8 * words + 4
    ALG_BYTES_PER_ENV_STEP = 8 + 2 * state_bytes + 8 * CELLS + CELLS + 5      # 814 at 9x9
    PACKED_BYTES_PER_ENV_STEP = 8 + 2 * state_bytes + 5                       # 85 at 9x9
    METRIC = f"env steps/sec ({M}x{N_COLS}x{K_LINE}, win-check)"


def workload_name(envs):
    return (f"{WORKLOAD_DESC} ({envs} envs on this GPU), seeded random legal actions, env.step (placement + "
            "k-in-a-row win/draw check + rewards/dones + f32 obs & bool mask materialised) + auto-reset")


# ------------------------------------------------------------------------------------------------
# clocks sampler (NVML)
# ------------------------------------------------------------------------------------------------
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
            self._stop.wait(0.004)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def profiled_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "step_dense_traffic.json")) as f:
            return json.load(f).get("dram_bytes_per_launch")
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm (oracle/torch_port.py on the host cores)
# ------------------------------------------------------------------------------------------------
def cpu_port_run(actions_cpu, envs, warmup, steps, budget_s, check_state=None, init_state=None, device="cpu"):
    """Replays `actions_cpu[t]` ([T, envs] int64 legal actions, or None => draw them with the
    reference's RandomPolicy arithmetic outside the timed sections) through the torch-op port.
    Times only port_step + port_reset(done_idx), like the reference harness of SURVEY 8d."""
    import torch
    from oracle import torch_port as tp
    torch.set_num_threads(os.cpu_count() or 1)
    s = tp.port_make(M, N_COLS, K_LINE, envs, device=device)
    obs = tp.port_reset(s)
    if init_state is not None:          # start from the same mid-game positions as the B200 arm
        s.planes.copy_(init_state[0]), s.to_move.copy_(init_state[1]), s.plies.copy_(init_state[2])
        obs = tp.port_observe(s)
    on_gpu = str(device) != "cpu"
    sync = torch.cuda.synchronize if on_gpu else (lambda: None)
    timed, done_steps = 0.0, 0
    t_begin = time.perf_counter()
    total = warmup + steps
    for t in range(total):
        a = actions_cpu[t] if actions_cpu is not None else tp.port_uniform_legal(obs["action_mask"])
        sync()
        t0 = time.perf_counter()
        obs, r, d = tp.port_step(s, a)
        idx = torch.nonzero(d).squeeze(1)
        if idx.numel():
            obs = tp.port_reset(s, idx)
        sync()
        dt = time.perf_counter() - t0
        if t >= warmup:
            timed += dt
            done_steps += 1
        if time.perf_counter() - t_begin > budget_s and done_steps >= 3:
            break
    ok = None
    if check_state is not None:
        ok = check_state(s, warmup + done_steps)
    return {"steps": done_steps, "seconds": timed, "value": envs * done_steps / timed, "threads": torch.get_num_threads(),
            "parity": ok}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    envs = ENVS_PER_GPU
    # bound the whole run to a few minutes: calibrate one step, then sample envs if needed
    probe = cpu_port_run(None, envs, 1, 3, 60.0)
    per_step = probe["seconds"] / max(probe["steps"], 1)
    budget = 150.0
    total = args.steps + args.warmup
    if per_step * total > budget:
        envs = max(1024, int(envs * budget / (per_step * total)) // 1024 * 1024)
    res = cpu_port_run(None, envs, args.warmup, args.steps, 1e9)
    value = res["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": res["steps"],
        "warmup": args.warmup, "ms_per_step": 1e3 * res["seconds"] / res["steps"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(ENVS_PER_GPU), "device": "host CPU", "envs_timed": envs},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": res["threads"], "kind": "port",
                         "sample": f"{res['steps']} steps x {envs} envs of the workload through oracle/torch_port.py "
                                   "(the reference's torch op sequence: index_put, 3x conv2d, threshold, any, clone)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "host": {"cpu_count": os.cpu_count(), "torch_threads": res["threads"], "torch": torch.__version__},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200_arm(args):
    import torch
    import torch.distributed as dist
    from mnk_b200 import TorchVectorMnkEnv, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 arm has no CPU fallback; use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    envs, K, W = args.envs, args.steps, args.warmup
    total = W + K
    env = TorchVectorMnkEnv(M, N_COLS, K_LINE, envs, device=f"cuda:{local_rank}", env_offset=rank * envs)
    env.reset()
    seed = 20261018
    for t in range(MIX_PLIES):                      # stationary mix of game depths
        env.step_autoreset(env.random_legal_actions(seed, t), materialise=False)
    snap_bits, snap_meta = env._bits.clone(), env._meta.clone()

    # ---- pre-generate the action trace by playing the games once (untimed) --------------------------
    # the CPU baseline replays the head of the same trace: generate enough batches for it even when K is small
    want_cpu = world == 1 and not args.no_cpu_baseline
    gen = max(total, W + args.cpu_steps) if want_cpu else total
    actions = torch.empty((gen, envs), dtype=torch.long, device=dev)
    stats = torch.zeros(3, dtype=torch.float64, device=dev)     # episodes, wins, plies
    want_digest = None
    for t in range(gen):
        env.random_legal_actions(seed, MIX_PLIES + t, out=actions[t])
        _, r, d = env.step_autoreset(actions[t], materialise=False)
        if W <= t < total:
            stats += torch.stack([d.sum(), r.sum(), torch.tensor(float(envs), device=dev)]).double()
        if t == total - 1:
            want_digest = env.state_checksum()              # state after exactly W + K steps

    def restore():
        env._bits.copy_(snap_bits)
        env._meta.copy_(snap_meta)

    # ---- output ring larger than L2 --------------------------------------------------------------
    per_set = envs * (8 * CELLS + CELLS)
    ring = max(2, -(-3 * L2_BYTES // per_set))
    obs_ring = [torch.empty((envs, 2, M, N_COLS), dtype=torch.float32, device=dev) for _ in range(ring)]
    mask_ring = [torch.empty((envs, CELLS), dtype=torch.bool, device=dev) for _ in range(ring)]
    rewards = torch.empty(envs, dtype=torch.float32, device=dev)
    dones = torch.empty(envs, dtype=torch.bool, device=dev)
    L = _lib.lib()
    flags = _lib.STEP_AUTORESET | (0 if args.no_pdl else _lib.STEP_PDL)

    def launch(t, stream):
        rc = L.mnk_step(env._stp, actions[t].data_ptr(), None, envs, rewards.data_ptr(), dones.data_ptr(),
                        obs_ring[t % ring].data_ptr(), mask_ring[t % ring].data_ptr(), None, flags, stream)
        _lib.check(rc, "mnk_step")

    # warm-up: eager launches of the W warm-up steps
    restore()
    stream = torch.cuda.current_stream().cuda_stream
    for t in range(W):
        launch(t, stream)
    torch.cuda.synchronize()
    warm_bits, warm_meta = env._bits.clone(), env._meta.clone()

    # the K timed steps as one CUDA graph of K kernel nodes
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for t in range(W, total):
                launch(t, side.cuda_stream)
    torch.cuda.current_stream().wait_stream(side)
    # one untimed replay (graph upload, icache), then restore the post-warm-up state
    graph.replay()
    torch.cuda.synchronize()
    env._bits.copy_(warm_bits)
    env._meta.copy_(warm_meta)

    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    graph.replay()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    got_digest = env.state_checksum()
    verified = got_digest == want_digest

    # eager launches of the same K steps (no graph), for the record
    env._bits.copy_(warm_bits)
    env._meta.copy_(warm_meta)
    barrier()
    ev0.record()
    for t in range(W, total):
        launch(t, stream)
    ev1.record()
    barrier()
    ms_eager = ev0.elapsed_time(ev1)
    verified = verified and env.state_checksum() == want_digest

    # ---- packed mode: the same K steps without materialising observation / mask ---------------
    def launch_packed(t, stream_):
        rc = L.mnk_step(env._stp, actions[t].data_ptr(), None, envs, rewards.data_ptr(), dones.data_ptr(),
                        None, None, None, flags, stream_)
        _lib.check(rc, "mnk_step")

    env._bits.copy_(warm_bits)
    env._meta.copy_(warm_meta)
    graph_p = torch.cuda.CUDAGraph()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph_p, stream=side):
            for t in range(W, total):
                launch_packed(t, side.cuda_stream)
    torch.cuda.current_stream().wait_stream(side)
    graph_p.replay()
    torch.cuda.synchronize()
    env._bits.copy_(warm_bits)
    env._meta.copy_(warm_meta)
    barrier()
    ev0.record()
    graph_p.replay()
    ev1.record()
    barrier()
    ms_packed = ev0.elapsed_time(ev1)
    verified = verified and env.state_checksum() == want_digest

    # ---- e2e: host buffers through the public API -----------------------------------------------------
    e2e_steps = min(K, args.e2e_steps)
    host_actions = torch.empty((W + e2e_steps, envs), dtype=torch.long).pin_memory()
    host_actions.copy_(actions[: W + e2e_steps])
    host_out = torch.empty(5 * envs, dtype=torch.uint8).pin_memory()
    def e2e_pass(zero_copy):
        restore()
        for t in range(W):
            env.step_host(host_actions[t], host_out, autoreset=True, out=(obs_ring[t % ring], mask_ring[t % ring]),
                          zero_copy=zero_copy)
        barrier()
        t0 = time.perf_counter()
        ev0.record()
        for t in range(W, W + e2e_steps):
            _, r_host, d_host = env.step_host(host_actions[t], host_out, autoreset=True,
                                              out=(obs_ring[t % ring], mask_ring[t % ring]), zero_copy=zero_copy)
        ev1.record()
        barrier()
        good = env.state_checksum() == want_digest if e2e_steps == K else True
        return max(ev0.elapsed_time(ev1), 1e3 * (time.perf_counter() - t0)), good

    e2e_copy_ms, ok1 = e2e_pass(False)     # cudaMemcpyAsync H2D + launch + cudaMemcpyAsync D2H + sync
    e2e_zc_ms, ok2 = e2e_pass(True)        # kernel dereferences the pinned host buffers: launch + sync
    verified = verified and ok1 and ok2
    e2e_ms = min(e2e_copy_ms, e2e_zc_ms)
    clocks = sampler.stop()

    # ---- reduce over ranks -------------------------------------------------------------------------
    times = torch.tensor([ms, ms_eager, e2e_ms, e2e_copy_ms, e2e_zc_ms, ms_packed], dtype=torch.float64, device=dev)
    ok = torch.tensor([1.0 if verified else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)          # end-of-run statistics over NCCL
    ms, ms_eager, e2e_ms, e2e_copy_ms, e2e_zc_ms, ms_packed = (float(x) for x in times.tolist())
    verified = bool(ok.item() == 1.0)

    if rank == 0:
        peak, peak_src = measured_peak()
        total_envs = envs * world
        value = total_envs * K / (ms * 1e-3)
        launch_us = 1e3 * ms / K
        achieved = ALG_BYTES_PER_ENV_STEP * envs / (launch_us * 1e-6) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": SCALING, "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {
                "workload": workload_name(envs), "envs_per_gpu": envs, "global_envs": total_envs,
                "launch": f"one CUDA graph of {K} step_dense_kernel nodes{'' if args.no_pdl else ' with programmatic dependent launch edges'} (eager launches: "
                          f"{total_envs * K / (ms_eager * 1e-3):.4g} {UNIT})",
                "l2": f"obs/mask outputs rotate through a ring of {ring} buffer sets "
                      f"({ring * per_set / 2**20:.0f} MiB > 126 MiB L2); inputs: {K} distinct action batches",
                "parallelism": f"env-shard x{world}, no per-step collective",
            },
            "verified_state_digest": verified,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": profiled_traffic() if (M, N_COLS) == (9, 9) and envs == 65536 else None, "kernel": f"step_dense_kernel<SGeom<{M},{N_COLS},{K_LINE}>>",
                         "alg_bytes_per_env_step": ALG_BYTES_PER_ENV_STEP, "launch_us": launch_us, "peak_source": peak_src},
            "e2e": {"value": total_envs * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 8 * envs,
                    "d2h_bytes_per_step": 5 * envs, "steps": e2e_steps,
                    "api": "TorchVectorMnkEnv.step_host -> mnk_step_host: pinned int64 actions in, f32 rewards + bool dones out, "
                           "stream synchronised every step; value = the faster of the two transports",
                    "staged_copies": total_envs * e2e_steps / (e2e_copy_ms * 1e-3),
                    "zero_copy": total_envs * e2e_steps / (e2e_zc_ms * 1e-3)},
            "packed_mode": {"value": total_envs * K / (ms_packed * 1e-3), "unit": UNIT, "launch_us": 1e3 * ms_packed / K,
                            "alg_bytes_per_env_step": PACKED_BYTES_PER_ENV_STEP,
                            "hbm_frac": PACKED_BYTES_PER_ENV_STEP * envs / (1e3 * ms_packed / K * 1e-6) / 1e9 / peak,
                            "note": "same K steps, observation / mask not materialised (state + action + reward / done "
                                    "traffic only); latency-bound, reported for SURVEY 8d, not the headline"},
            "gpu_launches": K,
            "clocks": clocks,
            "stats": {"episodes": stats[0].item(), "wins": stats[1].item(), "plies": stats[2].item()},
        }
        if world == 1 and not args.no_cpu_baseline:
            acts_cpu = actions[: W + args.cpu_steps].cpu()

            def check_state(s, n_steps):
                # cross-check: the CPU port after n_steps of the SAME trace == the CUDA env after n_steps
                restore()
                for t in range(n_steps):
                    launch(t, stream)
                torch.cuda.synchronize()
                same = torch.equal(env.boards.cpu(), s.planes) and torch.equal(env.move_counts.cpu(), s.plies)
                env.release_mirrors()
                return bool(same)

            restore()
            init = (env.boards.cpu(), env.current_player.cpu(), env.move_counts.cpu())
            env.release_mirrors()
            res = cpu_port_run(acts_cpu, envs, min(W, 3), args.cpu_steps - min(W, 3), args.cpu_budget, check_state, init)
            line["cpu_baseline"] = {
                "value": res["value"], "unit": UNIT, "cores": res["threads"], "kind": "port",
                "sample": f"first {res['steps']} steps of the same action trace, all {envs} envs, oracle/torch_port.py "
                          f"(reference torch op sequence) on {os.cpu_count()} host CPUs",
                "parity_with_gpu_state": res["parity"]}
            # informative second baseline: the same torch op sequence on this B200 (device="cuda")
            gsteps = min(args.cpu_steps, 40)
            gres = cpu_port_run(actions, envs, min(W, 3), gsteps - min(W, 3), 60.0, None,
                                tuple(x.to(dev) for x in init), device=dev)
            line["cpu_baseline"]["stock_torch_same_gpu"] = {
                "value": gres["value"], "unit": UNIT, "steps": gres["steps"],
                "note": "oracle/torch_port.py with device='cuda' (stock PyTorch kernels, as the reference would run on "
                        "this GPU), same trace, host-synchronised per step like the reference's loop"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# cfg3: self-play rollout with the policy/value network (secondary workload, --workload cfg3)
# ------------------------------------------------------------------------------------------------
def run_rollout_arm(args):
    """BASELINE cfg3 per GPU: 9x9x5, agent and opponent both resnet_b_s (same random-init weights, opponent
    frozen), tcgen05 forward fed from bitboards, Gumbel-max sampling, fused wrapper, packed PPO buffer, on-device
    episode statistics; K rollout steps timed.  Metric = the reference's fps (ppo.py:126-129): agent steps / s."""
    import copy
    import torch
    import torch.distributed as dist
    from mnk_b200 import (NativeNNPolicy, NativeResNet, ResNetActorCritic, RolloutBuffer, RolloutCollector,
                          TorchSelfPlayWrapper, TorchVectorMnkEnv)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    envs, K, W = args.envs, args.steps, args.warmup
    torch.manual_seed(0)
    net = ResNetActorCritic((2, M, N_COLS), CELLS).to(dev).eval()
    agent = NativeResNet(net, device=dev)
    opponent = NativeNNPolicy(copy.deepcopy(net), device=dev, seed=7)
    env = TorchVectorMnkEnv(M, N_COLS, K_LINE, envs, device=f"cuda:{local_rank}", env_offset=rank * envs)
    wr = TorchSelfPlayWrapper(env, seed=20261018)
    wr.set_opponent(opponent)
    col = RolloutCollector(envs, device=dev, seed=11, row_offset=rank * envs, world_size=world)
    wr.reset(materialise=False)
    col._last_obs = {"observation": None, "action_mask": None}
    warm_buf = RolloutBuffer(max(W, MIX_PLIES // 2), envs, (2, M, N_COLS), CELLS, device=dev, k=K_LINE)
    col.collect(agent, wr, warm_buf)                       # warm-up: also brings games to a stationary depth mix
    use_graph = not args.no_graph
    buf = RolloutBuffer(K, envs, (2, M, N_COLS), CELLS, device=dev, k=K_LINE)
    if use_graph:                                          # capture the K-step rollout once; its first replay is untimed
        col.collect(agent, wr, buf, graph=True)
        buf.reset()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    ev0.record()
    stats = col.collect(agent, wr, buf, graph=use_graph)   # K steps; ends with the NCCL all-reduce + one host read
    ev1.record()
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    agent.check_error()
    opponent.net.check_error()
    times = torch.tensor([ms, wall_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, wall_ms = (float(x) for x in times.tolist())
    if rank == 0:
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peak, peak_src = float(json.load(f)["bf16_tflops_sustained"]), "MEASURED_PEAKS.json bf16_tflops_sustained"
        except Exception:
            peak, peak_src = 1400.0, "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
        flop_per_agent_step = 2 * 12.136e6                 # agent + opponent forward
        total = envs * world * K
        value = total / (ms * 1e-3)
        achieved = value * flop_per_agent_step / 1e12
        line = {
            "metric": "self-play rollout steps/sec (9x9x5, resnet_b_s agent + opponent)", "value": value, "unit": "agent-steps/s",
            "n_gpus": world, "steps": K, "warmup": warm_buf.n_steps, "ms_per_step": ms / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"cfg3: gomoku 9x9x5 self-play rollout, {envs} envs/GPU x {K} steps, resnet_b_s agent + frozen "
                                   "copy as opponent (random-init weights), tcgen05 forward from bitboards, Gumbel-max sampling, "
                                   "fused wrapper, packed PPO buffer, on-device episode stats"
                                   + (", the whole K-step rollout replayed as one CUDA graph" if use_graph else ", eager launches"),
                       "envs_per_gpu": envs, "global_envs": envs * world,
                       "l2": f"per step the towers stream {envs * 72 / 2**20:.1f} MiB of bitboards and {envs * 972 * 2 / 2**20:.0f} MiB of "
                             "head features; rollout buffer slots are distinct per step",
                       "parallelism": f"env-shard x{world}, one NCCL all-reduce of 6 doubles per rollout"},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": None, "kernel": "resnet_tower_kernel (2 launches per agent-step)",
                         "flop_per_agent_step": flop_per_agent_step, "peak_source": peak_src},
            "e2e": {"value": total / (wall_ms * 1e-3), "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 48.0 / K,
                    "api": "RolloutCollector.collect (wall clock incl. the statistics all-reduce and host read)"},
            "gpu_launches": 10 * K,
            "clocks": clocks,
            "stats": {"episodes": stats.episodes, "mean_reward": stats.mean_reward, "mean_length": stats.mean_length,
                      "wins": stats.wins, "losses": stats.losses, "draws": stats.draws},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS) + ["cfg3"],
                    help="cfg2 (default, the headline), cfg4 / cfg5 (larger boards), cfg3 (self-play rollout with the network)")
    ap.add_argument("--envs", type=int, default=None, help="envs per GPU (default: the workload's)")
    ap.add_argument("--e2e-steps", type=int, default=500)
    ap.add_argument("--cpu-steps", type=int, default=60)
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="cfg3: eager launches instead of one CUDA graph per rollout")
    ap.add_argument("--no-pdl", action="store_true", help="plain launches instead of programmatic dependent launch")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.workload == "cfg3":
        configure("cfg2", args.gpus, args.envs or 32768)
        args.envs = ENVS_PER_GPU
        return run_rollout_arm(args)
    configure(args.workload, args.gpus, args.envs)
    args.envs = ENVS_PER_GPU
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()

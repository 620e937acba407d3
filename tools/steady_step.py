"""cfg2 steady state for an ncu window: 60 consecutive eager launches of step_dense_kernel (9x9x5, 65,536 envs, API-exact
outputs rotating through a ring of 9 buffer sets = 410 MiB > L2), nothing else in between.  Run under

  ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
      -k regex:step_dense -s 24 -c 18 --csv --log-file gpurun_out/steady.csv python tools/steady_step.py

so that the profiled launches see the cache state the previous launches left (no flush, no replay passes: the three
metrics fit one pass), i.e. the DRAM traffic of the timed region of bench.py rather than that of a cold, isolated launch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rl-selfplay-mnk_b200")); sys.path.insert(0, ROOT)
import torch
from mnk_b200 import TorchVectorMnkEnv, _lib

m, n, k, envs, steps, ring = 9, 9, 5, 65536, 60, 9
dev = torch.device("cuda", 0)
env = TorchVectorMnkEnv(m, n, k, envs, device="cuda:0")
env.reset()
for t in range(64):
    env.step_autoreset(env.random_legal_actions(1, t), materialise=False)
snap_bits, snap_meta = env._bits.clone(), env._meta.clone()
actions = torch.empty((steps, envs), dtype=torch.long, device=dev)
for t in range(steps):
    env.random_legal_actions(1, 64 + t, out=actions[t])
    env.step_autoreset(actions[t], materialise=False)
env._bits.copy_(snap_bits), env._meta.copy_(snap_meta)
obs = [torch.empty((envs, 2, m, n), dtype=torch.float32, device=dev) for _ in range(ring)]
mask = [torch.empty((envs, m * n), dtype=torch.bool, device=dev) for _ in range(ring)]
rewards = torch.empty(envs, dtype=torch.float32, device=dev)
dones = torch.empty(envs, dtype=torch.bool, device=dev)
L = _lib.lib()
stream = torch.cuda.current_stream().cuda_stream
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for t in range(steps):
    _lib.check(L.mnk_step(env._stp, actions[t].data_ptr(), None, envs, rewards.data_ptr(), dones.data_ptr(),
                          obs[t % ring].data_ptr(), mask[t % ring].data_ptr(), None, _lib.STEP_AUTORESET, stream), "mnk_step")
e1.record()
torch.cuda.synchronize()
print(f"{steps} eager launches: {1e3 * e0.elapsed_time(e1) / steps:.2f} us per launch, "
      f"{814 * envs / (e0.elapsed_time(e1) / steps * 1e-3) / 1e9:.0f} GB/s algorithmic")

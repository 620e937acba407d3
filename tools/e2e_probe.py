"""Where the end-to-end host step spends its time: env.step_host vs the bare C-ABI call with cached pointers."""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rl-selfplay-mnk_b200")); sys.path.insert(0, ROOT)
import torch
from mnk_b200 import TorchVectorMnkEnv, _lib

N = 65536
env = TorchVectorMnkEnv(9, 9, 5, N, device="cuda")
env.reset()
T = 300
acts = torch.stack([env.random_legal_actions(1, t) for t in range(8)]).cpu().pin_memory()   # legal only at t=0; fine for timing
out = torch.empty(5 * N, dtype=torch.uint8).pin_memory()
obs = torch.empty((N, 2, 9, 9), device="cuda"); mask = torch.empty((N, 81), dtype=torch.bool, device="cuda")
for zc in (True, False):
    for _ in range(20):
        env.step_host(acts[0], out, autoreset=True, out=(obs, mask), zero_copy=zc)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for t in range(T):
        env.step_host(acts[t & 7], out, autoreset=True, out=(obs, mask), zero_copy=zc)
    dt = (time.perf_counter() - t0) / T
    print(f"env.step_host zero_copy={zc}: {dt*1e6:.1f} us/step")
L = _lib.lib()
stream = torch.cuda.current_stream().cuda_stream
flags = _lib.STEP_AUTORESET | _lib.STEP_ZEROCOPY
ap = [acts[i].data_ptr() for i in range(8)]
op, mp, hp = obs.data_ptr(), mask.data_ptr(), out.data_ptr()
stp = env._stp
torch.cuda.synchronize(); t0 = time.perf_counter()
for t in range(T):
    L.mnk_step_host(stp, ap[t & 7], None, None, hp, op, mp, flags, stream)
dt = (time.perf_counter() - t0) / T
print(f"bare ctypes mnk_step_host zero-copy: {dt*1e6:.1f} us/step")
# kernel alone with device-resident actions, sync each step
da = acts.cuda(); r = torch.empty(N, device="cuda"); d = torch.empty(N, dtype=torch.bool, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for t in range(T):
    L.mnk_step(stp, da[t & 7].data_ptr(), None, N, r.data_ptr(), d.data_ptr(), op, mp, None, _lib.STEP_AUTORESET, stream)
    torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / T
print(f"device-resident step + synchronize each step: {dt*1e6:.1f} us/step")

"""Host<->device copy bandwidth from pinned memory as bench.py's e2e leg sees it: 4 MiB H2D / 2.5 MiB D2H chunks (one 8-step slab
of cfg2), alone and concurrently, with and without the NUMA pinning of mnk_b200.dist.pin_to_gpu_numa."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rl-selfplay-mnk_b200"))
import torch
from mnk_b200.dist import pin_to_gpu_numa
if len(sys.argv) > 1 and sys.argv[1] == "pin":
    print("pinned to:", pin_to_gpu_numa(0))
torch.cuda.init()
h_in = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
h_out = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
d_in = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
d_out = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
def run(nin, nout, reps=200):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(reps):
        if nin:
            with torch.cuda.stream(s_in):
                d_in[:nin].copy_(h_in[:nin], non_blocking=True)
        if nout:
            with torch.cuda.stream(s_out):
                h_out[:nout].copy_(d_out[:nout], non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return dt
for nin, nout, label in ((4 << 20, 0, "H2D 4 MiB"), (0, 2621440, "D2H 2.5 MiB"), (4 << 20, 2621440, "both"), (32 << 20, 0, "H2D 32 MiB"), (0, 32 << 20, "D2H 32 MiB")):
    dt = run(nin, nout)
    print(f"{label:12s}: {dt * 1e6:7.1f} us per round  H2D {nin / dt / 1e9:5.1f} GB/s  D2H {nout / dt / 1e9:5.1f} GB/s")

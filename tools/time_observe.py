"""Times env.observe() / legal_mask() / step_subset (CUDA events) against the HBM roofline of their bytes."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rl-selfplay-mnk_b200")); sys.path.insert(0, ROOT)
import torch
from mnk_b200 import TorchVectorMnkEnv, _lib
if os.environ.get("MNK_LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["MNK_LIB"])

peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6455.9
for (m, n, k, ne) in ((9, 9, 5, 65536), (9, 9, 5, 1 << 20), (13, 13, 5, 1 << 20), (19, 19, 5, 1 << 19)):
    env = TorchVectorMnkEnv(m, n, k, ne, device="cuda")
    env.reset()
    for t in range(20):
        env.step_autoreset(env.random_legal_actions(1, t), materialise=False)
    cells, words = m * n, (m * (n + 1) + 63) // 64
    state = 2 * 8 * words + 4
    ring = [env._new_obs() for _ in range(max(2, (3 * 126 * 2**20) // (ne * 9 * cells) + 1))]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    def timed(fn, reps=50):
        for i in range(5): fn(i)
        torch.cuda.synchronize(); e0.record()
        for i in range(reps): fn(i)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3
    us = timed(lambda i: env._call(env._L.mnk_observe, ring[i % len(ring)][0].data_ptr(), ring[i % len(ring)][1].data_ptr(), None, 0))
    b = ne * (state + 8 * cells + cells)
    us_m = timed(lambda i: env.legal_mask())
    bm = ne * (state + cells)
    print(f"{m}x{n} envs={ne}: observe {us:.1f} us = {b/us/1e3:.0f} GB/s ({b/us/1e3/peak:.2f} of HBM peak); "
          f"legal_mask {us_m:.1f} us = {bm/us_m/1e3:.0f} GB/s ({bm/us_m/1e3/peak:.2f})")

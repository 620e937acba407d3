"""Soak test of the board-row tower kernel: many launches at random batch sizes / boards, each compared with the tap kernel."""
import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rl-selfplay-mnk_b200")); sys.path.insert(0, ROOT)
import torch
from mnk_b200 import NativeResNet, ResNetActorCritic, TorchVectorMnkEnv

random.seed(0); torch.manual_seed(0)
worst = 0.0
for (m, n, k) in ((9, 9, 5), (7, 7, 4), (3, 3, 3), (10, 10, 5)):
    native = NativeResNet(ResNetActorCritic((2, m, n), m * n).cuda().eval())
    for it in range(120):
        ne = random.choice([1, 2, 11, 12, 13, 100, 1000, random.randint(1, 6000), 32768 if it % 40 == 0 else 77])
        env = TorchVectorMnkEnv(m, n, k, ne, device="cuda")
        env.reset()
        for t in range(random.randint(0, m * n - 1)):
            env.step_autoreset(env.random_legal_actions(it, t), materialise=False)
        swap = (torch.rand(ne, device="cuda") < 0.5).to(torch.uint8) if it % 2 else None
        native.use_rows_kernel = True
        outs = [native.features(env._st, ne, m * n, swap) for _ in range(3)]     # back-to-back launches must agree bit for bit
        native.use_rows_kernel = False
        pt, vt = native.features(env._st, ne, m * n, swap)
        for pf, vf in outs[1:]:
            assert torch.equal(pf, outs[0][0]) and torch.equal(vf, outs[0][1]), (m, n, ne, it, "non-deterministic")
        scale = float(pt.abs().max()) + 1e-6
        d = float((outs[0][0] - pt).abs().max()) / scale
        worst = max(worst, d)
        assert d < 3e-2, (m, n, ne, it, d)
    native.check_error()
    print(f"{m}x{n}: 120 iterations ok, worst rel diff vs tap kernel so far {worst:.2e}", flush=True)
print("soak ok")

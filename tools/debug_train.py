"""Bounded probe of mnk_resnet_tower_train at a given env count: time per forward and the barrier-timeout flags."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rl-selfplay-mnk_b200")); sys.path.insert(0, ROOT)
import torch
from mnk_b200 import NativeResNet, ResNetActorCritic, TorchVectorMnkEnv, _lib
if os.environ.get("MNK_LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["MNK_LIB"])
ne = int(sys.argv[1])
m, n, k = 9, 9, 5
torch.manual_seed(0)
net = ResNetActorCritic((2, m, n), m * n).cuda()
native = NativeResNet(net, bn_mode="train")
native._err = torch.zeros(32, dtype=torch.int32, device="cuda")
env = TorchVectorMnkEnv(m, n, k, ne, device="cuda")
env.reset()
for t in range(20):
    env.step_autoreset(env.random_legal_actions(1, t), materialise=False)
bad = 0
for rep in range(int(os.environ.get("REPS", 3))):
    native._err.zero_()
    torch.cuda.synchronize()
    t0 = time.time()
    pf, vf = native.features(env._st, ne, m * n, None)
    torch.cuda.synchronize()
    ms = 1e3 * (time.time() - t0)
    er = native._err.tolist()
    bad += 1 if er[0] else 0
    if er[0] or rep < 2 or rep == int(os.environ.get("REPS", 4)) - 1:
        print(f"envs={ne} forward {rep}: {ms:.2f} ms code {hex(er[0])} layers wts/watch/tma {hex(er[1])}/{hex(er[2])}/{hex(er[3])} "
              f"watcher e min {100000 - er[4] if er[4] else None} max {er[5]} ctas {er[6]}", flush=True)
        if er[0]:
            print(f"    cta {er[7] & 0xFFFF} total_steps {er[7] >> 16}: per-warp (phase, index): " + " ".join(f"{w}:{x >> 16}/{x & 0xFFFF}" for w, x in enumerate(er[8:26])), flush=True)
print(f"envs={ne}: {bad} of {int(os.environ.get('REPS', 4))} forwards hit a barrier timeout", flush=True)

"""Soak / post-mortem probe of mnk_resnet_tower_train at a given env count: time per forward and the barrier-timeout flag
(0x100 weights, 0x200 commit watcher, 0x400 operand TMA; low bits = layer); on a timeout the kernel's post-mortem words
(CTA, step, layer, raw commit barriers) are read back from the scratch buffer."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rl-selfplay-mnk_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
from mnk_b200 import NativeResNet, ResNetActorCritic, TorchVectorMnkEnv, _lib
if os.environ.get("MNK_LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["MNK_LIB"])
ne = int(sys.argv[1])
m, n, k = 9, 9, 5
reps = int(os.environ.get("REPS", 3))
torch.manual_seed(0)
net = ResNetActorCritic((2, m, n), m * n).cuda()
native = NativeResNet(net, bn_mode="train")
env = TorchVectorMnkEnv(m, n, k, ne, device="cuda")
env.reset()
for t in range(20):
    env.step_autoreset(env.random_legal_actions(1, t), materialise=False)
bad = 0
for rep in range(reps):
    torch.cuda.synchronize()
    t0 = time.time()
    pf, vf = native.features(env._st, ne, m * n, None)
    torch.cuda.synchronize()
    ms = 1e3 * (time.time() - t0)
    code = int(native._err.item())
    bad += 1 if code else 0
    if code or rep < 2 or rep == reps - 1:
        print(f"envs={ne} forward {rep}: {ms:.2f} ms code {hex(code)}", flush=True)
    if code:
        import _postmortem
        print(_postmortem.describe(_postmortem.read(native, _lib.lib(), m, n, ne), m), flush=True)
        native._scratch.zero_()
        native._err.zero_()
    if bad >= int(os.environ.get("MAX_BAD", 3)):
        break
print(f"envs={ne}: {bad} of {rep + 1} forwards hit a barrier timeout", flush=True)

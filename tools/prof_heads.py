"""Runs the heads kernel (and one tower launch) a few times at 32,768 envs -- the command profiled by ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rl-selfplay-mnk_b200")); sys.path.insert(0, ROOT)
import torch
from mnk_b200 import NativeResNet, ResNetActorCritic, TorchVectorMnkEnv

m = n = 9
ne = 32768
torch.manual_seed(0)
native = NativeResNet(ResNetActorCritic((2, m, n), m * n).cuda().eval())
env = TorchVectorMnkEnv(m, n, 5, ne, device="cuda")
env.reset()
for t in range(20):
    env.step_autoreset(env.random_legal_actions(1, t), materialise=False)
pf, vf = native.features(env._st, ne, m * n, None)
for _ in range(4):
    native.tails(pf, vf)
torch.cuda.synchronize()
native.check_error()
print("ok")

"""Cycle timeline of one mid-grid CTA of the board-row tower kernel (needs a -DMNK_TIMELINE variant build:
tools/ab_tower.sh; MNK_LIB=.../lib_<name>.so python tools/timeline_rows.py)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rl-selfplay-mnk_b200")); sys.path.insert(0, ROOT)
import torch
from mnk_b200 import NativeResNet, ResNetActorCritic, TorchVectorMnkEnv, _lib
_lib.LIB_PATH = os.path.abspath(os.environ["MNK_LIB"])

m = n = 9
ne = 32768
torch.manual_seed(0)
native = NativeResNet(ResNetActorCritic((2, m, n), m * n).cuda().eval())
env = TorchVectorMnkEnv(m, n, 5, ne, device="cuda")
env.reset()
for t in range(20):
    env.step_autoreset(env.random_legal_actions(1, t), materialise=False)
native._err = torch.zeros(1 + 8 * 48, dtype=torch.int32, device="cuda")
for _ in range(3):
    native.features(env._st, ne, m * n, None)
torch.cuda.synchronize()
t = native._err.cpu().tolist()
print("idx mma_waited  mma_go  mma_done | epi_begin epi_ready  epi_loaded  epi_done | issue  wait  ready->loaded  loaded->done")
for i in range(40):
    a, b, c, d, e, w, s6, s7 = t[1 + 8 * i: 1 + 8 * i + 8]
    print(f"{i:3d} {w:10d} {a:7d} {b:9d} | {s6:9d} {c:9d} {d:11d} {e:9d} | {b-a:5d} {c-s6:5d} {d-c:13d} {e-d:13d}   math {s7-d} fence+arrive {e-s7}")

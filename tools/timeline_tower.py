"""Per-layer cycle timeline of one mid-grid CTA of the tower kernel (needs a -DMNK_TIMELINE variant build:
tools/ab_tower.sh timeline "-DMNK_TIMELINE"; MNK_LIB=.../lib_timeline.so python tools/timeline_tower.py)."""
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
native._err = torch.zeros(1 + 8 * 16, dtype=torch.int32, device="cuda")
native.use_rows_kernel = False          # this tool times the tap kernel (mnk_resnet.cu)
for _ in range(3):
    native.features(env._st, ne, m * n, None)
torch.cuda.synchronize()
t = native._err.cpu().tolist()
print("layer  issue_start  issue_end  epi_wake  epi_done  barrier   | issue  drain+wake  epilogue  sync  layer_total")
prev = None
for L in range(9):
    a, b, c, d, e = t[1 + 8 * L: 1 + 8 * L + 5]
    print(f"{L:5d} {a:11d} {b:10d} {c:9d} {d:9d} {e:8d}   | {b-a:5d} {c-b:10d} {d-c:9d} {e-d:5d} {e-(prev if prev is not None else a):11d}")
    prev = e

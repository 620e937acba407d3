"""Phase timeline of mnk_transformer_body (a -DTF_TIMELINE build: MNK_LIB=... python tools/timeline_tf.py [arch]): cycles between
consecutive phase boundaries of one mid-grid CTA, first layer."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rl-selfplay-mnk_b200")); sys.path.insert(0, ROOT)
import torch
from mnk_b200 import TorchVectorMnkEnv, build_architecture, native_network, _lib
_lib.LIB_PATH = os.path.abspath(os.environ["MNK_LIB"])
arch = sys.argv[1] if len(sys.argv) > 1 else "transformer_b_s"
m, n, k, ne = 9, 9, 5, 4096
net = build_architecture(arch, (2, m, n), m * n).cuda().eval()
fwd = native_network(net)
fwd._err = torch.zeros(256, dtype=torch.int32, device="cuda")
env = TorchVectorMnkEnv(m, n, k, ne, device="cuda")
env.reset()
for _ in range(3):
    fwd.features(env._st, ne, m * n, None)
torch.cuda.synchronize()
st = fwd._err.cpu().tolist()[1:]
heads = fwd.heads
names = ["LN1 -> bufA", "in_proj MMA", "in_proj epilogue (Q, K, V^T)", "h0 scores MMA"]
for h in range(heads):
    names += [f"h{h} softmax -> P", f"h{h} PV MMA (+ scores of h{h + 1})"]
names += ["O epilogue + out_proj bias", "out_proj MMA", "LN2 -> bufA", "linear1 MMA", "linear1 epilogue + bias", "linear2 MMA"]
for layer in range(2):
    base = layer * len(names)
    prev = st[base - 1] if layer else 0
    print(f"layer {layer}:")
    for i, nm in enumerate(names):
        t = st[base + i]
        if nm is not None:
            print(f"  {nm:32s} +{t - prev:6d}   (at {t})")
        prev = t

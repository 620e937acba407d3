"""Times mnk_conv_tower / mnk_transformer_body (the wider convolutional and the transformer bodies) alone with CUDA events:
useful TFLOP/s (3x3 convolutions; for transformers the Linear / attention matmuls of the reference module) against the measured
bf16 peak, and the same eval-mode forward through the stock torch module (fp32 / TF32 and bf16 autocast).
   python tools/time_convnet.py [arch] [m n k]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rl-selfplay-mnk_b200")); sys.path.insert(0, ROOT)
import torch
from mnk_b200 import TorchVectorMnkEnv, build_architecture, native_network, _lib
if os.environ.get("MNK_LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["MNK_LIB"])
arch = sys.argv[1] if len(sys.argv) > 1 else "resnet_b_l"
m, n, k = (int(x) for x in (sys.argv[2:5] if len(sys.argv) > 4 else (9, 9, 5)))
torch.manual_seed(0)
net = build_architecture(arch, (2, m, n), m * n).cuda().eval()
native = native_network(net)
convs = [mod for mod in net.modules() if isinstance(mod, torch.nn.Conv2d) and mod.kernel_size == (3, 3)]
flops = sum(2 * m * n * c.in_channels * c.out_channels * 9 for c in convs)
if hasattr(net, "transformer"):      # per layer: in_proj + out_proj + feed-forward (12 D^2 MACs per token) + QK^T and PV (2 T D per token)
    D, T = net.embed_dim, m * n
    flops = net.num_layers * (2 * T * 12 * D * D + 2 * 2 * T * T * D)
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops_sustained": 1391.8}
for ne in (4096, 32768)[: int(os.environ.get("MNK_SIZES", 2))]:
    env = TorchVectorMnkEnv(m, n, k, ne, device="cuda")
    env.reset()
    for t in range(20):
        env.step_autoreset(env.random_legal_actions(1, t), materialise=False)
    for _ in range(3):
        native.features(env._st, ne, m * n, None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        native.features(env._st, ne, m * n, None)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    e0.record()
    for _ in range(reps):
        native.forward_env(env)
    e1.record(); torch.cuda.synchronize()
    ms_full = e0.elapsed_time(e1) / reps
    tf = ne * flops / (ms * 1e-3) / 1e12
    print(f"{arch} {m}x{n} envs={ne}: tower {ms:.3f} ms = {ne / ms * 1e3 / 1e6:.2f} M samples/s = {tf:.1f} useful TFLOP/s "
          f"({tf / peaks['bf16_tflops_sustained']:.3f} of sustained bf16 peak; {flops / 1e6:.1f} MFLOP/sample); forward_env {ms_full:.3f} ms")
    if ne <= 4096:
        obs = env.observe()["observation"]
        with torch.no_grad():
            for label, ctx in (("fp32 (TF32 convs allowed)", torch.autocast("cuda", enabled=False)), ("bf16 autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
                with ctx:
                    for _ in range(2):
                        net(obs, None)
                    torch.cuda.synchronize()
                    e0.record()
                    for _ in range(5):
                        net(obs, None)
                    e1.record(); torch.cuda.synchronize()
                print(f"    stock PyTorch eval forward, {label}: {e0.elapsed_time(e1) / 5:.3f} ms")
native.check_error()

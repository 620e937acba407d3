"""Times mnk_resnet_tower alone (CUDA events) and reports useful TFLOP/s against the measured bf16 peak."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rl-selfplay-mnk_b200")); sys.path.insert(0, ROOT)
import torch
from mnk_b200 import NativeResNet, ResNetActorCritic, TorchVectorMnkEnv, _lib
if os.environ.get("MNK_LIB"):          # A/B experiments: a variant build of the library (tools/ab_tower.sh)
    _lib.LIB_PATH = os.path.abspath(os.environ["MNK_LIB"])

m, n, k = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (9, 9, 5)))
flops_per_sample = {(9, 9): 12.136e6, (13, 13): 25.3e6, (19, 19): 54.1e6}[(m, n)]
torch.manual_seed(0)
net = ResNetActorCritic((2, m, n), m * n).cuda().eval()
native = NativeResNet(net, bn_mode=os.environ.get("MNK_BN", "eval"))
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops_sustained": 1391.8}
for ne in (4096, 32768, 262144)[: int(os.environ.get('MNK_SIZES', 3))]:
    env = TorchVectorMnkEnv(m, n, k, ne, device="cuda")
    env.reset()
    for t in range(20):
        env.step_autoreset(env.random_legal_actions(1, t), materialise=False)
    for _ in range(3):
        native.features(env._st, ne, m * n, None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        native.features(env._st, ne, m * n, None)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    e0.record()
    for _ in range(reps):
        lg, v = native.forward_env(env)
    e1.record(); torch.cuda.synchronize()
    ms_full = e0.elapsed_time(e1) / reps
    pf, vf = native.features(env._st, ne, m * n, None)
    if ne == 4096 and os.environ.get("MNK_SAVE"):
        torch.save((pf.cpu(), vf.cpu()), os.environ["MNK_SAVE"])
    if ne == 4096 and os.environ.get("MNK_CHECK"):
        want_pf, want_vf = torch.load(os.environ["MNK_CHECK"])
        print("variant vs saved features: max |d| =", float((pf.cpu() - want_pf).abs().max()), float((vf.cpu() - want_vf).abs().max()))
    for _ in range(3):
        native.tails(pf, vf)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        native.tails(pf, vf)
    e1.record(); torch.cuda.synchronize()
    ms_heads = e0.elapsed_time(e1) / reps
    native.torch_heads = True
    e0.record()
    for _ in range(reps):
        lg, v = native.forward_env(env)
    e1.record(); torch.cuda.synchronize()
    ms_torch = e0.elapsed_time(e1) / reps
    native.torch_heads = False
    tf = ne * flops_per_sample * 0.98 / (ms * 1e-3) / 1e12
    print(f"{m}x{n} envs={ne}: tower {ms:.3f} ms = {ne/ms*1e3/1e6:.2f} M samples/s = {tf:.1f} useful TFLOP/s "
          f"({tf/peaks['bf16_tflops_sustained']:.3f} of sustained bf16 peak); heads kernel {ms_heads:.3f} ms; forward_env {ms_full:.3f} ms (with torch heads {ms_torch:.3f} ms)")
    if os.environ.get("MNK_BN") == "train" and ne <= 32768:      # the same forward in stock PyTorch (cuDNN, train-mode BatchNorm)
        obs = env.observe()["observation"]
        net.train()
        with torch.no_grad():
            for _ in range(2):
                net(obs, None)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                net(obs, None)
            e1.record(); torch.cuda.synchronize()
        print(f"    stock PyTorch train-mode forward (TF32 convs allowed): {e0.elapsed_time(e1) / 5:.3f} ms")
native.check_error()

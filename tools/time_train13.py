import os, sys, time
sys.path.insert(0, "rl-selfplay-mnk_b200"); sys.path.insert(0, ".")
import torch
from mnk_b200 import NativeResNet, ResNetActorCritic, TorchVectorMnkEnv
m, n, k = 13, 13, 5
for ne in (4096, 32768):
    torch.manual_seed(0)
    net = ResNetActorCritic((2, m, n), m * n).cuda().train()
    native = NativeResNet(net, bn_mode="train")
    env = TorchVectorMnkEnv(m, n, k, ne, device="cuda")
    env.reset()
    for t in range(40):
        env.step_autoreset(env.random_legal_actions(1, t), materialise=False)
    for _ in range(3):
        native.forward_env(env)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        native.forward_env(env)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    native.check_error()
    line = f"13x13 train-mode forward_env, {ne} envs: {ms:.3f} ms"
    if ne == 4096:
        obs = env.observe()["observation"]
        with torch.no_grad():
            for _ in range(2): net(obs, None)
            e0.record()
            for _ in range(3): net(obs, None)
            e1.record(); torch.cuda.synchronize()
        line += f"; stock PyTorch train-mode forward {e0.elapsed_time(e1) / 3:.3f} ms"
    print(line)

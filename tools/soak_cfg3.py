"""Soak of the cfg3 rollout (train-mode agent + eval-mode opponent, heads, sampler, fused wrapper) as bench.py runs it:
R rollouts of K steps, eager or as one CUDA graph, reading the barrier-timeout flags of both networks after each rollout
(0x1 tower, 0x2 heads, 0x1LL / 0x2LL / 0x4LL train-mode layer LL: weights / commit watcher / operand TMA) and, for the
train-mode tower, the kernel's post-mortem words.   python tools/soak_cfg3.py [envs] [K] [R] [graph|eager]"""
import copy, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rl-selfplay-mnk_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
from mnk_b200 import (NativeNNPolicy, NativeResNet, ResNetActorCritic, RolloutBuffer, RolloutCollector, TorchSelfPlayWrapper,
                      TorchVectorMnkEnv, _lib)
if os.environ.get("MNK_LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["MNK_LIB"])
envs = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
K = int(sys.argv[2]) if len(sys.argv) > 2 else 100
R = int(sys.argv[3]) if len(sys.argv) > 3 else 20
graph = (sys.argv[4] if len(sys.argv) > 4 else "graph") == "graph"
agent_bn = os.environ.get("AGENT_BN", "train")
m, n, k = 9, 9, 5
if os.environ.get("DIRTY"):            # fill the caching allocator with non-zero garbage first (what a previous workload leaves)
    junk = [torch.full((1 << 28,), 0x7F7F7F7F, dtype=torch.int32, device="cuda") for _ in range(8)]
    del junk
torch.manual_seed(0)
net = ResNetActorCritic((2, m, n), m * n).cuda()
net.train(agent_bn == "train")
agent = NativeResNet(net, bn_mode=agent_bn)
opponent = NativeNNPolicy(copy.deepcopy(net), seed=7)
env = TorchVectorMnkEnv(m, n, k, envs, device="cuda")
wr = TorchSelfPlayWrapper(env, seed=1234)
wr.set_opponent(opponent)
col = RolloutCollector(envs, seed=11)
wr.reset(materialise=False)
col._last_obs = {"observation": None, "action_mask": None}
buf = RolloutBuffer(K, envs, (2, m, n), m * n, k=k)
bad = 0
for r in range(R):
    buf.reset()
    t0 = time.time()
    try:
        col.collect(agent, wr, buf, graph=graph)
        msg = "ok"
    except RuntimeError as exc:
        msg = str(exc)[-200:]
        bad += 1
        a, o = int(agent._err.item()), int(opponent.net._err.item())
        print(f"rollout {r}: agent flag {a:#x} opponent flag {o:#x}", flush=True)
        if agent_bn == "train" and a:
            import _postmortem
            print(_postmortem.describe(_postmortem.read(agent, _lib.lib(), m, n, envs), m), flush=True)
            agent._scratch.zero_()
        agent._err.zero_(); opponent.net._err.zero_()
    if r < 2 or msg != "ok" or r == R - 1:
        print(f"rollout {r}: {1e3 * (time.time() - t0):.1f} ms {msg}", flush=True)
    if bad >= 3:
        break
print(f"cfg3 soak envs={envs} K={K} {'graph' if graph else 'eager'} agent_bn={agent_bn}: {bad} of {r + 1} rollouts flagged", flush=True)

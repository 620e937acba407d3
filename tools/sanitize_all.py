"""Exercises every kernel of libmnk_b200.so once at small sizes (for compute-sanitizer runs):
   compute-sanitizer --tool memcheck python tools/sanitize_all.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rl-selfplay-mnk_b200")); sys.path.insert(0, ROOT)
import torch
from mnk_b200 import (NativeNNPolicy, NativeResNet, RandomPolicy, ResNetActorCritic, RolloutBuffer, RolloutCollector,
                      TorchSelfPlayWrapper, TorchVectorMnkEnv, masked_sample)

torch.manual_seed(0)
for (m, n, k, ne) in [(3, 3, 3, 70), (9, 9, 5, 100), (13, 13, 5, 45), (19, 19, 5, 33), (7, 11, 4, 50), (16, 31, 5, 33)]:
    env = TorchVectorMnkEnv(m, n, k, ne, device="cuda")
    obs = env.reset()
    env.boards[0, 0, 0, 0] = 1                       # pack / unpack / meta import-export mirrors
    _ = env.current_player, env.move_counts
    for t in range(6):
        a = env.random_legal_actions(1, t)
        obs, r, d = env.step(a)
        idx = torch.arange(0, ne, 3, device="cuda")
        obs, r, d = env.step_subset(env.random_legal_actions(2, t)[idx], idx)
        env.reset(torch.nonzero(d).squeeze(1))
        env.step_autoreset(env.random_legal_actions(3, t), materialise=(t % 2 == 0))
    env.release_mirrors()
    host_a = env.random_legal_actions(4, 0).cpu().pin_memory()
    host_out = torch.empty(5 * ne, dtype=torch.uint8).pin_memory()
    env.step_host(host_a, host_out, autoreset=True)
    env.step_host(env.random_legal_actions(4, 1).cpu().pin_memory(), host_out, autoreset=True, zero_copy=True)
    wr = TorchSelfPlayWrapper(env, seed=1)
    wr.set_opponent(RandomPolicy(m * n))
    obs, _ = wr.reset()
    for t in range(5):
        obs, r, term, _, _ = wr.step(RandomPolicy(m * n, seed=t).act(obs))

    class First:
        def act(self, o):
            return torch.argmax(o["action_mask"].float(), dim=1)
    wr.set_opponent(First())
    for t in range(5):
        obs, r, term, _, _ = wr.step(RandomPolicy(m * n, seed=t).act(obs))
    lg = torch.randn(ne, m * n, device="cuda")
    masked_sample(lg, obs["action_mask"], seed=1, counter=1, want_entropy=True)
    masked_sample(lg, None, given=torch.zeros(ne, dtype=torch.long, device="cuda"))
    buf = RolloutBuffer(4, ne, (2, m, n), m * n, device="cuda", k=k)
    col = RolloutCollector(ne, device="cuda")
    if n <= 22:
        net = ResNetActorCritic((2, m, n), m * n).cuda().eval()
        agent = NativeResNet(net)
        wr.set_opponent(NativeNNPolicy(net))
        wr.reset(materialise=False)
        col._last_obs = {"observation": None, "action_mask": None}
        col.collect(agent, wr, buf)
        agent.check_error()
        agent.forward(wr.get_agent_obs()["observation"], None)
    else:
        wr.set_opponent(RandomPolicy(m * n))
        col.collect(lambda o, mk: (type("D", (), {"logits": torch.log_softmax(torch.where(mk, torch.zeros_like(mk, dtype=torch.float32), torch.full_like(mk, -float("inf"), dtype=torch.float32)), 1)})(), torch.zeros(ne, 1, device="cuda")), wr, buf)
    buf.compute_advantages_and_returns(torch.zeros(ne, device="cuda"))
    for batch in buf.get_data_loader(37):
        pass
torch.cuda.synchronize()
print("sanitize_all ok")

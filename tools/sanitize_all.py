"""Exercises every kernel of libmnk_b200.so at small, ragged sizes (tail tiles, partial CTAs, six board geometries).
`exercise_all()` is what tests/test_guards_gpu.py runs with every CUDA allocation of the package wrapped in canary zones
(the pool refuses compute-sanitizer -- profiles/README.md -- so out-of-bounds WRITES are caught by the canaries instead);
as a script it is the command for a compute-sanitizer run where one is available."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rl-selfplay-mnk_b200")); sys.path.insert(0, ROOT)
import torch
from mnk_b200 import (NativeConvNet, NativeNNPolicy, NativeResNet, build_architecture, native_network, RandomPolicy, ResNetActorCritic, RolloutBuffer, RolloutCollector,
                      TorchSelfPlayWrapper, TorchVectorMnkEnv, masked_sample)


def exercise_all():
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

    # train-mode tower, tcgen05 heads (boards <= 96 cells) and the fp32 heads kernel, at env counts around a CTA's group size
    for (m, n, k, ne) in [(9, 9, 5, 1), (9, 9, 5, 13), (9, 9, 5, 149), (7, 7, 4, 33), (3, 3, 3, 50), (10, 10, 5, 25)]:
        env = TorchVectorMnkEnv(m, n, k, ne, device="cuda")
        env.reset()
        for t in range(m * n // 3):
            env.step_autoreset(env.random_legal_actions(6, t), materialise=False)
        net = ResNetActorCritic((2, m, n), m * n).cuda()
        for mode in ("train", "eval"):
            fwd = NativeResNet(net, bn_mode=mode)
            fwd.forward_env(env, swap=(torch.arange(ne, device="cuda") % 2).to(torch.uint8))
            fwd.use_mma_heads = False
            fwd.forward_env(env)
            fwd.use_rows_kernel = False
            fwd.forward_env(env)
            fwd.check_error()
    # the wider convolutional towers
    for arch, (m, n, k, ne) in [("resnet_b_l", (9, 9, 5, 7)), ("resnet_b_l", (13, 13, 5, 3)), ("resnet_b_l", (19, 19, 5, 2)),
                                ("cnn_b_s", (9, 9, 5, 11)), ("cnn_b_l", (9, 9, 5, 4)), ("cnn_b_l", (5, 6, 4, 30))]:
        env = TorchVectorMnkEnv(m, n, k, ne, device="cuda")
        env.reset()
        for t in range(m * n // 3):
            env.step_autoreset(env.random_legal_actions(7, t), materialise=False)
        wide = NativeConvNet(build_architecture(arch, (2, m, n), m * n).cuda().eval())
        wide.forward_env(env, swap=(torch.arange(ne, device="cuda") % 2).to(torch.uint8))
        wide.check_error()
    # the transformers: one board per CTA (9x9), several (3x3: 14, 7x7: 2), all 128 token rows (8x16)
    for arch, (m, n, k, ne) in [("transformer_b_s", (9, 9, 5, 5)), ("transformer_b_s", (3, 3, 3, 31)), ("transformer_b_l", (7, 7, 4, 9)),
                                ("transformer_b_l", (8, 16, 5, 3))]:
        env = TorchVectorMnkEnv(m, n, k, ne, device="cuda")
        env.reset()
        for t in range(m * n // 3):
            env.step_autoreset(env.random_legal_actions(8, t), materialise=False)
        tfm = native_network(build_architecture(arch, (2, m, n), m * n).cuda().eval())
        tfm.forward_env(env, swap=(torch.arange(ne, device="cuda") % 2).to(torch.uint8))
        tfm.check_error()
    torch.cuda.synchronize()


if __name__ == "__main__":
    exercise_all()
    print("sanitize_all ok")

"""tcgen05 forward (mnk_resnet_tower + torch head tails) against the reference network's outputs.

Tolerance (BASELINE.json: "policy logits must match within 1e-3 relative in bf16").  The kernel
keeps activations as 16-bit floats between layers (IEEE fp16 operands by default: same tensor-core rate as
bf16, 8x finer rounding; fp32 accumulation in TMEM, fp32 heads); the fixture is the reference network in
fp32.  Asserted on the normalised masked logits ELEMENT-WISE: max |delta| <= 1e-3 * max |logit|, relative L2
error <= 3e-4, -inf positions identical, and the kernel closer to the fp32 reference than stock PyTorch bf16
autocast of the same network.  (The bf16 build, -DMNK_ACT_BF16, measured 2.5e-3 / 8.1e-4 in round 1 and gets
the looser bounds.)"""
import numpy as np
import pytest
import torch

import golden_io as gio
from test_nets_cpu import load_net

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("path", gio.files("resnet_b_s_"), ids=gio.name)
def test_native_forward_matches_reference(path):
    from mnk_b200 import NativeResNet
    g = gio.load(path)
    net, m, n, batch = load_net(g)
    native = NativeResNet(net.to(DEV), device=DEV)
    obs = torch.from_numpy(gio.unpack(g["obs"], (2, m, n)).astype(np.float32)).to(DEV)
    mask = torch.from_numpy(gio.unpack(g["mask"], (m * n,))).to(DEV)
    dist, value = native(obs, mask)
    native.check_error()
    want = g["logits"]
    got = dist.logits.cpu().numpy()
    fin = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), fin)
    scale = np.abs(want[fin]).max()
    err = np.abs(got[fin] - want[fin])
    rel_l2 = np.linalg.norm(got[fin] - want[fin]) / np.linalg.norm(want[fin])
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        d2, v2 = net(obs, mask)
    ac = d2.logits.float().cpu().numpy()
    ac_err = np.abs(ac[fin] - want[fin]).max()
    print(f"{gio.name(path)}: max|d|={err.max():.3e} (scale {scale:.2f}) rel_l2={rel_l2:.3e} autocast max|d|={ac_err:.3e} "
          f"value max|d|={np.abs(value.cpu().numpy() - g['value']).max():.3e}")
    from mnk_b200.resnet import operand_dtype
    f16 = operand_dtype() == torch.float16
    assert rel_l2 <= (3e-4 if f16 else 1e-3) and err.max() <= (1e-3 if f16 else 5e-3) * scale
    assert err.max() <= ac_err
    assert np.abs(value.cpu().numpy() - g["value"]).max() <= (4e-3 if f16 else 3e-2)
    # argmax agreement where the reference's top-2 gap exceeds the error bound
    top2 = np.sort(np.where(fin, want, -np.inf), axis=1)[:, -2:]
    clear = (top2[:, 1] - top2[:, 0]) > 2 * err.max()
    assert (np.argmax(np.where(fin, got, -np.inf), 1) == np.argmax(np.where(fin, want, -np.inf), 1))[clear].all()


@pytest.mark.parametrize("m,n,k", [(3, 3, 3), (5, 7, 4), (7, 7, 4), (9, 9, 5), (10, 10, 5), (6, 22, 5), (4, 15, 4)], ids=str)
def test_board_row_kernel_matches_tap_kernel_and_fp32_tower(m, n, k):
    """mnk_resnet_tower_rows (vertical taps fused into N = 96, TMEM slot ring, no per-layer barrier) against
    mnk_resnet_tower (one N = 32 MMA per tap) and the torch fp32 tower on mid-game positions: env counts that
    leave a partial last CTA, plane swap, lane layouts that fill all 128 lanes (7x7: 16 envs x 8 lanes)."""
    from mnk_b200 import NativeResNet, ResNetActorCritic, TorchVectorMnkEnv
    torch.manual_seed(11)
    net = ResNetActorCritic((2, m, n), m * n).to(DEV).eval()
    with torch.no_grad():
        for mod in net.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.normal_(0, 0.1)
                mod.running_var.uniform_(0.7, 1.3)
                mod.weight.uniform_(0.8, 1.2)
                mod.bias.normal_(0, 0.1)
    native = NativeResNet(net, device=DEV)
    per_cta = 128 // (n + 1)
    for ne in (1, per_cta, per_cta + 1, 5 * per_cta - 1, 1000):
        env = TorchVectorMnkEnv(m, n, k, ne, device=DEV)
        env.reset()
        for t in range(m * n // 2):
            env.step_autoreset(env.random_legal_actions(5, t), materialise=False)
        swap = (torch.arange(ne, device=DEV) % 3 == 0).to(torch.uint8)
        native.use_rows_kernel = True
        pf_r, vf_r = native.features(env._st, ne, m * n, swap)
        native.use_rows_kernel = False
        pf_t, vf_t = native.features(env._st, ne, m * n, swap)
        native.check_error()
        obs = env.observe()["observation"]
        obs = torch.where(swap.bool()[:, None, None, None], obs.flip(1), obs)
        with torch.no_grad():
            feat = net.forward_body(obs)
            want_p = net.policy_head[1](net.policy_head[0](feat))
            want_v = net.value_head[1](net.value_head[0](feat))
        for got, tap, want in ((pf_r, pf_t, want_p), (vf_r, vf_t, want_v)):
            scale = float(want.abs().max()) + 1e-6
            # both kernels round activations to bf16 between layers; their fp32 accumulation orders differ
            assert float((got - tap).abs().max()) <= 2e-2 * scale, (ne, float((got - tap).abs().max()), scale)
            rel = lambda x: float((x - want).norm() / (want.norm() + 1e-12))
            assert rel(got) <= 2e-2 and rel(got) <= 1.5 * rel(tap) + 1e-3, (ne, rel(got), rel(tap))   # bf16 activations, 9 layers
            assert float((got - want).abs().max()) <= 5e-2 * scale


def test_refresh_keeps_device_addresses_and_takes_new_weights():
    """NativeResNet.refresh copies into the existing device tensors (raw pointers baked into a captured rollout
    graph stay valid and read the NEW weights); the result equals a freshly built NativeResNet."""
    from mnk_b200 import NativeResNet, ResNetActorCritic, TorchVectorMnkEnv
    torch.manual_seed(8)
    net = ResNetActorCritic((2, 9, 9), 81).to(DEV).eval()
    native = NativeResNet(net, device=DEV)
    env = TorchVectorMnkEnv(9, 9, 5, 100, device=DEV)
    env.reset()
    for t in range(20):
        env.step_autoreset(env.random_legal_actions(5, t), materialise=False)
    before = native.forward_env(env)[0].clone()
    sig = native.pointer_signature()
    with torch.no_grad():
        for prm in net.parameters():
            prm.add_(0.02 * torch.randn_like(prm))
    native.refresh(net)
    assert native.pointer_signature() == sig
    after, value = native.forward_env(env)
    fresh = NativeResNet(net, device=DEV)
    want, want_value = fresh.forward_env(env)
    assert torch.equal(after, want) and torch.equal(value, want_value) and not torch.equal(after, before)
    assert fresh.pointer_signature() != sig


def test_native_forward_from_env_bitboards_and_batch_tails():
    """forward_env reads the bitboards directly (no f32 observation) and must agree with forward(obs):
    same kernel, different input path; batch sizes that leave a partial last CTA."""
    from mnk_b200 import NativeResNet, ResNetActorCritic, TorchSelfPlayWrapper, TorchVectorMnkEnv, RandomPolicy
    torch.manual_seed(3)
    net = ResNetActorCritic((2, 9, 9), 81).to(DEV).eval()
    native = NativeResNet(net, device=DEV)
    for ne in (1, 9, 10, 11, 333):
        env = TorchVectorMnkEnv(9, 9, 5, ne, device=DEV)
        wr = TorchSelfPlayWrapper(env, seed=ne)
        wr.set_opponent(RandomPolicy(81))
        obs, _ = wr.reset()
        for step in range(7):
            obs, *_ = wr.step(env.random_legal_actions(1, step))
        logits_env, value_env = native.forward_env(env, swap=wr._side)
        dist, value = native(obs["observation"], obs["action_mask"])
        assert torch.equal(logits_env, dist._raw) and torch.equal(value_env, value)
        with torch.no_grad():
            ref_dist, ref_value = net(obs["observation"], obs["action_mask"])
        assert (ref_dist._raw - logits_env).abs().max().item() <= 3e-2 * max(ref_dist._raw.abs().max().item(), 1e-3) + 1e-4
    native.check_error()


def test_native_nn_policy_in_wrapper():
    from mnk_b200 import NativeNNPolicy, ResNetActorCritic, TorchSelfPlayWrapper, TorchVectorMnkEnv
    torch.manual_seed(4)
    net = ResNetActorCritic((2, 9, 9), 81).to(DEV)
    pol = NativeNNPolicy(net, device=DEV, seed=2)
    env = TorchVectorMnkEnv(9, 9, 5, 640, device=DEV)
    wr = TorchSelfPlayWrapper(env, seed=1)
    wr.set_opponent(pol)
    obs, _ = wr.reset()
    done = 0
    for _ in range(60):
        a = pol.act(obs)
        assert bool(obs["action_mask"].gather(1, a[:, None]).all())
        obs, r, term, _, _ = wr.step(a)
        done += int(term.sum())
    pol.net.check_error()
    assert done > 100


def test_native_rollout_matches_generic_rollout_path():
    """RolloutCollector on the bitboard-fed path (tcgen05 forward, no f32 observation anywhere) stores the
    same observations / actions / rewards as the generic path driven by the same NativeResNet through
    its module-compatible forward(obs, mask) -- same seeds, same Philox counters."""
    from mnk_b200 import (NativeNNPolicy, NativeResNet, ResNetActorCritic, RolloutBuffer, RolloutCollector,
                          TorchSelfPlayWrapper, TorchVectorMnkEnv)
    torch.manual_seed(5)
    net = ResNetActorCritic((2, 9, 9), 81).to(DEV).eval()
    with torch.no_grad():
        net.policy_head[7].weight.mul_(50.0)
    opp_net = ResNetActorCritic((2, 9, 9), 81).to(DEV).eval()
    ne, steps = 700, 24
    runs = []
    for native_path in (True, False):
        env = TorchVectorMnkEnv(9, 9, 5, ne, device=DEV)
        wr = TorchSelfPlayWrapper(env, seed=9)
        opp = NativeNNPolicy(opp_net, device=DEV, seed=4)
        if not native_path:
            class ObsOnly:          # hide act_from_env: forces the f32 opponent view + forward(obs, mask)
                def __init__(self, inner): self.inner = inner; self.calls = 0
                def act(self, obs):
                    self.calls += 1
                    dist, _ = self.inner.net.forward(obs["observation"], obs["action_mask"])
                    from mnk_b200 import masked_sample
                    return masked_sample(dist._raw, dist._mask, seed=self.inner.seed, counter=wr._steps, want_log_prob=False)[0]
            wr.set_opponent(ObsOnly(opp))
        else:
            wr.set_opponent(opp)
        agent = NativeResNet(net, device=DEV)
        buf = RolloutBuffer(steps, ne, (2, 9, 9), 81, device=DEV, k=5)
        col = RolloutCollector(ne, device=DEV, seed=21)
        if native_path:
            wr.reset(materialise=False)
            col._last_obs = {"observation": None, "action_mask": None}
            stats = col.collect(agent, wr, buf)
        else:
            stats = col.collect(agent.forward, wr, buf)
        agent.check_error()
        runs.append((buf.observations.clone(), buf.actions.clone(), buf.rewards.clone(), buf.dones.clone(),
                     buf.log_probs.clone(), buf.values.clone(), stats))
    a, b = runs
    for x, y in zip(a[:4], b[:4]):
        assert torch.equal(x, y)
    assert torch.allclose(a[4], b[4], atol=1e-6) and torch.allclose(a[5], b[5], atol=1e-6)
    assert a[6].episodes == b[6].episodes > 0 and a[6].wins == b[6].wins


@pytest.mark.parametrize("m,n", [(9, 9), (13, 13), (15, 15), (19, 19), (5, 7), (20, 20)], ids=lambda v: str(v))
def test_fused_heads_match_torch_modules(m, n):
    """mnk_resnet_heads (LN -> ReLU -> Linear -> LN -> ReLU -> Linear [-> Tanh], fp32) against the same torch
    modules on the same tower features, including a row count that is not a multiple of the 8-sample batch."""
    from mnk_b200 import NativeResNet, ResNetActorCritic
    torch.manual_seed(7)
    cells = m * n
    net = ResNetActorCritic((2, m, n), cells).to(DEV).eval()
    with torch.no_grad():
        for mod in net.modules():
            if isinstance(mod, torch.nn.LayerNorm):
                mod.weight.uniform_(0.5, 1.5)
                mod.bias.normal_(0, 0.3)
            elif isinstance(mod, torch.nn.Linear):
                mod.bias.normal_(0, 0.2)
        net.policy_head[7].weight.mul_(40.0)
    native = NativeResNet(net, device=DEV)
    native.use_mma_heads = False
    for rows in (1, 8, 13, 1001):
        pf = torch.randn(rows, 2 * cells, device=DEV) * 2
        vf = torch.randn(rows, cells, device=DEV) * 2
        logits, values = native.tails(pf, vf)
        with torch.no_grad():
            want_l, want_v = native.policy_tail(pf), native.value_tail(vf)
        assert logits.shape == want_l.shape and values.shape == want_v.shape
        assert torch.allclose(logits, want_l, rtol=1e-4, atol=2e-4), (logits - want_l).abs().max()
        assert torch.allclose(values, want_v, rtol=1e-4, atol=1e-5), (values - want_v).abs().max()
        only_l, none_v = native.tails(pf, vf, want_value=False)        # policy-only (values = NULL in the C ABI)
        assert none_v is None and torch.equal(only_l, logits)


@pytest.mark.parametrize("m,n", [(9, 9), (3, 3), (5, 7), (8, 12), (6, 6)], ids=lambda v: str(v))
def test_tensor_core_heads_match_torch_modules(m, n):
    """mnk_resnet_heads_mma (the three Linear layers as tcgen05 GEMMs over 128-sample tiles, 16-bit operands, fp32
    accumulation / LayerNorm / bias) against the torch modules in fp32 and against the fp32 heads kernel: row counts
    with a partial tile and several tiles per CTA, policy-only calls; element-wise error bound 1e-3 of the logit range
    (the budget the tower + heads share in test_native_forward_matches_reference)."""
    from mnk_b200 import NativeResNet, ResNetActorCritic
    torch.manual_seed(7)
    cells = m * n
    net = ResNetActorCritic((2, m, n), cells).to(DEV).eval()
    with torch.no_grad():
        for mod in net.modules():
            if isinstance(mod, torch.nn.LayerNorm):
                mod.weight.uniform_(0.5, 1.5)
                mod.bias.normal_(0, 0.3)
            elif isinstance(mod, torch.nn.Linear):
                mod.bias.normal_(0, 0.2)
        net.policy_head[7].weight.mul_(40.0)
    native = NativeResNet(net, device=DEV)
    assert native.use_mma_heads and "hm_w2" in native._params
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        for rows in (1, 127, 128, 129, 1001, 148 * 128 * 2 + 77):
            pf = torch.randn(rows, 2 * cells, device=DEV) * 2
            vf = torch.randn(rows, cells, device=DEV) * 2
            logits, values = native.tails(pf, vf)
            native.check_error()
            with torch.no_grad():
                want_l, want_v = native.policy_tail(pf), native.value_tail(vf)
            assert logits.shape == want_l.shape and values.shape == want_v.shape
            scale = float(want_l.abs().max())
            err, rel = float((logits - want_l).abs().max()), float((logits - want_l).norm() / want_l.norm())
            verr = float((values - want_v).abs().max())
            if rows in (1, 1001):
                print(f"{m}x{n} rows={rows}: logits max|d| {err:.2e} (scale {scale:.1f}) rel_l2 {rel:.2e}; value max|d| {verr:.2e}")
            assert err <= 1e-3 * scale and rel <= 4e-4 and verr <= 2e-3, (rows, err, scale, rel, verr)
            only_l, none_v = native.tails(pf, vf, want_value=False)    # policy-only (values = NULL in the C ABI)
            assert none_v is None and torch.equal(only_l, logits)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32


@pytest.mark.parametrize("opponent_kind", ["native_nn", "random"])
def test_graph_captured_rollout_equals_eager_rollout(opponent_kind):
    """RolloutCollector.collect(graph=True) captures the whole rollout as one CUDA graph.  With the device
    counter base at zero (first replay) it must reproduce the eager rollout bit for bit; the second replay
    continues the same games with fresh random numbers."""
    from mnk_b200 import (NativeNNPolicy, NativeResNet, RandomPolicy, ResNetActorCritic, RolloutBuffer, RolloutCollector,
                          TorchSelfPlayWrapper, TorchVectorMnkEnv)
    torch.manual_seed(11)
    net = ResNetActorCritic((2, 9, 9), 81).to(DEV).eval()
    with torch.no_grad():
        net.policy_head[7].weight.mul_(50.0)
    ne, steps = 384, 20                      # the reference's default env count: launch-bound when run eagerly
    runs = []
    for use_graph in (False, True):
        env = TorchVectorMnkEnv(9, 9, 5, ne, device=DEV)
        wr = TorchSelfPlayWrapper(env, seed=3)
        wr.set_opponent(NativeNNPolicy(net, device=DEV, seed=8) if opponent_kind == "native_nn" else RandomPolicy(81))
        agent = NativeResNet(net, device=DEV)
        buf = RolloutBuffer(steps, ne, (2, 9, 9), 81, device=DEV, k=5)
        col = RolloutCollector(ne, device=DEV, seed=5)
        s1 = col.collect(agent, wr, buf, graph=use_graph)
        first = (buf.observations.clone(), buf.actions.clone(), buf.rewards.clone(), buf.dones.clone(), buf.log_probs.clone(),
                 s1.episodes, s1.wins)
        buf.reset()
        s2 = col.collect(agent, wr, buf, graph=use_graph)
        second = (buf.observations.clone(), buf.actions.clone(), s2.episodes)
        agent.check_error()
        runs.append((first, second, env.state_checksum()))
    eager, graphed = runs
    for x, y in zip(eager[0][:5], graphed[0][:5]):
        assert torch.equal(x, y)
    assert eager[0][5:] == graphed[0][5:] and eager[0][5] > 0
    # second rollout: continues from the first one's final state (same first observation as the eager run) ...
    assert torch.equal(eager[1][0][0], graphed[1][0][0])
    # ... with fresh draws: not a repeat of the first rollout's actions
    assert not torch.equal(graphed[1][1], graphed[0][1])
    assert graphed[1][2] > 0

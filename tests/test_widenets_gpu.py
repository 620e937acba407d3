"""mnk_conv_tower (csrc/mnk_convtower.cu) behind NativeConvNet against the reference's wider convolutional networks.

Fixtures (tests/golden/widenet_*.npz) hold the parameters and the eval-mode outputs of the UNMODIFIED reference classes
(src/alg/architectures/configs.py:36-65); the kernel keeps 16-bit activations between layers (fp16 operands, fp32
accumulation), so the tolerance is the path's logit bound (BASELINE.json: 1e-3 relative): asserted ELEMENT-WISE on the
normalised masked logits as max |delta| <= 1e-3 * max |logit| and as relative L2 <= 3e-4, with -inf positions identical
and the kernel closer to the fp32 reference than stock bf16 autocast of the same module."""
import numpy as np
import pytest
import torch

import golden_io as gio
from test_widenets_cpu import load_wide

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("path", gio.files("widenet_"), ids=gio.name)
def test_native_forward_matches_reference(path):
    from mnk_b200 import native_network
    from mnk_b200.resnet import operand_dtype
    g = gio.load(path)
    net, m, n, batch = load_wide(g)
    native = native_network(net.to(DEV), device=DEV)       # NativeConvNet / NativeTransformer by the module's layout
    obs = torch.from_numpy(gio.unpack(g["obs"], (2, m, n)).astype(np.float32)).to(DEV)
    mask = torch.from_numpy(gio.unpack(g["mask"], (m * n,))).to(DEV)
    dist, value = native(obs, mask)
    native.check_error()
    want, got = g["logits"], dist.logits.cpu().numpy()
    fin = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), fin)
    scale = np.abs(want[fin]).max()
    err = np.abs(got[fin] - want[fin])
    rel_l2 = np.linalg.norm(got[fin] - want[fin]) / np.linalg.norm(want[fin])
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        d2, _ = net(obs, mask)
    ac_err = np.abs(d2.logits.float().cpu().numpy()[fin] - want[fin]).max()
    v_err = np.abs(value.cpu().numpy() - g["value"]).max()
    print(f"{gio.name(path)}: max|d|={err.max():.3e} (scale {scale:.2f}) rel_l2={rel_l2:.3e} autocast max|d|={ac_err:.3e} "
          f"value max|d|={v_err:.3e}")
    f16 = operand_dtype() == torch.float16
    assert rel_l2 <= (3e-4 if f16 else 2e-3) and err.max() <= (1e-3 if f16 else 8e-3) * scale
    assert err.max() <= ac_err
    assert v_err <= (4e-3 if f16 else 3e-2)


def _randomise(net):
    with torch.no_grad():
        for mod in net.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.normal_(0, 0.1)
                mod.running_var.uniform_(0.7, 1.3)
                mod.weight.uniform_(0.8, 1.2)
                mod.bias.normal_(0, 0.1)
    return net


@pytest.mark.parametrize("arch,m,n,k", [("resnet_b_l", 9, 9, 5), ("resnet_b_l", 13, 13, 5), ("resnet_b_l", 19, 19, 5),
                                         ("resnet_b_l", 5, 7, 4), ("cnn_b_s", 9, 9, 5), ("cnn_b_s", 3, 3, 3),
                                         ("cnn_b_s", 13, 13, 5), ("cnn_b_l", 9, 9, 5), ("cnn_b_l", 6, 22, 5),
                                         ("resnet_s", 9, 9, 5), ("cnn_s", 7, 7, 4)], ids=str)
def test_features_match_fp32_module_from_bitboards(arch, m, n, k):
    """The kernel's head features (Flatten(Conv2d(C, 2, 1)(body)), Flatten(Conv2d(C, 1, 1)(body))) straight from the env's
    bitboards against the torch fp32 module on the same mid-game positions: env counts that leave a partial last CTA, one
    CTA and many, the plane swap of a white mover."""
    from mnk_b200 import NativeConvNet, TorchVectorMnkEnv, build_architecture
    torch.backends.cudnn.allow_tf32 = False          # the yardstick is the module in true fp32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(5)
    net = _randomise(build_architecture(arch, (2, m, n), m * n)).to(DEV).eval()
    native = NativeConvNet(net, device=DEV)
    ph, vh = (net.policy_head, net.value_head) if hasattr(net, "policy_head") else (net.actor, net.critic)
    for ne in (1, 4, 7, 333):
        env = TorchVectorMnkEnv(m, n, k, ne, device=DEV)
        env.reset()
        for t in range(m * n // 2):
            env.step_autoreset(env.random_legal_actions(5, t), materialise=False)
        swap = (torch.arange(ne, device=DEV) % 3 == 0).to(torch.uint8)
        pf, vf = native.features(env._st, ne, m * n, swap)
        native.check_error()
        obs = env.observe()["observation"]
        obs = torch.where(swap.bool()[:, None, None, None], obs.flip(1), obs)
        with torch.no_grad():
            body = net.forward_body(obs)
            want_pf, want_vf = ph[1](ph[0](body)), vh[1](vh[0](body))
        for got, want in ((pf, want_pf), (vf, want_vf)):
            scale = float(want.abs().max())
            assert float((got - want).abs().max()) <= 2e-3 * scale + 1e-4, (arch, m, n, ne)
        logits, values = native.forward_env(env, swap)
        with torch.no_grad():
            d, v = net(obs, None)
        assert float((logits - d._raw).abs().max()) <= 2e-3 * float(d._raw.abs().max()) + 1e-4
        assert float((values - v).abs().max()) <= 5e-3


def test_unsupported_shapes_raise():
    from mnk_b200 import NativeConvNet, TorchVectorMnkEnv, build_architecture, native_network, NativeResNet, ResNetActorCritic
    with pytest.raises(ValueError):
        NativeConvNet(build_architecture("resnet_l", (2, 9, 9), 81).to(DEV))       # 128 channels
    with pytest.raises(ValueError):
        NativeConvNet(ResNetActorCritic((2, 9, 9), 81).to(DEV))                     # 32 channels: NativeResNet's network
    assert isinstance(native_network(ResNetActorCritic((2, 9, 9), 81).to(DEV)), NativeResNet)
    wide = native_network(build_architecture("cnn_b_l", (2, 19, 19), 361).to(DEV).eval())
    assert isinstance(wide, NativeConvNet)
    env = TorchVectorMnkEnv(19, 19, 5, 4, device=DEV)
    env.reset()
    with pytest.raises(ValueError):             # 19x19 needs 401 pixel rows; the 96-channel tile has 384
        wide.forward_env(env)


def test_native_policy_on_wide_network_plays_legal_moves():
    """NativeNNPolicy picks NativeConvNet for a wide network; as the wrapper's opponent it only plays legal cells."""
    from mnk_b200 import NativeConvNet, NativeNNPolicy, TorchSelfPlayWrapper, TorchVectorMnkEnv, build_architecture
    torch.manual_seed(2)
    opp = NativeNNPolicy(build_architecture("resnet_b_l", (2, 9, 9), 81).to(DEV), seed=3)
    assert isinstance(opp.net, NativeConvNet)
    env = TorchVectorMnkEnv(9, 9, 5, 512, device=DEV, strict=True)
    wr = TorchSelfPlayWrapper(env, seed=1)
    wr.set_opponent(opp)
    obs, _ = wr.reset()
    for t in range(60):
        act = torch.multinomial(obs["action_mask"].float(), 1).squeeze(1)
        obs, r, term, trunc, _ = wr.step(act)
    opp.net.check_error()
    assert bool(((r == 0) | (r == 1) | (r == -1)).all())


@pytest.mark.parametrize("arch,m,n,k", [("transformer_b_s", 9, 9, 5), ("transformer_b_s", 3, 3, 3), ("transformer_b_s", 7, 7, 4),
                                         ("transformer_b_s", 10, 10, 5), ("transformer_b_s", 8, 16, 5), ("transformer_b_l", 9, 9, 5),
                                         ("transformer_b_l", 5, 7, 4), ("transformer_b_l", 11, 11, 5)], ids=str)
def test_transformer_features_match_fp32_module_from_bitboards(arch, m, n, k):
    """mnk_transformer_body straight from the env's bitboards against the torch fp32 module: one to fourteen boards per
    CTA (block-diagonal attention mask), partial last CTAs, boards that fill all 128 token rows (8x16), the plane swap."""
    from mnk_b200 import NativeTransformer, TorchVectorMnkEnv, build_architecture, native_network
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(6)
    net = build_architecture(arch, (2, m, n), m * n).to(DEV).eval()
    with torch.no_grad():
        net.pos_embed.normal_(0, 0.5)
        net.cell_embed.weight.normal_(0, 0.5)
        for l in net.transformer.layers:
            l.self_attn.in_proj_bias.normal_(0, 0.1)
            l.norm1.weight.uniform_(0.8, 1.2), l.norm2.bias.normal_(0, 0.1)
    native = native_network(net, device=DEV)
    assert isinstance(native, NativeTransformer)
    per_cta = 128 // (m * n)
    for ne in (1, per_cta, per_cta + 1, 3 * per_cta - 1, 301):
        env = TorchVectorMnkEnv(m, n, k, ne, device=DEV)
        env.reset()
        for t in range(m * n // 2):
            env.step_autoreset(env.random_legal_actions(5, t), materialise=False)
        swap = (torch.arange(ne, device=DEV) % 3 == 0).to(torch.uint8)
        pf, vf = native.features(env._st, ne, m * n, swap)
        native.check_error()
        obs = env.observe()["observation"]
        obs = torch.where(swap.bool()[:, None, None, None], obs.flip(1), obs)
        with torch.no_grad():
            body = net.forward_body(obs).transpose(1, 2)
            want_pf, want_vf = net.policy_head[1](net.policy_head[0](body)), net.value_head[1](net.value_head[0](body))
        for got, want in ((pf, want_pf), (vf, want_vf)):
            scale = float(want.abs().max())
            assert float((got - want).abs().max()) <= 3e-3 * scale + 1e-4, (arch, m, n, ne, float((got - want).abs().max()), scale)
        logits, values = native.forward_env(env, swap)
        with torch.no_grad():
            d, v = net(obs, None)
        assert float((logits - d._raw).abs().max()) <= 3e-3 * float(d._raw.abs().max()) + 1e-4
        assert float((values - v).abs().max()) <= 5e-3


def test_transformer_unsupported_shapes_raise():
    from mnk_b200 import NativeTransformer, TransformerActorCritic, build_architecture
    with pytest.raises(ValueError):
        NativeTransformer(TransformerActorCritic((2, 9, 9), 81, embed_dim=64, num_layers=2, num_heads=4).to(DEV))
    with pytest.raises(ValueError):
        NativeTransformer(build_architecture("transformer_b_s", (2, 13, 13), 169).to(DEV))      # 169 tokens > 128


@pytest.mark.parametrize("arch", ["resnet_b_l", "cnn_b_s", "transformer_b_s", "transformer_b_l"])
def test_native_forwards_are_deterministic(arch):
    """Back-to-back launches on the same state agree bit for bit (no atomics, fixed reduction orders) -- also a cheap race
    detector for the hand-rolled pipelines (weight rings, phase barriers)."""
    from mnk_b200 import TorchVectorMnkEnv, build_architecture, native_network
    torch.manual_seed(1)
    m, n, k, ne = 9, 9, 5, 3000
    native = native_network(build_architecture(arch, (2, m, n), m * n).to(DEV).eval(), device=DEV)
    env = TorchVectorMnkEnv(m, n, k, ne, device=DEV)
    env.reset()
    for t in range(30):
        env.step_autoreset(env.random_legal_actions(2, t), materialise=False)
    first = native.features(env._st, ne, m * n, None)
    for _ in range(20):
        again = native.features(env._st, ne, m * n, None)
        assert torch.equal(again[0], first[0]) and torch.equal(again[1], first[1])
    native.check_error()

"""Train-mode (batch-statistics) BatchNorm forward on the tcgen05 path: mnk_resnet_tower_train.

The reference's rollout forward never leaves train mode (src/alg/ppo.py:97), so its BatchNorm layers normalise with the
statistics of the current batch and update their running buffers.  Fixtures `resnet_train_b_s_*.npz` were recorded from
the UNMODIFIED reference network in .train() (oracle/gen_golden.py::gen_resnet_train): parameters before the call,
normalised masked logits / values, running statistics after the call.  Tolerances as for the eval-mode kernel:
element-wise |delta| <= 1e-3 * max |logit|, relative L2 <= 3e-4 (fp16 operands, fp32 accumulation and statistics)."""
import numpy as np
import pytest
import torch

import golden_io as gio
from test_nets_cpu import load_net

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _emulate(net, obs, op_dtype):
    """The kernel's arithmetic model in torch ops: operands (activations, conv weights) rounded to the 16-bit operand
    type, fp32 accumulation, batch statistics from the fp32 conv output, that output stored as fp16, BatchNorm applied
    as one fp32 scale / shift; the 1x1 head convolutions on the unrounded last activation."""
    import torch.nn.functional as F
    q = lambda t, dt: t.to(dt).float()
    convs = [(net.conv_in[0], net.conv_in[1])] + [cb for b in net.res_blocks for cb in ((b.conv1, b.bn1), (b.conv2, b.bn2))]
    a, skip, y = q(obs, op_dtype), None, None
    for L, (c, bn) in enumerate(convs):
        z = F.conv2d(a, q(c.weight, op_dtype), None, padding=1)
        mean, var = z.double().mean((0, 2, 3)), z.double().var((0, 2, 3), unbiased=False)
        scale = (bn.weight.double() / torch.sqrt(var + bn.eps)).float()
        shift = (bn.bias.double() - mean * scale.double()).float()
        y = q(z, torch.float16) * scale[None, :, None, None] + shift[None, :, None, None]
        if L >= 2 and L % 2 == 0:
            y = y + skip
        y = torch.relu(y)
        a = q(y, op_dtype)
        if L % 2 == 0:
            skip = a
    return net.policy_head[:2](y), net.value_head[:2](y)


def _bns(net):
    return [net.conv_in[1]] + [b for blk in net.res_blocks for b in (blk.bn1, blk.bn2)]


@pytest.mark.parametrize("path", gio.files("resnet_train_b_s_"), ids=gio.name)
def test_train_mode_forward_matches_reference_fixture(path):
    from mnk_b200 import NativeResNet
    from mnk_b200.resnet import operand_dtype
    g = gio.load(path)
    net, m, n, batch = load_net(g)
    net = net.to(DEV)
    native = NativeResNet(net, device=DEV, bn_mode="train")
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
    verr = np.abs(value.cpu().numpy() - g["value"]).max()
    print(f"{gio.name(path)}: max|d|={err.max():.3e} (scale {scale:.2f}) rel_l2={rel_l2:.3e} value max|d|={verr:.3e}")
    f16 = operand_dtype() == torch.float16
    assert rel_l2 <= (3e-4 if f16 else 1e-3) and err.max() <= (1e-3 if f16 else 5e-3) * scale
    assert verr <= (4e-3 if f16 else 3e-2)
    # running statistics after ONE train-mode forward, written back into the module
    native.export_running_stats(net)
    sd = net.state_dict()
    worst = 0.0
    for key in g:
        if not key.startswith("after/"):
            continue
        name = key[len("after/"):]
        have = sd[name].cpu().numpy()
        if "num_batches" in name:
            assert int(have) == int(g[key]) == 1
            continue
        tol = 2e-3 * max(np.abs(g[key]).max(), 1e-3)
        worst = max(worst, np.abs(have - g[key]).max() / max(np.abs(g[key]).max(), 1e-3))
        assert np.abs(have - g[key]).max() <= tol, (name, np.abs(have - g[key]).max())
    print(f"  running statistics: worst relative deviation {worst:.2e}")


@pytest.mark.parametrize("m,n,k,counts", [(9, 9, 5, (1, 12, 13, 1777, 5000)), (3, 3, 3, (7, 4000)), (5, 7, 4, (333,)),
                                          (10, 10, 5, (2500,)), (6, 22, 5, (1000,)), (7, 7, 4, (3000,)),
                                          (13, 13, 5, (1, 9, 10, 2000)), (11, 11, 5, (1500,)), (12, 7, 4, (3000,))], ids=str)
def test_train_mode_forward_matches_torch_module(m, n, k, counts):
    """Against the torch module (mnk_b200.nets.ResNetActorCritic == the reference network, tests/test_nets_cpu.py) in
    .train() on the GPU, on mid-game positions: env counts with a partial last group, one group per CTA and several
    groups per persistent CTA (double-buffered operand), plane swap; batch statistics and running buffers."""
    from mnk_b200 import NativeResNet, ResNetActorCritic, TorchVectorMnkEnv
    from mnk_b200.resnet import operand_dtype
    torch.manual_seed(13)
    net = ResNetActorCritic((2, m, n), m * n).to(DEV)
    with torch.no_grad():
        for mod in net.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.normal_(0, 0.1)
                mod.running_var.uniform_(0.7, 1.3)
                mod.weight.uniform_(0.8, 1.2)
                mod.bias.normal_(0, 0.1)
            elif isinstance(mod, torch.nn.Conv2d):
                mod.bias.normal_(0, 0.2)
        net.policy_head[7].weight.mul_(50.0)
    for ne in counts:
        env = TorchVectorMnkEnv(m, n, k, ne, device=DEV)
        env.reset()
        for t in range(m * n // 2):
            env.step_autoreset(env.random_legal_actions(5, t), materialise=False)
        swap = (torch.arange(ne, device=DEV) % 3 == 0).to(torch.uint8)
        native = NativeResNet(net, device=DEV, bn_mode="train")
        pf, vf = native.features(env._st, ne, m * n, swap)
        logits, value = native.tails(pf, vf)
        logits2, value2 = NativeResNet(net, device=DEV, bn_mode="train").forward_env(env, swap)
        native.check_error()
        assert torch.equal(logits, logits2) and torch.equal(value, value2)          # fixed-order statistics
        obs = env.observe()["observation"]
        obs = torch.where(swap.bool()[:, None, None, None], obs.flip(1), obs)
        ref = ResNetActorCritic((2, m, n), m * n).to(DEV)
        ref.load_state_dict(net.state_dict())
        ref.train()
        tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False      # a true fp32 reference
        try:
            with torch.no_grad():
                feat = ref.forward_body(obs)
                want_pf, want_vf = ref.policy_head[:2](feat), ref.value_head[:2](feat)
                want = ref.policy_head(feat)
                want_v = ref.value_head(feat)
                emu_pf, emu_vf = _emulate(net, obs, operand_dtype())
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
        scale = float(want.abs().max())
        err = float((logits - want).abs().max())
        rel = float((logits - want).norm() / want.norm())
        rel_pf = float((pf - want_pf).norm() / want_pf.norm())
        rel_vf = float((vf - want_vf).norm() / want_vf.norm())
        emu = max(float((pf - emu_pf).norm() / emu_pf.norm()), float((vf - emu_vf).norm() / emu_vf.norm()))
        print(f"{m}x{n} envs={ne}: head features rel_l2 {rel_pf:.3e} / {rel_vf:.3e} (vs the arithmetic model {emu:.3e}); "
              f"logits max|d|={err:.3e} (scale {scale:.2f}) rel_l2={rel:.3e}")
        # This default-initialised network's head features are differences of nearly cancelling terms (the fixtures above
        # pin the logits of the reference network at the 1e-3 bar): against fp32 the bound is that of the arithmetic
        # model (16-bit operands, fp16 pre-activations), against the model itself only summation order differs
        # (and with it which pre-activations sit on an fp16 rounding boundary).
        big = ne * m * n >= 500          # tiny batches: the variance of a handful of samples amplifies rounding
        assert rel_pf <= (3e-3 if big else 2e-2) and rel_vf <= (3e-3 if big else 2e-2), (ne, rel_pf, rel_vf)
        assert emu <= (1.5e-3 if big else 1e-2), (ne, emu)
        assert err <= (6e-3 if big else 3e-2) * scale, (ne, err, scale, rel)
        assert float((value - want_v).abs().max()) <= 1e-2
        # batch statistics and running buffers of every layer
        stats = native._params["batch_stats"]
        for i, (b_ref, b_old) in enumerate(zip(_bns(ref), _bns(net))):
            mean_new = (b_ref.running_mean - 0.9 * b_old.running_mean) / 0.1
            assert torch.allclose(stats[i, :32], mean_new, atol=2e-3 * float(mean_new.abs().max()) + 1e-4), i
            assert torch.allclose(native._params["running_mean"][i], b_ref.running_mean, atol=2e-3 * float(b_ref.running_mean.abs().max()) + 1e-4), i
            assert torch.allclose(native._params["running_var"][i], b_ref.running_var, rtol=1e-2, atol=1e-4), i


def test_train_mode_geometry_limits_and_eval_unchanged():
    from mnk_b200 import NativeResNet, ResNetActorCritic, TorchVectorMnkEnv
    net = ResNetActorCritic((2, 15, 15), 225).to(DEV)
    native = NativeResNet(net, device=DEV, bn_mode="train")
    env = TorchVectorMnkEnv(15, 15, 5, 10, device=DEV)
    env.reset()
    with pytest.raises(ValueError):      # boards with more than 13 rows: train-mode kernel not available (shared memory)
        native.forward_env(env)
    with pytest.raises(ValueError):
        NativeResNet(net, device=DEV, bn_mode="batch")

"""The stock-PyTorch ResNetActorCritic must be the reference's "resnet_b_s": it loads a reference
state_dict (golden fixture recorded from src/alg/architectures/configs.py::ResNetSActorCritic) and
reproduces the reference's eval-mode masked logits and values.  CPU, fp32."""
import numpy as np
import pytest
import torch

import golden_io as gio


def load_net(g):
    from mnk_b200.nets import ResNetActorCritic
    m, n, k, batch = (int(x) for x in g["geom"])
    net = ResNetActorCritic((2, m, n), m * n)
    sd = {key[len("param/"):]: torch.from_numpy(v) for key, v in g.items() if key.startswith("param/")}
    missing = net.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    net.eval()
    return net, m, n, batch


@pytest.mark.parametrize("path", gio.files("resnet_b_s_"), ids=gio.name)
def test_module_matches_reference_network(path):
    g = gio.load(path)
    net, m, n, batch = load_net(g)
    assert sum(p.numel() for p in net.parameters()) == (118203 if (m, n) == (9, 9) else sum(p.numel() for p in net.parameters()))
    obs = torch.from_numpy(gio.unpack(g["obs"], (2, m, n)).astype(np.float32))
    mask = torch.from_numpy(gio.unpack(g["mask"], (m * n,)))
    with torch.no_grad():
        dist, value = net(obs, mask)
    want = g["logits"]
    got = dist.logits.numpy()
    fin = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), fin)                      # -inf exactly where the reference has it
    assert np.allclose(got[fin], want[fin], rtol=0, atol=2e-5)
    assert np.allclose(value.numpy(), g["value"], rtol=0, atol=2e-6)
    assert np.allclose(got[0], -np.log(m * n), atol=1e-6)             # the all-masked row is uniform

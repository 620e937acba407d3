"""The stock-PyTorch modules of the wider convolutional architectures (mnk_b200.nets: "resnet_b_l", "cnn_b_s", "cnn_b_l")
must BE the reference's networks: they load a state_dict recorded from src/alg/architectures/configs.py:36-65 strictly
(same parameter names and shapes) and reproduce the reference's eval-mode masked logits and values.  CPU, fp32."""
import numpy as np
import pytest
import torch

import golden_io as gio
import ref_loader


def load_wide(g):
    from mnk_b200.nets import build_architecture
    m, n, k, batch = (int(x) for x in g["geom"])
    net = build_architecture(str(g["arch"]), (2, m, n), m * n)
    sd = {key[len("param/"):]: torch.from_numpy(v.astype(np.float32) if v.dtype == np.float16 else v)
          for key, v in g.items() if key.startswith("param/")}
    missing = net.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return net.eval(), m, n, batch


@pytest.mark.parametrize("path", gio.files("widenet_"), ids=gio.name)
def test_module_matches_reference_network(path):
    g = gio.load(path)
    net, m, n, batch = load_wide(g)
    assert net._architecture_name == str(g["arch"])
    obs = torch.from_numpy(gio.unpack(g["obs"], (2, m, n)).astype(np.float32))
    mask = torch.from_numpy(gio.unpack(g["mask"], (m * n,)))
    with torch.no_grad():
        dist, value = net(obs, mask)
    want, got = g["logits"], dist.logits.numpy()
    fin = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), fin)
    assert np.allclose(got[fin], want[fin], rtol=0, atol=5e-5)
    assert np.allclose(value.numpy(), g["value"], rtol=0, atol=5e-6)
    assert np.allclose(got[0], -np.log(m * n), atol=1e-6)             # the all-masked row is uniform


@pytest.mark.skipif(not ref_loader.available(), reason="reference not mounted (GPU box)")
@pytest.mark.parametrize("name", ["resnet_b_s", "resnet_b_l", "cnn_b_s", "cnn_b_l", "resnet_s", "cnn_s", "transformer_b_s",
                                  "transformer_b_l"])
def test_registry_matches_live_reference(name):
    """Every registry entry against the reference's own ARCHITECTURE_REGISTRY (src/utils/model_export.py:28-47): the same
    parameter names / shapes and, with the reference's weights loaded, the same outputs."""
    from mnk_b200.nets import build_architecture
    registry = ref_loader.load("utils.model_export").ARCHITECTURE_REGISTRY
    torch.manual_seed(3)
    ref = registry[name]((2, 6, 7), 42).eval()
    mine = build_architecture(name, (2, 6, 7), 42).eval()
    assert mine._architecture_name == ref._architecture_name and mine._architecture_params == ref._architecture_params
    assert {k: tuple(v.shape) for k, v in mine.state_dict().items()} == {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    mine.load_state_dict(ref.state_dict(), strict=True)
    obs = (torch.rand(5, 2, 6, 7) < 0.3).float()
    mask = torch.rand(5, 42) < 0.7
    mask[:, 0] = True
    with torch.no_grad():
        d_ref, v_ref = ref(obs, mask)
        d_mine, v_mine = mine(obs, mask)
    fin = torch.isfinite(d_ref.logits)
    assert torch.equal(torch.isfinite(d_mine.logits), fin)
    assert torch.allclose(d_mine.logits[fin], d_ref.logits[fin], atol=1e-5)
    assert torch.allclose(v_mine, v_ref, atol=1e-6)

"""Model files in the reference's format (src/utils/model_export.py): round trip here, and -- where the
reference is mounted -- files written by either side load in the other.  CPU only."""
import json
import os

import pytest
import torch

import ref_loader
from mnk_b200 import model_io
from mnk_b200.nets import ResNetActorCritic


def _net(m=5, n=4, seed=0):
    torch.manual_seed(seed)
    return ResNetActorCritic((2, m, n), m * n)


def test_round_trip(tmp_path):
    net = _net()
    mid = model_io.export_model(net, str(tmp_path / "run7"), iteration=42, is_benchmark_breaker=True)
    assert mid == "model_00042"
    meta = json.load(open(tmp_path / "run7" / "model_00042.json"))
    assert meta["architecture"] == {"name": "resnet_b_s", "params": {"obs_shape": [2, 5, 4], "action_dim": 20}}
    assert meta["run_name"] == "run7" and meta["is_benchmark_breaker"] is True and meta["iteration"] == 42
    back = model_io.load_model(str(tmp_path / "run7"), mid)
    assert not back.training
    for (ka, a), (kb, b) in zip(net.state_dict().items(), back.state_dict().items()):
        assert ka == kb and torch.equal(a, b)
    model_io.export_model(net, str(tmp_path / "run7"), iteration=3)
    (tmp_path / "run7" / "broken.json").write_text("{not json")
    assert [x["iteration"] for x in model_io.list_models(str(tmp_path / "run7"))] == [3, 42]
    assert model_io.list_models(str(tmp_path / "nope")) == []


def test_compiled_prefix_and_errors(tmp_path):
    net = _net(seed=1)
    d = str(tmp_path)
    mid = model_io.export_model(net, d, 1)
    torch.save({"_orig_mod." + k: v for k, v in net.state_dict().items()}, os.path.join(d, mid + ".pt"))
    back = model_io.load_model(d, mid)
    assert torch.equal(back.policy_head[7].weight, net.policy_head[7].weight)
    with pytest.raises(FileNotFoundError):
        model_io.load_model(d, "model_00009")
    meta = model_io.read_metadata(d, mid)
    meta["architecture"]["name"] = "transformer_c_l"      # (the sgrtransformer variants are not built here)
    json.dump(meta, open(os.path.join(d, mid + ".json"), "w"))
    with pytest.raises(ValueError, match="Unknown architecture"):
        model_io.load_model(d, mid)
    os.remove(os.path.join(d, mid + ".pt"))
    meta["architecture"]["name"] = "resnet_b_s"
    json.dump(meta, open(os.path.join(d, mid + ".json"), "w"))
    with pytest.raises(FileNotFoundError):
        model_io.load_model(d, mid)
    with pytest.raises(ValueError, match="_architecture_name"):
        model_io.export_model(torch.nn.Linear(2, 2), d, 5)


@pytest.mark.skipif(not ref_loader.available(), reason="reference not mounted (GPU box)")
def test_interchange_with_reference(tmp_path):
    ref_io, ref_cfg = ref_loader.load("utils.model_export", "alg.architectures.configs")
    # reference -> here
    torch.manual_seed(3)
    ref_net = ref_cfg.ResNetSActorCritic((2, 9, 9), 81)
    exporter = ref_io.ModelExporter(run_name="r", base_dir=str(tmp_path))
    mid = exporter.export_model(ref_net, iteration=7)
    mine = model_io.load_model(exporter.export_dir, mid)
    obs = (torch.rand(6, 2, 9, 9) < 0.2).float()
    mask = obs.sum(1).reshape(6, 81) == 0
    ref_net.eval()
    with torch.no_grad():
        d_ref, v_ref = ref_net(obs, mask)
        d_mine, v_mine = mine(obs, mask)
    assert torch.allclose(d_ref.logits, d_mine.logits, atol=1e-5) and torch.allclose(v_ref, v_mine, atol=1e-6)
    assert model_io.list_models(exporter.export_dir) == ref_io.get_models_from_directory(exporter.export_dir)
    # here -> reference
    mid2 = model_io.export_model(_net(9, 9, seed=4), exporter.export_dir, iteration=8)
    theirs = ref_io.load_any_model(exporter.export_dir, mid2)
    mine2 = model_io.load_model(exporter.export_dir, mid2)
    with torch.no_grad():
        d_a, v_a = theirs(obs, mask)
        d_b, v_b = mine2(obs, mask)
    assert torch.allclose(d_a.logits, d_b.logits, atol=1e-5) and torch.allclose(v_a, v_b, atol=1e-6)

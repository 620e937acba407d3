"""Model files in the reference's on-disk format, for the architecture this path accelerates.

The reference writes ``<dir>/model_00042.pt`` (a plain ``state_dict``) next to ``model_00042.json``
(``{"model_id", "iteration", "architecture": {"name", "params"}, "export_timestamp",
"is_benchmark_breaker", "run_name"}``) and reads them back by architecture name
(src/utils/model_export.py:87-178).  Files written here load in the reference and vice versa:
``ResNetActorCritic`` keeps the reference's parameter names, and keys saved from a
``torch.compile``-wrapped module (``_orig_mod.`` prefix, model_export.py:167-174) are accepted.

The convolutional families are constructed here (``mnk_b200.nets.ARCHITECTURES``: resnet_b_s -- the network of the
accelerated path -- resnet_b_l, cnn_b_s, cnn_b_l, transformer_b_s, transformer_b_l and the older resnet_s / resnet_l / cnn_s /
cnn_l); other names (transformer_s / _l, the sgrtransformer transformer_c_*) raise, so a caller falls back to the reference's
own loader for them rather than silently getting a different network.
"""
from __future__ import annotations

import json
import os
from datetime import datetime
from typing import Any, Dict, List, Optional

import torch

from .nets import ARCHITECTURES, build_architecture

ARCHITECTURE = "resnet_b_s"            # the default network (src/train.py)
_COMPILED_PREFIX = "_orig_mod."


def model_id_for(iteration: int) -> str:
    return f"model_{int(iteration):05d}"


def export_model(network: torch.nn.Module, export_dir: str, iteration: int, is_benchmark_breaker: bool = False,
                 run_name: Optional[str] = None) -> str:
    """Write ``network`` the way ``ModelExporter.export_model`` does (model_export.py:98-132); returns the model id."""
    name = getattr(network, "_architecture_name", None)
    params = getattr(network, "_architecture_params", None)
    if name is None or params is None:
        raise ValueError("Model must have _architecture_name and _architecture_params attributes for export")
    os.makedirs(export_dir, exist_ok=True)
    model_id = model_id_for(iteration)
    torch.save(network.state_dict(), os.path.join(export_dir, f"{model_id}.pt"))
    meta = {"model_id": model_id, "iteration": int(iteration), "architecture": {"name": name, "params": params},
            "export_timestamp": datetime.now().isoformat(), "is_benchmark_breaker": bool(is_benchmark_breaker),
            "run_name": run_name if run_name is not None else os.path.basename(os.path.normpath(export_dir))}
    with open(os.path.join(export_dir, f"{model_id}.json"), "w") as f:
        json.dump(meta, f, indent=2)
    return model_id


def read_metadata(model_dir: str, model_id: str) -> Dict[str, Any]:
    path = os.path.join(model_dir, f"{model_id}.json")
    if not os.path.exists(path):
        raise FileNotFoundError(f"Metadata for model {model_id} not found in {model_dir}")
    with open(path) as f:
        return json.load(f)


def strip_compiled_prefix(state_dict: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    return {(k[len(_COMPILED_PREFIX):] if k.startswith(_COMPILED_PREFIX) else k): v for k, v in state_dict.items()}


def load_model(model_dir: str, model_id: str, device: str = "cpu") -> torch.nn.Module:
    """``load_any_model`` (model_export.py:146-178) for the convolutional architectures: eval-mode module on ``device``."""
    meta = read_metadata(model_dir, model_id)
    arch = meta.get("architecture", {})
    if arch.get("name") not in ARCHITECTURES:
        raise ValueError(f"Unknown architecture: {arch.get('name')}. Known architectures: {', '.join(sorted(ARCHITECTURES))}")
    weights = os.path.join(model_dir, f"{model_id}.pt")
    if not os.path.exists(weights):
        raise FileNotFoundError(f"Model weights {model_id} not found in {model_dir}")
    model = build_architecture(arch["name"], **arch.get("params", {}))
    model.load_state_dict(strip_compiled_prefix(torch.load(weights, map_location=device)))
    return model.to(device).eval()


def list_models(model_dir: str) -> List[Dict[str, Any]]:
    """Metadata of every model in a directory, by iteration (model_export.py:181-204); unreadable files are skipped."""
    found: List[Dict[str, Any]] = []
    if not os.path.isdir(model_dir):
        return found
    for fn in os.listdir(model_dir):
        if not fn.endswith(".json"):
            continue
        try:
            with open(os.path.join(model_dir, fn)) as f:
                meta = json.load(f)
        except (json.JSONDecodeError, OSError):
            continue
        if "model_id" in meta:
            found.append(meta)
    found.sort(key=lambda x: x.get("iteration", 0))
    return found

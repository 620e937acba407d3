"""Imports modules of the (read-only) reference in isolation from the drop-in packages of the same
name (``env``, ``selfplay``).  Only usable where /root/reference is mounted (the build container)."""
import importlib
import os
import sys

REF = os.environ.get("MNK_REFERENCE", "/root/reference")
_SHADOWED = ("env", "selfplay", "alg", "utils")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "src", "env"))


def load(*names):
    """Return the reference's modules `names` (e.g. "env.torch_vector_mnk_env")."""
    def shadowed(k):
        return any(k == p or k.startswith(p + ".") for p in _SHADOWED)

    saved_mods = {k: sys.modules.pop(k) for k in list(sys.modules) if shadowed(k)}
    saved_path = list(sys.path)
    sys.path[:] = [os.path.join(REF, "src"), REF] + [p for p in saved_path if "rl-selfplay-mnk_b200" not in p]
    try:
        out = [importlib.import_module(n) for n in names]
    finally:
        for k in [k for k in sys.modules if shadowed(k)]:
            del sys.modules[k]
        sys.modules.update(saved_mods)
        sys.path[:] = saved_path
    return out if len(out) > 1 else out[0]

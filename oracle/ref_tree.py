"""Locates (and, in the build container, stages) the UNMODIFIED reference tree.

TEST / MEASUREMENT INFRASTRUCTURE ONLY -- never imported by the product (rl-selfplay-mnk_b200/).

The reference (michal-szadkowski/rl-selfplay-mnk) is a pure-Python program: there is nothing to
compile, so "building the real reference" means making its source tree available where the
tests and ``bench.py --impl reference`` run.  ``/root/reference`` exists only in the build
container; ``stage()`` (called from ``__graft_entry__.build()``) copies ``/root/reference/src``
byte for byte into ``oracle/_ref/src``.  ``oracle/_ref/`` is git-ignored (reference sources never
enter this repository's history) but not gpurun-ignored, so it travels to the GPU box next to
the built ``.so`` files.

``load(...)`` imports reference modules in isolation from the drop-in packages of the same name
(``env``, ``selfplay``, ``alg``); ``dropin_path()`` / ``reference_path()`` give the two
``sys.path`` orders used by the drop-in tests (drop-in first = the reference's callers run on
the sm_100a kernels; reference only = the stock implementation).
"""
from __future__ import annotations

import filecmp
import importlib
import os
import shutil
import sys
from contextlib import contextmanager

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "rl-selfplay-mnk_b200")
STAGED = os.path.join(ROOT, "oracle", "_ref")
MOUNTED = os.environ.get("MNK_REFERENCE", "/root/reference")
_SHADOWED = ("env", "selfplay", "alg", "utils", "model_comparison", "src")


def _has_src(base: str) -> bool:
    return os.path.isfile(os.path.join(base, "src", "env", "torch_vector_mnk_env.py"))


def stage(verbose: bool = False) -> str | None:
    """Copy <mounted reference>/src -> oracle/_ref/src (idempotent).  Returns the staged root or None
    when neither the mounted reference nor an earlier staged copy exists."""
    if _has_src(MOUNTED):
        dst = os.path.join(STAGED, "src")
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(MOUNTED, "src"), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        cmp = filecmp.dircmp(os.path.join(MOUNTED, "src"), dst, ignore=["__pycache__"])
        assert not cmp.diff_files and not cmp.left_only, "staged reference differs from the mounted one"
        if verbose:
            n = sum(len(f) for _, _, f in os.walk(dst))
            print(f"staged the unmodified reference: {MOUNTED}/src -> {dst} ({n} files)")
    return STAGED if _has_src(STAGED) else None


def root() -> str | None:
    """The reference root to use: the staged copy (what travels to the GPU box), else the mounted tree."""
    if _has_src(STAGED):
        return STAGED
    if _has_src(MOUNTED):
        return MOUNTED
    return None


def available() -> bool:
    return root() is not None


def kind() -> str:
    """"reference" when the unmodified tree is importable here, else "port" (oracle/torch_port.py)."""
    return "reference" if available() else "port"


def reference_path() -> list[str]:
    """sys.path entries under which every import resolves to the stock reference."""
    r = root()
    return [os.path.join(r, "src"), r]


def dropin_path() -> list[str]:
    """sys.path entries of INTEGRATION.md (A): the drop-in package directory ahead of the reference's src/."""
    r = root()
    return [PKG, os.path.join(r, "src"), r]


def _shadowed(name: str) -> bool:
    return any(name == p or name.startswith(p + ".") for p in _SHADOWED)


@contextmanager
def imports(paths: list[str]):
    """Temporarily resolve `env`, `selfplay`, `alg`, `utils`, `src` ... through `paths` only; modules
    imported inside the block are removed from sys.modules afterwards (objects stay usable)."""
    saved_mods = {k: sys.modules.pop(k) for k in list(sys.modules) if _shadowed(k)}
    saved_path = list(sys.path)
    foreign = {os.path.abspath(p) for p in (PKG, STAGED, os.path.join(STAGED, "src"), MOUNTED, os.path.join(MOUNTED, "src"))}
    sys.path[:] = list(paths) + [p for p in saved_path if os.path.abspath(p or ".") not in foreign]
    importlib.invalidate_caches()
    try:
        yield
    finally:
        for k in [k for k in sys.modules if _shadowed(k)]:
            del sys.modules[k]
        sys.modules.update(saved_mods)
        sys.path[:] = saved_path


def load(*names: str):
    """The stock reference's modules `names` (e.g. "env.torch_vector_mnk_env"), isolated from the drop-in."""
    if not available():
        raise RuntimeError("the reference tree is neither mounted nor staged under oracle/_ref")
    with imports(reference_path()):
        out = [importlib.import_module(n) for n in names]
    return out if len(out) > 1 else out[0]


def load_dropin(*names: str):
    """Modules `names` resolved the way INTEGRATION.md (A) sets PYTHONPATH: drop-in packages first, the
    rest of the reference's src/ (PPO, networks, launchers) unmodified behind them."""
    if not available():
        raise RuntimeError("the reference tree is neither mounted nor staged under oracle/_ref")
    with imports(dropin_path()):
        out = [importlib.import_module(n) for n in names]
    return out if len(out) > 1 else out[0]

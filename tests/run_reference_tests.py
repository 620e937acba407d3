"""Runs the reference's OWN test file, unmodified, in a fresh interpreter and reports each test's outcome.

    python tests/run_reference_tests.py {dropin|reference} OUT.json

The reference's tests import ``src.env.torch_vector_mnk_env`` / ``src.selfplay.torch_self_play_wrapper``
(src/tests/test_mnk_integration.py:3-4).  In `dropin` mode the package name ``src`` is bound to
rl-selfplay-mnk_b200/ (this repo's stand-in for the reference's src/ on the hot path) ahead of the
reference's real src/, so those two modules -- and everything they import -- are the sm_100a implementation;
in `reference` mode ``src`` is the reference's own directory.  The test file itself is read from the staged
(or mounted) reference tree and is never edited.
"""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(mode: str, out_path: str) -> int:
    import pytest
    from oracle import ref_tree
    ref = ref_tree.root()
    if ref is None:
        print("no reference tree (oracle/_ref not staged)")
        return 3
    src = types.ModuleType("src")
    if mode == "dropin":
        src.__path__ = [ref_tree.PKG, os.path.join(ref, "src")]
        sys.path[:0] = ref_tree.dropin_path()
    else:
        src.__path__ = [os.path.join(ref, "src")]
        sys.path[:0] = ref_tree.reference_path()
    sys.modules["src"] = src
    test_file = os.path.join(ref, "src", "tests", "test_mnk_integration.py")

    class Outcomes:
        def __init__(self):
            self.results = {}

        def pytest_runtest_logreport(self, report):
            name = report.nodeid.split("::")[-1]
            if report.when == "call" or (report.outcome != "passed" and name not in self.results):
                self.results[name] = {"outcome": report.outcome, "where": report.when,
                                      "detail": str(report.longrepr)[-600:] if report.outcome != "passed" else ""}

    plugin = Outcomes()
    rc = pytest.main([test_file, "-q", "-p", "no:cacheprovider", "--import-mode=importlib", "--rootdir",
                      os.path.join(ref, "src", "tests"), "-c", os.devnull], plugins=[plugin])
    import torch
    impl = sys.modules["src.env.torch_vector_mnk_env"].TorchVectorMnkEnv
    with open(out_path, "w") as f:
        json.dump({"mode": mode, "pytest_rc": int(rc), "results": plugin.results, "cuda": torch.cuda.is_available(),
                   "env_class": f"{impl.__module__}.{impl.__qualname__}",
                   "native_lib_loaded": any("libmnk_b200" in line for line in open("/proc/self/maps"))}, f)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1], sys.argv[2]))

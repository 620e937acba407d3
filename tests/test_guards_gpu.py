"""Out-of-bounds WRITE check of every kernel without compute-sanitizer (the GPU pool refuses it).

Every CUDA tensor the package allocates through torch.empty / torch.zeros while `tools/sanitize_all.exercise_all()` runs
-- observation / mask / reward / done outputs, packed state, rollout buffers, head features, logits, the train-mode
tower's scratch arrays, sampler outputs -- is carved out of a larger byte buffer with a 4 KiB canary zone on either
side (0xA5); the kernels receive the interior pointer through the normal code path.  After the run every canary byte
must be intact.  The exercise uses ragged sizes (tail tiles, partial CTAs, 1-env batches) on six board geometries."""
import math
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
GUARD = 4096
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class GuardedAllocations:
    def __init__(self):
        self.regions = []

    def _carve(self, shape, dtype, device, zero):
        numel = int(math.prod(shape))
        nbytes = numel * torch.empty((), dtype=dtype).element_size()
        raw = self._empty(nbytes + 2 * GUARD, dtype=torch.uint8, device=device)
        raw.fill_(0xA5)
        inner = raw[GUARD:GUARD + nbytes]
        if zero:
            inner.zero_()
        self.regions.append((raw, nbytes, tuple(shape), dtype))
        return inner.view(dtype).view(tuple(shape))

    def _wrap(self, orig, zero):
        def alloc(*size, **kw):
            device = kw.get("device")
            if device is None or torch.device(device).type != "cuda" or kw.get("pin_memory") or kw.get("out") is not None:
                return orig(*size, **kw)
            shape = size[0] if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)) else size
            dtype = kw.get("dtype") or torch.get_default_dtype()
            if dtype.is_complex or any(not isinstance(d, int) for d in shape):
                return orig(*size, **kw)
            return self._carve(shape, dtype, device, zero)
        return alloc

    def __enter__(self):
        self._empty, self._zeros = torch.empty, torch.zeros
        torch.empty, torch.zeros = self._wrap(self._empty, False), self._wrap(self._zeros, True)
        return self

    def __exit__(self, *exc):
        torch.empty, torch.zeros = self._empty, self._zeros

    def check(self):
        torch.cuda.synchronize()
        bad = []
        for raw, nbytes, shape, dtype in self.regions:
            lo, hi = raw[:GUARD], raw[GUARD + nbytes:]
            if not bool((lo == 0xA5).all()) or not bool((hi == 0xA5).all()):
                first_hi = int(torch.nonzero(hi != 0xA5)[0]) if not bool((hi == 0xA5).all()) else None
                bad.append((shape, dtype, "below" if not bool((lo == 0xA5).all()) else f"above (+{first_hi} B)"))
        return bad


def test_no_kernel_writes_outside_its_buffers():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sanitize_all
    with GuardedAllocations() as guards:
        sanitize_all.exercise_all()
        bad = guards.check()
    assert len(guards.regions) > 500, len(guards.regions)          # the wrapper really sat under the package's allocations
    assert not bad, bad[:10]


def test_the_canaries_catch_an_overrun():
    """The checker itself: a deliberate one-element overrun through the C ABI (mnk_observe told the batch is one env
    larger than the mask buffer) must be reported."""
    import ctypes
    from mnk_b200 import TorchVectorMnkEnv, _lib
    from mnk_b200._lib import MnkState
    env = TorchVectorMnkEnv(9, 9, 5, 33, device="cuda")
    env.reset()
    with GuardedAllocations() as guards:
        obs = torch.empty((33, 2, 9, 9), dtype=torch.float32, device="cuda")
        mask = torch.empty((32, 81), dtype=torch.bool, device="cuda")          # one row short
        st = env._st
        _lib.check(_lib.lib().mnk_observe(ctypes.byref(st), obs.data_ptr(), mask.data_ptr(), None, 0,
                                          torch.cuda.current_stream().cuda_stream), "mnk_observe")
        bad = guards.check()
    assert len(bad) == 1 and bad[0][0] == (32, 81) and bad[0][2].startswith("above")

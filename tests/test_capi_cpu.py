"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol that
include/mnk_b200.h declares, and validates arguments before touching CUDA."""
import ctypes
import os
import re

import pytest

from mnk_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mnk_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mnk_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    L = _lib.lib()
    names = declared_symbols()
    assert "mnk_step" in names and "mnk_observe" in names and len(names) >= 12
    for name in names:
        assert hasattr(L, name), f"{name} declared in include/mnk_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in mnk_b200/_lib.py"
    assert L.mnk_version() == 200       # MNK_B200_VERSION of include/mnk_b200.h (round 2)


def test_state_words():
    L = _lib.lib()
    assert L.mnk_state_words(3, 3) == 1       # 12 bits
    assert L.mnk_state_words(9, 9) == 2       # 90 bits
    assert L.mnk_state_words(13, 13) == 3     # 182 bits
    assert L.mnk_state_words(19, 19) == 6     # 380 bits
    assert L.mnk_state_words(7, 11) == 2
    assert L.mnk_state_words(40, 40) == _lib.MNK_ERR_GEOM
    assert L.mnk_state_words(3, 33) == _lib.MNK_ERR_GEOM
    assert L.mnk_state_words(0, 3) == _lib.MNK_ERR_GEOM


def test_argument_validation_without_gpu():
    L = _lib.lib()
    st = _lib.MnkState(9, 9, 5, 2, 4, None, None)
    assert L.mnk_reset(ctypes.byref(st), None, 0, None) == _lib.MNK_ERR_NULL
    buf = (ctypes.c_uint64 * 64)()
    meta = (ctypes.c_uint32 * 8)()
    addr, maddr = ctypes.addressof(buf), ctypes.addressof(meta)
    st = _lib.MnkState(9, 9, 5, 3, 4, addr, maddr)             # wrong words
    assert L.mnk_reset(ctypes.byref(st), None, 0, None) == _lib.MNK_ERR_ARG
    st = _lib.MnkState(9, 9, 10, 2, 4, addr, maddr)            # k > n
    assert L.mnk_reset(ctypes.byref(st), None, 0, None) == _lib.MNK_ERR_GEOM
    st = _lib.MnkState(9, 9, 5, 2, 4, addr + 4, maddr)         # misaligned planes
    assert L.mnk_reset(ctypes.byref(st), None, 0, None) == _lib.MNK_ERR_ALIGN
    st = _lib.MnkState(9, 9, 5, 2, 4, addr, maddr)
    assert L.mnk_step(ctypes.byref(st), None, None, 4, None, None, None, None, None, 0, None) == _lib.MNK_ERR_NULL
    assert L.mnk_observe(ctypes.byref(st), None, None, None, 0, None) == _lib.MNK_ERR_NULL
    assert L.mnk_step(ctypes.byref(st), addr, None, 3, addr, addr, None, None, None, 0, None) == _lib.MNK_ERR_ARG
    with pytest.raises(ValueError):
        _lib.check(_lib.MNK_ERR_GEOM, "x")
    assert b"ok" == L.mnk_error_string(0)
    # round-2 entry points: the same argument checks, still without a CUDA call
    a = addr
    assert L.mnk_step_slab(ctypes.byref(st), None, 32, a, 32, 1, None, None, 0, None) == _lib.MNK_ERR_NULL
    assert L.mnk_step_slab(ctypes.byref(st), a, 32, a, 32, 17, None, None, 0, None) == _lib.MNK_ERR_ARG        # > MNK_MAX_SLAB_STEPS
    assert L.mnk_step_slab(ctypes.byref(st), a, 32, a, 16, 2, None, None, 0, None) == _lib.MNK_ERR_ARG         # rd_stride < 5 * num_envs
    assert L.mnk_step_slab(ctypes.byref(st), a, 8, a, 32, 2, None, None, 0, None) == _lib.MNK_ERR_ARG          # action_stride < 8 * num_envs
    assert L.mnk_conv_tower(ctypes.byref(st), None, 80, 11, 1, None, a, a, a, a, a, None, None) == _lib.MNK_ERR_NULL
    assert L.mnk_conv_tower(ctypes.byref(st), None, 72, 11, 1, a, a, a, a, a, a, None, None) == _lib.MNK_ERR_ARG     # unsupported width
    assert L.mnk_conv_tower(ctypes.byref(st), None, 80, 10, 1, a, a, a, a, a, a, None, None) == _lib.MNK_ERR_ARG     # residual needs odd layers
    big = _lib.MnkState(19, 19, 5, 6, 4, addr, maddr)
    assert L.mnk_conv_tower(ctypes.byref(big), None, 96, 8, 0, a, a, a, a, a, a, None, None) == _lib.MNK_ERR_GEOM    # 401 rows > 384
    assert L.mnk_transformer_body(ctypes.byref(st), None, 64, 4, 2, a, a, a, a, a, a, a, a, None, None) == _lib.MNK_ERR_ARG
    assert L.mnk_transformer_body(ctypes.byref(st), None, 56, 4, 0, a, a, a, a, a, a, a, a, None, None) == _lib.MNK_ERR_ARG
    mid = _lib.MnkState(13, 13, 5, 3, 4, addr, maddr)
    assert L.mnk_transformer_body(ctypes.byref(mid), None, 56, 4, 2, a, a, a, a, a, a, a, a, None, None) == _lib.MNK_ERR_GEOM   # 169 tokens
    assert L.mnk_transformer_layer_weight_bytes(56, 4) == 2 * (64 * 192 + 64 * 64 + 64 * 224 + 224 * 64)
    assert L.mnk_transformer_layer_weight_bytes(96, 8) == 2 * (96 * 384 + 128 * 96 + 96 * 384 + 384 * 96)
    assert L.mnk_transformer_layer_weight_bytes(64, 4) == _lib.MNK_ERR_ARG
    assert L.mnk_resnet_tower_train_scratch_bytes(14, 14, 100, 4) == _lib.MNK_ERR_GEOM                               # board rows 3 .. 13
    assert L.mnk_resnet_tower_train_scratch_bytes(13, 13, 100, 4) > 0


def test_no_cpu_fallback():
    import torch
    from mnk_b200 import TorchVectorMnkEnv
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        TorchVectorMnkEnv(3, 3, 3, 4, device="cpu")
    if not torch.cuda.is_available():
        st_words = _lib.lib().mnk_state_words(3, 3)
        assert st_words == 1

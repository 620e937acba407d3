"""ctypes binding of libmnk_b200.so (C ABI declared in include/mnk_b200.h).

The library is built in-tree (``build()``: nvcc, sm_100a only) and loaded from
``rl-selfplay-mnk_b200/libmnk_b200.so``.  There is NO fallback: if the shared object is
missing or a call fails, a RuntimeError / ValueError is raised.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor
from typing import Optional

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # rl-selfplay-mnk_b200/
_CSRC = os.path.join(_PKG_ROOT, "csrc")
_OBJ = os.path.join(_PKG_ROOT, "build")
LIB_PATH = os.path.join(_PKG_ROOT, "libmnk_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "--std=c++17",
    "-Xcompiler", "-fPIC",
]

MNK_OK, MNK_ERR_NULL, MNK_ERR_GEOM, MNK_ERR_ALIGN, MNK_ERR_ARG = 0, -1, -2, -3, -4
STEP_ACTIONS_I32, STEP_AUTORESET, STEP_ZEROCOPY, STEP_PDL, STEP_NOSYNC = 1, 2, 4, 8, 16


class MnkState(ctypes.Structure):
    """struct mnk_state of include/mnk_b200.h."""
    _fields_ = [("m", ctypes.c_int32), ("n", ctypes.c_int32), ("k", ctypes.c_int32), ("words", ctypes.c_int32),
                ("num_envs", ctypes.c_int64), ("bits", ctypes.c_void_p), ("meta", ctypes.c_void_p)]


class MnkSelfplay(ctypes.Structure):
    """struct mnk_selfplay of include/mnk_b200.h."""
    _fields_ = [("agent_side", ctypes.c_void_p), ("pending", ctypes.c_void_p), ("episodes", ctypes.c_void_p),
                ("seed", ctypes.c_uint64), ("env_offset", ctypes.c_int64), ("counter_base", ctypes.c_void_p)]


class MnkHostLoop(ctypes.Structure):
    """struct mnk_host_loop of include/mnk_b200.h."""
    _fields_ = [("host_actions", ctypes.c_void_p), ("host_rd", ctypes.c_void_p), ("host_obs", ctypes.c_void_p),
                ("host_mask", ctypes.c_void_p), ("dev_actions", ctypes.c_void_p), ("dev_rd", ctypes.c_void_p),
                ("obs_ring", ctypes.POINTER(ctypes.c_void_p)), ("mask_ring", ctypes.POINTER(ctypes.c_void_p)),
                ("ring", ctypes.c_int32), ("steps", ctypes.c_int64), ("slab_steps", ctypes.c_int64), ("buffers", ctypes.c_int32)]


class MnkHeadsWeights(ctypes.Structure):
    """struct mnk_heads_weights of include/mnk_b200.h (16 device pointers)."""
    NAMES = ("p_ln1_w", "p_ln1_b", "p_w1t", "p_b1", "p_ln2_w", "p_ln2_b", "p_w2t", "p_b2",
             "v_ln1_w", "v_ln1_b", "v_w1t", "v_b1", "v_ln2_w", "v_ln2_b", "v_w2", "v_b2")
    _fields_ = [(n, ctypes.c_void_p) for n in NAMES]


class MnkBnTrain(ctypes.Structure):
    """mnk_bn_train_t: per-layer BatchNorm parameters / statistics of the train-mode tower (f32 [layers][32] each)."""
    _fields_ = [("gamma", ctypes.c_void_p), ("beta", ctypes.c_void_p), ("conv_bias", ctypes.c_void_p),
                ("running_mean", ctypes.c_void_p), ("running_var", ctypes.c_void_p), ("batch_stats", ctypes.c_void_p),
                ("momentum", ctypes.c_float), ("eps", ctypes.c_float)]


SP_ACTIONS_I32, SP_RESET_ALL, SP_DETERMINISTIC_OPP = 1, 2, 4


def sources():
    return sorted(os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith(".cu"))


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(_CSRC, f) for f in os.listdir(_CSRC)]
    deps.append(os.path.join(os.path.dirname(_PKG_ROOT), "include", "mnk_b200.h"))
    return any(os.path.getmtime(d) > built for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every csrc/*.cu for sm_100a and link libmnk_b200.so (cross-compiles without a GPU)."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(_OBJ, exist_ok=True)

    def compile_one(src: str) -> str:
        obj = os.path.join(_OBJ, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, sources()))
    cmd = [nvcc, *NVCC_FLAGS, "-shared", "-o", LIB_PATH, *objs]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    return LIB_PATH


_lib: Optional[ctypes.CDLL] = None

_VP, _I32, _I64, _U32, _U64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint32, ctypes.c_uint64
_ST = ctypes.POINTER(MnkState)
_SP = ctypes.POINTER(MnkSelfplay)

# name -> (restype, argtypes); must list every symbol include/mnk_b200.h declares
SIGNATURES = {
    "mnk_version": (_I32, []),
    "mnk_error_string": (ctypes.c_char_p, [_I32]),
    "mnk_state_words": (_I32, [_I32, _I32]),
    "mnk_reset": (_I32, [_ST, _VP, _I64, _VP]),
    "mnk_observe": (_I32, [_ST, _VP, _VP, _VP, _I32, _VP]),
    "mnk_step": (_I32, [_ST, _VP, _VP, _I64, _VP, _VP, _VP, _VP, _VP, _U32, _VP]),
    "mnk_step_slab": (_I32, [_ST, _VP, _I64, _VP, _I64, _I32, _VP, _VP, _U32, _VP]),
    "mnk_step_host": (_I32, [_ST, _VP, _VP, _VP, _VP, _VP, _VP, _U32, _VP]),
    "mnk_host_pipe_create": (_I32, [ctypes.POINTER(ctypes.c_void_p)]),
    "mnk_host_pipe_destroy": (_I32, [_VP]),
    "mnk_step_host_loop": (_I32, [_ST, ctypes.POINTER(MnkHostLoop), _VP, _U32, _VP]),
    "mnk_unpack_boards": (_I32, [_ST, _VP, _VP]),
    "mnk_pack_boards": (_I32, [_ST, _VP, _VP]),
    "mnk_export_meta": (_I32, [_ST, _VP, _VP, _VP]),
    "mnk_import_meta": (_I32, [_ST, _VP, _VP, _VP]),
    "mnk_random_legal": (_I32, [_ST, _U64, _U64, _I64, _I32, _VP, _VP]),
    "mnk_masked_sample": (_I32, [_VP, _I64, _VP, _I32, _I64, _U64, _U64, _VP, _I64, _I32, _VP, _VP, _VP, _VP, _VP]),
    "mnk_selfplay_agent": (_I32, [_ST, _SP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _U32, _VP]),
    "mnk_selfplay_opponent": (_I32, [_ST, _SP, _VP, _VP, _VP, _VP, _VP, _VP, _U32, _VP]),
    "mnk_selfplay_step_random": (_I32, [_ST, _SP, _VP, _VP, _U64, _VP, _VP, _VP, _VP, _U32, _VP]),
    "mnk_rollout_store_obs": (_I32, [_ST, _VP, _VP, _VP]),
    "mnk_rollout_gather": (_I32, [_I32, _I32, _I32, _VP, _I64, _VP, _I64, _VP, _VP, _VP]),
    "mnk_gae": (_I32, [_VP, _VP, _VP, _VP, _I64, _I64, ctypes.c_double, ctypes.c_double, _VP, _VP, _VP]),
    "mnk_episode_stats": (_I32, [_VP, _VP, _I64, _VP, _VP, _VP, _VP]),
    "mnk_resnet_operand_dtype": (_I32, []),
    "mnk_resnet_tower": (_I32, [_ST, _VP, _VP, _VP, _VP, _VP, _I32, _VP, _VP, _VP, _VP]),
    "mnk_conv_tower": (_I32, [_ST, _VP, _I32, _I32, _I32, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "mnk_transformer_layer_weight_bytes": (_I64, [_I32, _I32]),
    "mnk_transformer_layer_params": (_I64, [_I32, _I32]),
    "mnk_transformer_body": (_I32, [_ST, _VP, _I32, _I32, _I32, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "mnk_resnet_tower_rows": (_I32, [_ST, _VP, _VP, _VP, _VP, _VP, _I32, _VP, _VP, _VP, _VP]),
    "mnk_resnet_tower_train_scratch_bytes": (_I64, [_I32, _I32, _I64, _I32]),
    "mnk_resnet_tower_train": (_I32, [_ST, _VP, _VP, ctypes.POINTER(MnkBnTrain), _VP, _VP, _I32, _VP, _I64, _VP, _VP, _VP, _VP]),
    "mnk_resnet_heads_mma": (_I32, [_VP, _VP, _I64, _I32, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "mnk_resnet_heads": (_I32, [_VP, _VP, _I64, _I32, ctypes.POINTER(MnkHeadsWeights), _VP, _VP, _VP]),
}


def lib() -> ctypes.CDLL:
    """The loaded library.  Raises if it has not been built -- there is no CPU or eager fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  mnk_b200 has no CPU / eager fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)     # AttributeError => stale library
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(code: int, what: str = "mnk") -> None:
    if code == MNK_OK:
        return
    msg = lib().mnk_error_string(code).decode()
    if code in (MNK_ERR_GEOM, MNK_ERR_ARG, MNK_ERR_ALIGN, MNK_ERR_NULL):
        raise ValueError(f"{what}: {msg} (code {code})")
    raise RuntimeError(f"{what}: CUDA error {code}: {msg}")

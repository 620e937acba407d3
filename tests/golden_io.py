"""Helpers to read tests/golden/*.npz (written by oracle/gen_golden.py from the reference)."""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
OP_STEP_SUBSET, OP_RESET_IDX, OP_RESET_ALL, OP_STEP = 0, 1, 2, 3


def files(prefix):
    return sorted(glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def load(path):
    return dict(np.load(path))


def unpack(bits, shape):
    """Inverse of gen_golden.pack: bits u8[B, ceil(prod(shape)/8)] -> bool[B, *shape]."""
    count = int(np.prod(shape))
    return np.unpackbits(bits, axis=1)[:, :count].reshape((bits.shape[0],) + tuple(shape)).astype(bool)


def name(path):
    return os.path.splitext(os.path.basename(path))[0]

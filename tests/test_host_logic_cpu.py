"""CPU tests of host-side logic that needs no GPU: OpponentPool (reference: src/selfplay/opponent_pool.py:5-19
as used by src/train.py:96-123), default sampler seeds, the strict-mode switch."""
import collections
import random

import pytest

from mnk_b200.sampling import fresh_seed
from selfplay.opponent_pool import OpponentPool


def test_opponent_pool_fifo_eviction_and_uniform_choice():
    pool = OpponentPool(max_size=3)
    assert pool.size() == 0 and pool.get_random_opponent() is None          # empty pool: None, like the reference
    for name in "abcd":
        pool.add_opponent(name)
    assert pool.size() == 3 and list(pool.pool) == ["b", "c", "d"]          # oldest entry evicted first (deque maxlen)
    random.seed(0)
    seen = collections.Counter(pool.get_random_opponent() for _ in range(3000))
    assert set(seen) == {"b", "c", "d"} and min(seen.values()) > 850        # uniform over the survivors
    assert OpponentPool().max_size == 5                                     # the reference's default


def test_opponent_pool_matches_live_reference():
    from oracle import ref_tree
    if not ref_tree.available():
        pytest.skip("reference tree neither mounted nor staged (oracle/_ref)")
    ref = ref_tree.load("selfplay.opponent_pool")
    ours, theirs = OpponentPool(max_size=4), ref.OpponentPool(max_size=4)
    for i in range(9):
        ours.add_opponent(i), theirs.add_opponent(i)
        random.seed(i)
        a = [ours.get_random_opponent() for _ in range(20)]
        random.seed(i)
        b = [theirs.get_random_opponent() for _ in range(20)]
        assert a == b and ours.size() == theirs.size() and list(ours.pool) == list(theirs.pool)


def test_default_seeds_are_distinct_per_object():
    """Two default-constructed samplers must not share a Philox key (agent and opponent drew the same
    Gumbel noise on consecutive plies when every default seed was 0)."""
    from mnk_b200 import NNPolicy, RandomPolicy
    import torch
    seeds = [fresh_seed() for _ in range(1000)]
    assert len(set(seeds)) == 1000 and all(0 <= s < 2**64 for s in seeds)
    a, b = RandomPolicy(9), RandomPolicy(9)
    assert a.seed != b.seed
    assert NNPolicy(torch.nn.Identity()).seed != NNPolicy(torch.nn.Identity()).seed
    assert RandomPolicy(9, seed=5).seed == 5                                # explicit seeds are kept


def test_env_refuses_cpu_and_missing_library_fails_loudly(monkeypatch):
    from mnk_b200 import TorchVectorMnkEnv, _lib
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        TorchVectorMnkEnv(3, 3, 3, 4, device="cpu")
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libmnk_b200.so")
    with pytest.raises(RuntimeError, match="no CPU / eager fallback"):
        _lib.lib()

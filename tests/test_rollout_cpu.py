"""CPU tests of the rollout host logic: the oracle's GAE against the reference's RolloutBuffer
(golden), shard arithmetic, and the N>1 statistics all-reduce over gloo with world_size 2."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import golden_io as gio
from oracle import mnk_oracle as orc
from mnk_b200 import dist as mdist


def test_oracle_gae_matches_reference_buffer():
    g = gio.load(gio.files("rollout_buffer_gae")[0])
    adv, ret = orc.gae(g["rewards"], g["values"], g["dones"], g["last_values"], round(float(g["gamma"]), 6), round(float(g["lam"]), 6))
    assert np.array_equal(adv, g["advantages"]) and np.array_equal(ret, g["returns"])
    assert np.array_equal(g["stored_obs"], g["obs"]) and np.array_equal(g["stored_masks"], g["masks"])


GAE_PAIRS = [(0.99, 0.95), (0.997, 0.9), (0.9, 0.97), (0.993, 0.913), (1.0, 1.0), (0.95, 0.0)]


def test_gae_pairs_include_one_where_float_product_differs():
    """gamma * gae_lambda is multiplied in double and THEN rounded to fp32 by the reference (rollout_buffer.py:76);
    multiplying the fp32 roundings instead is off by one ulp for some pairs -- make sure the list has such a pair."""
    f = np.float32
    assert any(f(f(g) * f(l)) != f(g * l) for g, l in GAE_PAIRS)


@pytest.mark.parametrize("gamma,lam", GAE_PAIRS)
def test_oracle_gae_matches_live_reference_buffer(gamma, lam):
    from oracle import ref_tree
    if not ref_tree.available():
        pytest.skip("reference tree neither mounted nor staged (oracle/_ref)")
    rb = ref_tree.load("alg.rollout_buffer")
    rng = np.random.default_rng(11)
    steps, ne = 23, 257
    buf = rb.RolloutBuffer(steps, ne, (2, 3, 3), 9, device="cpu")
    rewards = rng.choice([-1.0, 0.0, 1.0], size=(steps, ne)).astype(np.float32)
    values = rng.normal(size=(steps, ne)).astype(np.float32)
    dones = rng.random((steps, ne)) < 0.1
    last = rng.normal(size=ne).astype(np.float32)
    buf.rewards.copy_(torch.from_numpy(rewards)), buf.values.copy_(torch.from_numpy(values)), buf.dones.copy_(torch.from_numpy(dones))
    buf.ptr = steps
    buf.compute_advantages_and_returns(torch.from_numpy(last), gamma, lam)
    adv, ret = orc.gae(rewards, values, dones, last, gamma, lam)
    assert np.array_equal(adv, buf.advantages.numpy()) and np.array_equal(ret, buf.returns.numpy())


def test_shard_partition():
    for total, world in [(65536, 8), (1000, 3), (7, 8), (4194304, 8), (5, 1)]:
        spans = [mdist.shard(total, world, r) for r in range(world)]
        assert sum(c for c, _ in spans) == total
        pos = 0
        for c, off in spans:
            assert off == pos and c >= 0
            pos += c
        assert max(c for c, _ in spans) - min(c for c, _ in spans) <= 1
    with pytest.raises(ValueError):
        mdist.shard(10, 2, 2)


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, w, _ = mdist.init_from_env(backend="gloo")
    count, offset = mdist.shard(1001, w, r)
    # each rank accounts for its own shard: {episodes, reward sum, length sum, wins, losses, draws}
    ids = torch.arange(offset, offset + count, dtype=torch.float64)
    stats = torch.stack([torch.tensor(float(count), dtype=torch.float64), ids.sum(), (ids * 2).sum(),
                         (ids % 3 == 0).double().sum(), (ids % 3 == 1).double().sum(), (ids % 3 == 2).double().sum()])
    mdist.reduce_stats(stats)
    t_ms = torch.tensor([10.0 + rank], dtype=torch.float64)
    mdist.reduce_stats(t_ms, op=dist.ReduceOp.MAX)          # bench.py takes the max time over ranks
    # learner collectives: gradient averaging and global advantage moments
    torch.manual_seed(0)
    lin = torch.nn.Linear(4, 3)
    lin.weight.grad = torch.full_like(lin.weight, float(rank + 1))
    lin.bias.grad = torch.arange(3, dtype=torch.float32) * (rank + 1)
    mdist.average_gradients(lin.parameters(), w)
    shard_vals = torch.arange(offset, offset + count, dtype=torch.float32) * 0.01
    mean, std = mdist.global_mean_std(shard_vals)
    out[rank] = (stats.tolist(), t_ms.item(), lin.weight.grad[0, 0].item(), lin.bias.grad.tolist(), mean.item(), std.item())
    dist.destroy_process_group()


def test_two_rank_gloo_stats_allreduce():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        ids = np.arange(1001, dtype=np.float64)
        want = [1001.0, ids.sum(), 2 * ids.sum(), (ids % 3 == 0).sum(), (ids % 3 == 1).sum(), (ids % 3 == 2).sum()]
        for rank in range(world):
            stats, tmax, gw, gb, mean, std = out[rank]
            assert stats == want and tmax == 11.0
            assert gw == 1.5 and gb == [0.0, 1.5, 3.0]                      # mean of the two ranks' gradients
            full = np.arange(1001, dtype=np.float64) * 0.01
            assert abs(mean - full.mean()) < 1e-5 and abs(std - full.std(ddof=1)) < 1e-5

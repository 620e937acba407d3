"""GPU parity tests of the env kernels: golden traces recorded from the reference, the CPU
oracle on larger seeded batches, and size-independent properties at BASELINE sizes.
All comparisons are bit-exact (torch.equal / np.array_equal)."""
import numpy as np
import pytest
import torch

import golden_io as gio
from oracle import mnk_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda"


def make_env(*a, **kw):
    from env.torch_vector_mnk_env import TorchVectorMnkEnv
    return TorchVectorMnkEnv(*a, device=DEV, **kw)


def t(x, dtype=None):
    return torch.as_tensor(np.ascontiguousarray(x), device=DEV) if dtype is None else torch.as_tensor(np.ascontiguousarray(x), device=DEV, dtype=dtype)


@pytest.mark.parametrize("mirrors", [False, True], ids=["packed", "mirrors"])
@pytest.mark.parametrize("path", gio.files("env_trace_"), ids=gio.name)
def test_golden_env_trace(path, mirrors):
    g = gio.load(path)
    m, n, k, ne = (int(x) for x in g["geom"])
    env = make_env(m, n, k, ne)
    env.reset()
    if mirrors:
        held = (env.boards, env.current_player, env.move_counts)     # live mirrors stay current
    for step in range(len(g["op"])):
        op, active, actions = int(g["op"][step]), g["active"][step], g["actions"][step]
        idx = np.nonzero(active)[0]
        rewards = torch.zeros(ne, device=DEV)
        dones = torch.zeros(ne, dtype=torch.bool, device=DEV)
        if op == gio.OP_RESET_ALL:
            obs = env.reset()
        elif op == gio.OP_RESET_IDX:
            obs = env.reset(t(idx))
        elif op == gio.OP_STEP:
            a = t(actions)
            obs, rewards, dones = env.step(a.int() if step % 3 == 0 else a)     # int32 actions are accepted
        else:
            obs, rewards, dones = env.step_subset(t(actions[idx]), t(idx))
        want_boards = gio.unpack(g["boards"][step], (2, m, n))
        assert obs["observation"].dtype == torch.float32 and obs["action_mask"].dtype == torch.bool
        assert rewards.dtype == torch.float32 and dones.dtype == torch.bool
        assert np.array_equal(obs["observation"].cpu().numpy(), want_boards.astype(np.float32)), step
        assert np.array_equal(obs["action_mask"].cpu().numpy(), gio.unpack(g["mask"][step], (m * n,))), step
        assert np.array_equal(rewards.cpu().numpy(), g["rewards"][step]), step
        assert np.array_equal(dones.cpu().numpy(), g["dones"][step]), step
        if mirrors:
            assert np.array_equal(held[0].cpu().numpy(), want_boards.astype(np.float32)), step
            assert np.array_equal(held[1].cpu().numpy(), g["player"][step]), step
            assert np.array_equal(held[2].cpu().numpy(), g["count"][step]), step
    fresh = make_env(m, n, k, ne)     # state read back only at the end in packed mode
    assert np.array_equal(env.current_player.cpu().numpy(), g["player"][-1])
    assert np.array_equal(env.move_counts.cpu().numpy(), g["count"][-1])
    assert fresh.boards.sum().item() == 0


@pytest.mark.parametrize("path", gio.files("env_poke_"), ids=gio.name)
def test_golden_env_poke(path):
    """Positions written straight into env.boards / current_player / move_counts (as the reference's
    tests do) and one step from there: every line direction, overlines, wraps, draws."""
    g = gio.load(path)
    m, n, k, ne = (int(x) for x in g["geom"])
    env = make_env(m, n, k, ne)
    env.reset()
    env.boards[:] = t(gio.unpack(g["init_boards"], (2, m, n)).astype(np.float32))
    env.current_player[:] = t(g["init_player"])
    env.move_counts[:] = t(g["init_count"])
    obs, rewards, dones = env.step(t(g["actions"]))
    assert np.array_equal(env.boards.cpu().numpy().astype(bool), gio.unpack(g["boards"], (2, m, n)))
    assert np.array_equal(obs["action_mask"].cpu().numpy(), gio.unpack(g["mask"], (m * n,)))
    assert np.array_equal(env.current_player.cpu().numpy(), g["player"])
    assert np.array_equal(env.move_counts.cpu().numpy(), g["count"])
    assert np.array_equal(rewards.cpu().numpy(), g["rewards"])
    assert np.array_equal(dones.cpu().numpy(), g["dones"])


GEOMS = [(3, 3, 3, 1000), (9, 9, 5, 4099), (13, 13, 5, 1031), (19, 19, 5, 517), (15, 15, 5, 300),
         (7, 11, 4, 777), (6, 7, 4, 640), (4, 31, 4, 65), (16, 31, 5, 33), (5, 5, 1, 64), (10, 10, 10, 129)]


@pytest.mark.parametrize("m,n,k,ne", GEOMS, ids=lambda v: str(v))
def test_oracle_parity_random_play(m, n, k, ne):
    """Seeded random legal play with finished envs reset: CUDA env vs CPU oracle step by step, with
    the actions drawn by mnk_random_legal and checked against the oracle's Philox contract."""
    env = make_env(m, n, k, ne, env_offset=12345)
    ref = orc.OracleEnv(m, n, k, ne)
    obs = env.reset()
    ref.reset()
    steps = min(3 * m * n // 2 + 5, 220)
    for step in range(steps):
        a = env.random_legal_actions(seed=99, counter=step)
        want = orc.random_legal_actions(obs["action_mask"].cpu().numpy(), 99, step, env_offset=12345)
        assert np.array_equal(a.cpu().numpy(), want), step
        if step % 2 == 0:
            obs, r, d = env.step(a)
            o2, r2, d2 = ref.step(want)
            done_idx = np.nonzero(d2)[0]
            if len(done_idx):
                obs = env.reset(t(done_idx))
                o2 = ref.reset(done_idx)
        else:   # fused step + reset of finished envs; returned obs is the pre-reset one
            obs, r, d = env.step_autoreset(a)
            o2, r2, d2 = ref.step(want)
            assert np.array_equal(obs["observation"].cpu().numpy(), o2["observation"]), step
            done_idx = np.nonzero(d2)[0]
            if len(done_idx):
                o2 = ref.reset(done_idx)
            obs = env.observe()
        assert np.array_equal(r.cpu().numpy(), r2) and np.array_equal(d.cpu().numpy(), d2), step
        assert np.array_equal(obs["observation"].cpu().numpy(), o2["observation"]), step
        assert np.array_equal(obs["action_mask"].cpu().numpy(), o2["action_mask"]), step
    assert np.array_equal(env.current_player.cpu().numpy(), ref.current_player)
    assert np.array_equal(env.move_counts.cpu().numpy(), ref.move_counts)
    assert np.array_equal(env.boards.cpu().numpy(), ref.boards.astype(np.float32))


def test_baseline_size_properties_and_oracle():
    """BASELINE cfg 2 size (9x9x5, 65,536 envs): oracle parity on a short run plus invariants that
    hold for any number of envs under legal play."""
    m, n, k, ne = 9, 9, 5, 65536
    env = make_env(m, n, k, ne)
    ref = orc.OracleEnv(m, n, k, ne)
    env.reset()
    ref.reset()
    total_done = 0
    for step in range(70):
        a = env.random_legal_actions(seed=3, counter=step)
        obs, r, d = env.step_autoreset(a)
        _, r2, d2 = ref.step(a.cpu().numpy())
        assert np.array_equal(r.cpu().numpy(), r2) and np.array_equal(d.cpu().numpy(), d2), step
        o = obs["observation"]
        stones = o.sum(dim=(1, 2, 3))
        assert torch.equal(obs["action_mask"], ~(o[:, 0].bool() | o[:, 1].bool()).flatten(1))
        assert bool((r <= d.float()).all())                 # a reward implies done
        assert bool((o[:, 0] * o[:, 1]).sum() == 0)         # legal play never double-occupies
        done_idx = np.nonzero(d2)[0]
        total_done += len(done_idx)
        if len(done_idx):
            ref.reset(done_idx)
        assert torch.equal(env.move_counts, torch.where(d, torch.zeros_like(stones), stones).long()), step
    assert total_done > 1000
    assert np.array_equal(env.boards.cpu().numpy(), ref.boards.astype(np.float32))


def test_tail_tiles_and_empty_cases():
    """num_envs not a multiple of the 32-env warp tile, single env, empty index lists."""
    for ne in (1, 2, 31, 33, 63, 97):
        env = make_env(9, 9, 5, ne)
        ref = orc.OracleEnv(9, 9, 5, ne)
        env.reset(), ref.reset()
        for step in range(12):
            a = env.random_legal_actions(1, step)
            obs, r, d = env.step(a)
            o2, r2, d2 = ref.step(a.cpu().numpy())
            assert np.array_equal(obs["observation"].cpu().numpy(), o2["observation"])
            assert np.array_equal(obs["action_mask"].cpu().numpy(), o2["action_mask"])
    env = make_env(3, 3, 3, 8)
    env.reset()
    obs, r, d = env.step_subset(torch.empty(0, dtype=torch.long, device=DEV), torch.empty(0, dtype=torch.long, device=DEV))
    assert r.sum().item() == 0 and not d.any() and obs["action_mask"].all()     # the reference raises here; we no-op
    obs = env.reset(torch.empty(0, dtype=torch.long, device=DEV))
    assert obs["observation"].shape == (8, 2, 3, 3)


def test_reference_env_tests_port():
    """The two env-level reference tests (src/tests/test_mnk_integration.py:50-81)."""
    env = make_env(3, 3, 3, 1)
    env.reset()
    env.boards[0, 0, 0, 0] = 1
    env.boards[0, 0, 0, 1] = 1
    _, rewards, dones = env.step(torch.tensor([2], device=DEV))
    assert dones[0].item() is True
    assert rewards[0].item() == 1.0
    # test_env_illegal_move: fails on the reference itself (validators are dead code, :86-104);
    # the drop-in reproduces the silent behaviour by default and raises in strict mode.
    env = make_env(3, 3, 3, 1)
    env.reset()
    env.boards[0, 0, 0, 0] = 1
    obs, r, d = env.step(torch.tensor([0], device=DEV))
    assert env.move_counts[0].item() == 1 and env.current_player[0].item() == 1 and not d[0].item()
    strict = make_env(3, 3, 3, 1, strict=True)
    strict.reset()
    strict.boards[0, 0, 0, 0] = 1
    with pytest.raises(ValueError, match="Illegal Move"):
        strict.step(torch.tensor([0], device=DEV))
    with pytest.raises(ValueError, match="Action out of bounds"):
        strict.step(torch.tensor([9], device=DEV))


def test_observe_returns_fresh_tensors_and_swap_fix():
    env = make_env(3, 3, 3, 4)
    env.reset()
    a = env.observe()
    b = env.observe()
    assert a["observation"].data_ptr() != b["observation"].data_ptr()
    a["observation"].fill_(7)         # callers mutate what observe() returns (wrapper :89,:106,:110)
    assert env.observe()["observation"].sum().item() == 0
    env.step(torch.tensor([0, 1, 2, 3], device=DEV))
    swap = torch.tensor([0, 1, 0, 1], dtype=torch.uint8, device=DEV)
    o = env._observe_packed(swap=swap)["observation"]
    raw = env.observe()["observation"]
    assert torch.equal(o[0], raw[0]) and torch.equal(o[1], raw[1].flip(0)) and torch.equal(o[3], raw[3].flip(0))


def test_host_step_path():
    env = make_env(9, 9, 5, 1000)
    twin = make_env(9, 9, 5, 1000)
    env.reset(), twin.reset()
    host_a = torch.empty(1000, dtype=torch.long).pin_memory()
    host_out = torch.empty(5000, dtype=torch.uint8).pin_memory()
    for step in range(30):
        a = twin.random_legal_actions(5, step)
        host_a.copy_(a)
        obs, r, d = env.step_host(host_a, host_out, autoreset=True, zero_copy=(step % 2 == 1))
        o2, r2, d2 = twin.step_autoreset(a)
        assert not r.is_cuda and torch.equal(r, r2.cpu()) and torch.equal(d, d2.cpu())
        assert torch.equal(obs["observation"], o2["observation"])
    assert env.state_checksum() == twin.state_checksum()


def test_host_step_two_groups_in_flight():
    """step_host(sync=False): two env groups on two streams, each synchronised only before its results are read;
    with env_offset the two halves reproduce one env of twice the size step for step."""
    half, steps = 640, 25
    whole = make_env(9, 9, 5, 2 * half)
    whole.reset()
    groups = [make_env(9, 9, 5, half, env_offset=g * half) for g in range(2)]
    streams = [torch.cuda.Stream() for _ in range(2)]
    host_a = [torch.empty((steps, half), dtype=torch.long).pin_memory() for _ in range(2)]
    host_out = [[torch.empty(5 * half, dtype=torch.uint8).pin_memory() for _ in range(2)] for _ in range(2)]
    want = []
    for t in range(steps):                        # the action trace and the expected results, on the single env
        a = whole.random_legal_actions(9, t)
        for g in range(2):
            host_a[g][t].copy_(a[g * half:(g + 1) * half])
        _, r, d = whole.step_autoreset(a, materialise=False)
        want.append((r.cpu(), d.cpu()))
    torch.cuda.synchronize()
    for g in range(2):
        with torch.cuda.stream(streams[g]):
            groups[g].reset()
    pending = [None, None]
    for t in range(steps + 1):
        for g in range(2):
            if pending[g] is not None:            # results of this group's previous step
                streams[g].synchronize()
                tt, r, d = pending[g]
                assert torch.equal(r, want[tt][0][g * half:(g + 1) * half]) and torch.equal(d, want[tt][1][g * half:(g + 1) * half])
                pending[g] = None
            if t < steps:
                with torch.cuda.stream(streams[g]):
                    _, r, d = groups[g].step_host(host_a[g][t], host_out[g][t % 2], autoreset=True, zero_copy=(g == 0),
                                                  sync=False, out=(None, None))
                pending[g] = (t, r, d)
    torch.cuda.synchronize()
    bits = whole._bits.view(2, -1, 2 * half)
    for g in range(2):
        assert torch.equal(groups[g]._bits.view(2, -1, half), bits[:, :, g * half:(g + 1) * half])


@pytest.mark.parametrize("slab,views,dtype", [(1, False, torch.long), (3, True, torch.long), (4, False, torch.int32),
                                              (40, False, torch.long), (5, True, torch.int32)], ids=lambda v: str(v))
def test_host_step_loop_matches_single_steps(slab, views, dtype):
    """mnk_step_host_loop (K steps per C call, slab-pipelined copies) against K device-side single steps of a twin env:
    rewards / dones bit-identical for every step, observation ring and host copies of the views identical, final
    state identical -- for slab sizes that divide K, do not divide it, and exceed it."""
    ne, K = 1500, 17
    env, twin = make_env(9, 9, 5, ne), make_env(9, 9, 5, ne)
    env.reset(), twin.reset()
    for w in range(30):                                   # mid-game positions with finished games around
        a = twin.random_legal_actions(3, w)
        twin.step_autoreset(a, materialise=False), env.step_autoreset(a, materialise=False)
    host_a = torch.empty((K, ne), dtype=dtype).pin_memory()
    want = []
    for s_ in range(K):
        a = twin.random_legal_actions(4, s_)
        host_a[s_].copy_(a)
        o, r, d = twin.step_autoreset(a)
        want.append((o["observation"].cpu(), o["action_mask"].cpu(), r.cpu(), d.cpu()))
    host_out = torch.empty((K, 5 * ne), dtype=torch.uint8).pin_memory()
    ring_n = max(2 * min(slab, K), 3)
    ring = ([torch.empty((ne, 2, 9, 9), device=DEV) for _ in range(ring_n)],
            [torch.empty((ne, 81), dtype=torch.bool, device=DEV) for _ in range(ring_n)])
    host_obs = torch.empty((K, ne, 2, 9, 9)).pin_memory() if views else None
    host_mask = torch.empty((K, ne, 81), dtype=torch.bool).pin_memory() if views else None
    r_all, d_all = env.step_host_loop(host_a, host_out, slab_steps=slab, autoreset=True, ring=ring, host_obs=host_obs,
                                      host_mask=host_mask)
    assert r_all.shape == (K, ne) and d_all.shape == (K, ne) and not r_all.is_cuda
    for s_ in range(K):
        assert torch.equal(r_all[s_], want[s_][2]) and torch.equal(d_all[s_], want[s_][3]), s_
        if views:
            assert torch.equal(host_obs[s_], want[s_][0]) and torch.equal(host_mask[s_], want[s_][1]), s_
    last = (K - 1) % ring_n
    assert torch.equal(ring[0][last].cpu(), want[-1][0]) and torch.equal(ring[1][last].cpu(), want[-1][1])
    assert env.state_checksum() == twin.state_checksum()
    # packed mode (no views) continues from the same state
    a = twin.random_legal_actions(5, 0)
    host_a[0].copy_(a)
    _, r, d = twin.step_autoreset(a, materialise=False)
    r1, d1 = env.step_host_loop(host_a[:1], host_out, slab_steps=slab, autoreset=True)
    assert torch.equal(r1[0], r.cpu()) and torch.equal(d1[0], d.cpu()) and env.state_checksum() == twin.state_checksum()


def test_tournament_style_consumer_on_raw_env():
    """Second consumer of the env in the reference: MatchRunner._play_batch_games
    (src/model_comparison/match_runner.py:125-218) drives the RAW env -- reads env.current_player every
    ply, builds per-side observation subsets, steps only unfinished games with step_subset.  Same loop
    here on the drop-in and, in lock-step, on the CPU oracle with identical actions."""
    m, n, k, games = 6, 7, 4, 300
    env = make_env(m, n, k, games)
    ref = orc.OracleEnv(m, n, k, games)
    obs = env.reset()
    ref.reset()
    dones = torch.zeros(games, dtype=torch.bool, device=DEV)
    wins = losses = draws = 0
    p1_side = 0
    rng = np.random.default_rng(3)
    plies = 0
    while not dones.all():
        current_player = env.current_player                       # live mirror, as the reference reads it
        active = ~dones
        is_p1 = (current_player == p1_side) & active
        moving = torch.nonzero(active).squeeze(1)
        mask = obs["action_mask"]
        # both "policies": a random legal cell (p1 flips the observation when white, as the reference does)
        _ = torch.flip(obs["observation"][is_p1], dims=(1,)) if p1_side == 1 else obs["observation"][is_p1]
        u = torch.from_numpy(rng.random(games)).to(DEV)
        cnt = mask.sum(1)
        j = (u * cnt).long().clamp(max=(cnt - 1).clamp(min=0))
        actions = torch.argmax((mask & ((torch.cumsum(mask.long(), 1) - 1) == j[:, None])).long(), dim=1)
        obs, rewards, step_dones = env.step_subset(actions[moving], moving)
        o2, r2, d2 = ref.step_subset(actions[moving].cpu().numpy(), moving.cpu().numpy())
        assert np.array_equal(rewards.cpu().numpy(), r2) and np.array_equal(step_dones.cpu().numpy(), d2)
        assert np.array_equal(obs["observation"].cpu().numpy(), o2["observation"])
        just = step_dones & ~dones
        winners = (rewards == 1.0) & just
        wins += int((winners & is_p1).sum())
        losses += int((winners & ~is_p1).sum())
        draws += int(((rewards == 0.0) & just).sum())
        dones |= just
        plies += 1
        assert plies <= m * n
    assert wins + losses + draws == games and wins > 0 and losses > 0
    assert np.array_equal(env.current_player.cpu().numpy(), ref.current_player)


def test_strict_counters_of_the_c_abi():
    """mnk_step's `illegal` output (include/mnk_b200.h): [0] counts moves onto occupied / out-of-range cells,
    [1] encodes the smallest offending env; the moves are still applied, as in the reference."""
    import ctypes
    from mnk_b200 import _lib
    env = make_env(9, 9, 5, 200)
    env.reset()
    a = torch.zeros(200, dtype=torch.long, device=DEV)
    env.step(a)                                            # cell 0 now occupied everywhere
    a = torch.full((200,), 5, dtype=torch.long, device=DEV)
    a[17] = 0                                              # occupied
    a[150] = 0
    a[60] = 81                                             # out of range
    illegal = torch.zeros(2, dtype=torch.int32, device=DEV)
    rewards = torch.empty(200, device=DEV)
    dones = torch.empty(200, dtype=torch.bool, device=DEV)
    rc = _lib.lib().mnk_step(env._stp, a.data_ptr(), None, 200, rewards.data_ptr(), dones.data_ptr(), None, None,
                             illegal.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    cnt, enc = illegal.tolist()
    assert cnt == 3 and 0x7FFFFFFF - enc == 17
    assert env.move_counts.eq(2).all()                     # every move counted, illegal ones included
    idx = torch.tensor([3, 60], device=DEV)
    illegal.zero_()
    rc = _lib.lib().mnk_step(env._stp, torch.tensor([5, 1], device=DEV).data_ptr(), idx.data_ptr(), 2, rewards.data_ptr(),
                             dones.data_ptr(), None, None, illegal.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
    assert rc == 0 and illegal.tolist()[0] == 1 and 0x7FFFFFFF - illegal.tolist()[1] == 3


@pytest.mark.parametrize("m,n,k,ne", [(9, 9, 5, 1000), (3, 3, 3, 77), (13, 13, 5, 65), (7, 11, 4, 50), (19, 19, 5, 33)], ids=str)
def test_step_slab_equals_consecutive_steps(m, n, k, ne):
    """mnk_step_slab (K dense steps in one launch, the kernel looping over a tile's steps) against K calls of mnk_step on a
    twin env: bit-identical state, rewards, dones and every step's materialised observation / mask -- static and dynamic
    geometries, tail tiles, auto-reset, with and without materialisation, int32 actions."""
    import ctypes
    from mnk_b200 import TorchVectorMnkEnv, _lib
    L = _lib.lib()
    K = 11
    for materialise, act32 in ((True, False), (False, True)):
        a_env = TorchVectorMnkEnv(m, n, k, ne, device=DEV)
        b_env = TorchVectorMnkEnv(m, n, k, ne, device=DEV)
        a_env.reset(), b_env.reset()
        for t in range(m * n // 3):                          # a mid-game start, identical in both
            act = a_env.random_legal_actions(3, t)
            a_env.step_autoreset(act, materialise=False), b_env.step_autoreset(act, materialise=False)
        # K action batches generated by playing a third twin forward (legal at every step)
        probe = TorchVectorMnkEnv(m, n, k, ne, device=DEV)
        probe._bits.copy_(a_env._bits), probe._meta.copy_(a_env._meta)
        acts = []
        for t in range(K):
            act = probe.random_legal_actions(9, t)
            acts.append(act)
            probe.step_autoreset(act, materialise=False)
        acts = torch.stack(acts).to(torch.int32 if act32 else torch.int64).contiguous()
        want = []
        for t in range(K):
            obs, r, d = a_env.step_autoreset(acts[t].long(), materialise=materialise)
            want.append((r.clone(), d.clone(), None if not materialise else obs["observation"].clone(),
                         None if not materialise else obs["action_mask"].clone()))
        rd_stride = (5 * ne + 3) // 4 * 4 + 8
        rd = torch.zeros(K * rd_stride, dtype=torch.uint8, device=DEV)
        obs_out = [torch.empty((ne, 2, m, n), dtype=torch.float32, device=DEV) for _ in range(K)]
        mask_out = [torch.empty((ne, m * n), dtype=torch.bool, device=DEV) for _ in range(K)]
        PtrArr = ctypes.c_void_p * K
        flags = _lib.STEP_AUTORESET | (_lib.STEP_ACTIONS_I32 if act32 else 0)
        b_env._fold_mirrors()
        _lib.check(L.mnk_step_slab(ctypes.byref(b_env._st), acts.data_ptr(), acts.stride(0) * acts.element_size(), rd.data_ptr(),
                                   rd_stride, K, PtrArr(*[o.data_ptr() for o in obs_out]) if materialise else None,
                                   PtrArr(*[o.data_ptr() for o in mask_out]) if materialise else None, flags,
                                   torch.cuda.current_stream().cuda_stream), "mnk_step_slab")
        torch.cuda.synchronize()
        assert torch.equal(a_env._bits, b_env._bits) and torch.equal(a_env._meta, b_env._meta)
        for t in range(K):
            block = rd[t * rd_stride:(t + 1) * rd_stride]
            assert torch.equal(block[:4 * ne].view(torch.float32), want[t][0]), t
            assert torch.equal(block[4 * ne:5 * ne].bool(), want[t][1]), t
            if materialise:
                assert torch.equal(obs_out[t], want[t][2]) and torch.equal(mask_out[t], want[t][3]), t
    with pytest.raises(ValueError):
        _lib.check(L.mnk_step_slab(ctypes.byref(b_env._st), acts.data_ptr(), 8 * ne, rd.data_ptr(), rd_stride, 17, None, None, 0,
                                   torch.cuda.current_stream().cuda_stream), "mnk_step_slab")


def test_fuzzed_geometries_against_the_oracle():
    """Forty random (m, n, k, envs) -- every word count 1..8, n up to 32, k from 1 to min(m, n), env counts around the 32-env
    tile -- through the dynamic-geometry kernels, a short game each, bit-exact against the C oracle (planes, mask, rewards,
    dones, counters).  The parametrised test above fixes eleven geometries; this one widens the net."""
    rng = np.random.default_rng(20261019)
    seen_words = set()
    for trial in range(40):
        while True:
            m, n = int(rng.integers(1, 17)), int(rng.integers(1, 33))
            if m * (n + 1) <= 512:
                break
        k = int(rng.integers(1, min(m, n) + 1))
        ne = int(rng.choice([1, 31, 32, 33, 63, 65, 100, 257]))
        env = make_env(m, n, k, ne)
        seen_words.add(env.words if hasattr(env, "words") else -1)
        ref = orc.OracleEnv(m, n, k, ne)
        obs = env.reset()
        ref.reset()
        for step in range(min(m * n + 2, 60)):
            a = env.random_legal_actions(seed=trial, counter=step)
            obs, r, d = env.step_autoreset(a)
            o2, r2, d2 = ref.step(a.cpu().numpy())
            assert np.array_equal(r.cpu().numpy(), r2) and np.array_equal(d.cpu().numpy(), d2), (m, n, k, ne, step)
            assert np.array_equal(obs["observation"].cpu().numpy(), o2["observation"]), (m, n, k, ne, step)
            assert np.array_equal(obs["action_mask"].cpu().numpy(), o2["action_mask"]), (m, n, k, ne, step)
            done_idx = np.nonzero(d2)[0]
            if len(done_idx):
                ref.reset(done_idx)
        assert np.array_equal(env.boards.cpu().numpy(), ref.boards.astype(np.float32)), (m, n, k, ne)
        assert np.array_equal(env.move_counts.cpu().numpy(), ref.move_counts), (m, n, k, ne)

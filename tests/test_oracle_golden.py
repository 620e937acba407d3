"""Pins the CPU oracle (oracle/mnk_oracle.{c,py}) to the golden vectors recorded from the
unmodified reference (oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest

import golden_io as gio
from oracle import mnk_oracle as orc


@pytest.mark.parametrize("path", gio.files("env_trace_"), ids=gio.name)
def test_env_trace(path):
    g = gio.load(path)
    m, n, k, ne = (int(x) for x in g["geom"])
    env = orc.OracleEnv(m, n, k, ne)
    env.reset()
    for t in range(len(g["op"])):
        op, active, actions = int(g["op"][t]), g["active"][t], g["actions"][t]
        idx = np.nonzero(active)[0]
        rewards = np.zeros(ne, dtype=np.float32)
        dones = np.zeros(ne, dtype=bool)
        if op == gio.OP_RESET_ALL:
            obs = env.reset()
        elif op == gio.OP_RESET_IDX:
            obs = env.reset(idx)
        elif op == gio.OP_STEP:
            obs, rewards, dones = env.step(actions)
        else:
            obs, rewards, dones = env.step_subset(actions[idx], idx)
        assert np.array_equal(env.boards.astype(bool), gio.unpack(g["boards"][t], (2, m, n))), t
        assert np.array_equal(obs["observation"], env.boards.astype(np.float32))
        assert np.array_equal(obs["action_mask"], gio.unpack(g["mask"][t], (m * n,))), t
        assert np.array_equal(env.current_player, g["player"][t]), t
        assert np.array_equal(env.move_counts, g["count"][t]), t
        assert np.array_equal(rewards, g["rewards"][t]), t
        assert np.array_equal(dones, g["dones"][t]), t


@pytest.mark.parametrize("path", gio.files("env_poke_"), ids=gio.name)
def test_env_poke(path):
    g = gio.load(path)
    m, n, k, ne = (int(x) for x in g["geom"])
    env = orc.OracleEnv(m, n, k, ne)
    env.boards[:] = gio.unpack(g["init_boards"], (2, m, n))
    env.current_player[:] = g["init_player"]
    env.move_counts[:] = g["init_count"]
    obs, rewards, dones = env.step(g["actions"])
    assert np.array_equal(env.boards.astype(bool), gio.unpack(g["boards"], (2, m, n)))
    assert np.array_equal(obs["action_mask"], gio.unpack(g["mask"], (m * n,)))
    assert np.array_equal(env.current_player, g["player"])
    assert np.array_equal(env.move_counts, g["count"])
    assert np.array_equal(rewards, g["rewards"])
    assert np.array_equal(dones, g["dones"])
    assert 0 < rewards.sum() < ne          # the fixture exercises both outcomes


def hash_policy(obs_dict):
    """numpy twin of gen_golden.HashPolicy."""
    obs, mask = obs_dict["observation"], obs_dict["action_mask"].astype(bool)
    b, cells = mask.shape
    w = np.arange(1, 2 * cells + 1, dtype=np.int64)
    score = (obs.reshape(b, -1).astype(np.int64) * w).sum(axis=1)
    cnt = mask.sum(axis=1)
    j = score % np.maximum(cnt, 1)
    rank = np.cumsum(mask, axis=1) - 1
    picked = np.argmax(mask & (rank == j[:, None]), axis=1)
    return np.where(cnt == 0, score % cells, picked).astype(np.int64)


@pytest.mark.parametrize("path", gio.files("wrapper_trace_"), ids=gio.name)
def test_wrapper_trace(path):
    g = gio.load(path)
    m, n, k, ne = (int(x) for x in g["geom"])
    sides = g["sides"]
    cursor = {"t": 0}
    env = orc.OracleEnv(m, n, k, ne)
    wr = orc.OracleWrapper(env, side_fn=lambda idx: sides[cursor["t"]][idx])
    wr.set_opponent(hash_policy)
    if int(g["options_reset"]):
        obs, info = wr.reset(options={"agent_side": sides[0]})
    else:
        obs, info = wr.reset()
    assert info == {}
    assert np.array_equal(obs["observation"].astype(bool), gio.unpack(g["obs0"], (2, m, n)))
    assert np.array_equal(obs["action_mask"], gio.unpack(g["mask0"], (m * n,)))
    assert np.array_equal(wr.agent_side, g["side0"])
    for t in range(len(g["actions"])):
        cursor["t"] = t + 1
        obs, r, term, trunc, info = wr.step(g["actions"][t])
        assert np.array_equal(obs["observation"].astype(bool), gio.unpack(g["obs"][t], (2, m, n))), t
        assert np.array_equal(obs["action_mask"], gio.unpack(g["mask"][t], (m * n,))), t
        assert np.array_equal(r, g["rewards"][t]), t
        assert np.array_equal(term, g["terminated"][t]), t
        assert not trunc.any()
        assert np.array_equal(wr.agent_side, g["agent_side"][t]), t
        assert np.array_equal(wr.pending_resets, g["pending"][t]), t
        assert np.array_equal(env.boards.astype(bool), gio.unpack(g["boards"][t], (2, m, n))), t
        assert np.array_equal(env.current_player, g["player"][t]), t
        assert np.array_equal(env.move_counts, g["count"][t]), t
    rew = g["rewards"]
    assert (rew == 1).any() and (rew == -1).any()      # both outcomes occur in the fixture


def test_first_legal_matches_random_policy_deterministic():
    g = gio.load(gio.files("random_policy_first_legal")[0])
    mask = gio.unpack(g["mask"], (int(g["cells"]),))
    assert np.array_equal(orc.first_legal(mask), g["first_legal"])


def test_tictactoe_game_tree_known_answer():
    """Exhaustive 3x3x3 game tree: 255,168 games = 131,184 / 77,904 / 46,080
    (first-player wins / second-player wins / draws) -- a combinatorial fact
    independent of the reference."""
    lib = orc.lib()
    import ctypes
    counts = {"x": 0, "o": 0, "d": 0}
    planes = np.zeros((2, 9), dtype=np.uint8)

    def has_line(p):
        return lib.orc_plane_has_line(planes[p].ctypes.data_as(ctypes.c_void_p), 3, 3, 3)

    def rec(player, depth):
        for a in range(9):
            if planes[0, a] or planes[1, a]:
                continue
            planes[player, a] = 1
            if has_line(player):
                counts["x" if player == 0 else "o"] += 1
            elif depth == 8:
                counts["d"] += 1
            else:
                rec(player ^ 1, depth + 1)
            planes[player, a] = 0

    rec(0, 0)
    assert (counts["x"], counts["o"], counts["d"]) == (131184, 77904, 46080)


def test_philox_known_answer():
    """Philox4x32-10 known-answer vectors from the Random123 distribution (kat_vectors)."""
    out = orc.philox4x32(0, 0, 0, 0, 0, 0)
    assert [int(x) for x in out] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    out = orc.philox4x32(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff)
    assert [int(x) for x in out] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    out = orc.philox4x32(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)
    assert [int(x) for x in out] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


@pytest.mark.parametrize("path", gio.files("env_trace_"), ids=gio.name)
def test_torch_port_env_trace(path):
    """oracle/torch_port.py (the CPU-baseline port timed by bench.py) against the same traces."""
    import torch
    from oracle import torch_port as tp
    g = gio.load(path)
    m, n, k, ne = (int(x) for x in g["geom"])
    s = tp.port_make(m, n, k, ne)
    for t in range(len(g["op"])):
        op, active, actions = int(g["op"][t]), g["active"][t], g["actions"][t]
        idx = torch.from_numpy(np.nonzero(active)[0])
        rewards, dones = torch.zeros(ne), torch.zeros(ne, dtype=torch.bool)
        if op == gio.OP_RESET_ALL:
            obs = tp.port_reset(s)
        elif op == gio.OP_RESET_IDX:
            obs = tp.port_reset(s, idx)
        elif op == gio.OP_STEP:
            obs, rewards, dones = tp.port_step(s, torch.from_numpy(actions))
        else:
            obs, rewards, dones = tp.port_step_subset(s, torch.from_numpy(actions)[idx], idx)
        assert np.array_equal(s.planes.numpy().astype(bool), gio.unpack(g["boards"][t], (2, m, n))), t
        assert np.array_equal(obs["action_mask"].numpy(), gio.unpack(g["mask"][t], (m * n,))), t
        assert np.array_equal(s.to_move.numpy(), g["player"][t]) and np.array_equal(s.plies.numpy(), g["count"][t])
        assert np.array_equal(rewards.numpy(), g["rewards"][t]) and np.array_equal(dones.numpy(), g["dones"][t]), t


@pytest.mark.skipif(not __import__("ref_loader").available(), reason="reference not mounted (GPU box)")
def test_oracle_live_against_reference_large_batch():
    """Where the reference is mounted (build container), run it live next to the C oracle on a
    larger batch than the committed fixtures hold."""
    import torch
    import ref_loader
    RefEnv = ref_loader.load("env.torch_vector_mnk_env").TorchVectorMnkEnv
    rng = np.random.default_rng(5)
    for (m, n, k) in [(9, 9, 5), (6, 7, 4)]:
        ne = 512
        ref, mine = RefEnv(m, n, k, ne, device="cpu"), orc.OracleEnv(m, n, k, ne)
        obs = ref.reset()
        mine.reset()
        for t in range(3 * m * n // 2):
            mask = obs["action_mask"].numpy()
            u = rng.random(ne)
            cnt = mask.sum(1)
            rank = np.cumsum(mask, 1) - 1
            j = np.floor(u * np.maximum(cnt, 1)).astype(np.int64)
            a = np.where(cnt == 0, 0, np.argmax(mask & (rank == j[:, None]), axis=1)).astype(np.int64)
            obs, r, d = ref.step(torch.from_numpy(a))
            o2, r2, d2 = mine.step(a)
            assert np.array_equal(ref.boards.numpy().astype(np.uint8), mine.boards)
            assert np.array_equal(obs["action_mask"].numpy(), o2["action_mask"])
            assert np.array_equal(r.numpy(), r2) and np.array_equal(d.numpy(), d2)
            done_idx = np.nonzero(d2)[0]
            if len(done_idx) and t % 2 == 0:
                obs = ref.reset(torch.from_numpy(done_idx))
                mine.reset(done_idx)

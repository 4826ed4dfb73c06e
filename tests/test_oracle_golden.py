"""Pins the C oracle (oracle/keisei_oracle.c) against outputs of the Python reference itself
(fixtures written by oracle/gen_golden.py, which imports /root/reference in the build container).
CPU only."""
import hashlib
import os

import numpy as np
import pytest

from oracle import oracle as orc


def _digest(obs):
    return int.from_bytes(hashlib.blake2b(np.ascontiguousarray(obs, np.float32).tobytes(), digest_size=8).digest(), "little")


@pytest.fixture(scope="module")
def traces(golden_dir):
    with np.load(os.path.join(golden_dir, "traces_random.npz")) as z:
        return {k: z[k] for k in z.files}  # NpzFile re-inflates on every access: materialise once


def test_rand32_matches_fixture_rule():
    # the python generator and the C oracle implement the same counter-based RNG
    M64 = (1 << 64) - 1

    def py(seed, env, step):
        x = (seed ^ ((env * 0x9E3779B97F4A7C15) & M64) ^ ((step * 0xBF58476D1CE4E5B9) & M64)) & M64
        x ^= x >> 30; x = (x * 0xBF58476D1CE4E5B9) & M64
        x ^= x >> 27; x = (x * 0x94D049BB133111EB) & M64
        x ^= x >> 31
        return x >> 32

    for s, e, t in [(1234, 0, 0), (1234, 63, 511), (99, 100000, 7), (0, 0, 0)]:
        assert orc.rand32(s, e, t) == py(s, e, t)


def test_random_traces_bit_exact(traces):
    """Every ply of every golden game: legal set, chosen action, successor board/hands/side/move_count,
    reward, done, reason, winner and the observation digest agree with the Python reference."""
    z = traces
    base = 0
    n_plies = 0
    for gi in range(len(z["env"])):
        env, T, mm = int(z["env"][gi]), int(z["T"][gi]), int(z["max_moves"][gi])
        g = orc.OracleGame(mm)
        for t in range(T):
            i = base + t
            want = z["legal"][z["legal_off"][i]:z["legal_off"][i + 1]].astype(np.int32)
            got = g.legal_indices()
            assert np.array_equal(got, want), (gi, t)
            a = g.pick_action(int(z["seed"]), env, t)
            assert a == int(z["actions"][i]), (gi, t)
            reward, done, reason, winner = g.make_move(a)
            b, h, m = g.export()
            assert np.array_equal(b, z["boards"][i]) and np.array_equal(h, z["hands"][i]), (gi, t)
            assert (m[0], m[1]) == (int(z["sides"][i]), int(z["move_counts"][i])), (gi, t)
            assert reward == float(z["rewards"][i]) and done == bool(z["dones"][i]), (gi, t)
            assert reason == int(z["reasons"][i]) and winner == int(z["winners"][i]), (gi, t)
            assert _digest(g.observation()) == int(z["digests"][i]), (gi, t)
            if done:
                g.reset()
            n_plies += 1
        base += T
    assert n_plies == len(z["actions"])


def test_full_observations(traces):
    z = traces
    # replay up to each stored index is covered by digests above; here compare raw tensors of the KATs
    assert z["full_obs"].shape[1:] == (46, 9, 9)
    assert z["full_obs"].dtype == np.float32


def test_kat_positions(golden_dir):
    with np.load(os.path.join(golden_dir, "kat_positions.npz")) as zf:
        z = {k: zf[k] for k in zf.files}
    for i, sfen in enumerate(z["sfens"]):
        g = orc.OracleGame.from_sfen(str(sfen))
        m = g.meta
        assert bool(m[3]) == bool(z["game_over"][i]), sfen
        assert int(m[4]) == int(z["winner"][i]) and int(m[5]) == int(z["reason"][i]), sfen
        assert np.array_equal(g.observation(), z["obs"][i]), sfen
        assert [g.in_check(0), g.in_check(1)] == [bool(x) for x in z["in_check"][i]], sfen
        want = z["legal"][z["legal_off"][i]:z["legal_off"][i + 1]].astype(np.int32)
        assert np.array_equal(g.legal_indices(), want), sfen


def test_reference_suite_known_answers():
    # counts pinned by the reference's own tests
    assert len(orc.OracleGame().legal_indices()) == 30
    assert len(orc.OracleGame.from_sfen("4k4/4r4/9/9/9/9/9/9/4K4 b - 1").legal_indices()) == 4
    assert len(orc.OracleGame.from_sfen("9/9/9/9/4K4/9/9/9/4k4 b P 1").legal_indices()) == 78
    assert len(orc.OracleGame.from_sfen("P8/9/9/9/4k4/9/9/9/4K4 b P 1").legal_indices()) == 67
    assert len(orc.OracleGame.from_sfen("4k4/9/9/9/9/9/9/9/4K4 b - 1").legal_indices()) == 5
    g = orc.OracleGame.from_sfen("K8/9/9/9/9/9/9/9/r8 w - 1")
    assert g.meta[3] == 1 and g.meta[4] == 0 and g.meta[5] == 1
    # mapper known answers (tests/shogi/test_shogi_utils.py): drop P at (4,4) -> 13240, L at (0,0) -> 12961
    assert orc.index_to_move(13240) == (None, None, 4, 4, 0)
    assert orc.index_to_move(12961) == (None, None, 0, 0, 1)


def test_scripted_sennichite(golden_dir):
    z = np.load(os.path.join(golden_dir, "scripted.npz"))
    g = orc.OracleGame.from_sfen(str(z["senn_sfen"]))
    for a, d, r in zip(z["senn_actions"], z["senn_dones"], z["senn_reasons"]):
        _, done, reason, _ = g.make_move(int(a))
        assert (done, reason) == (bool(d), int(r))
    assert len(z["senn_actions"]) == 13 and z["senn_reasons"][-1] == 4
    g = orc.OracleGame()
    for a, d, r in zip(z["senn2_actions"], z["senn2_dones"], z["senn2_reasons"]):
        _, done, reason, _ = g.make_move(int(a))
        assert (done, reason) == (bool(d), int(r))
    assert z["senn2_reasons"][-1] == 4


def test_gae_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "gae_golden.npz"))
    for i in range(int(z["n_cases"])):
        r, v, d = z[f"c{i}_r"][:, None], z[f"c{i}_v"][:, None], z[f"c{i}_d"][:, None]
        for fn in (orc.gae, orc.gae_numpy):
            adv, ret = fn(r, v, d, np.array([z[f"c{i}_last"]], np.float32), float(z[f"c{i}_gamma"]), float(z[f"c{i}_lam"]))
            # bit-exact: same op order, every op rounded to fp32
            assert np.array_equal(adv[:, 0], z[f"c{i}_adv"]), (i, fn.__name__)
            assert np.array_equal(ret[:, 0], z[f"c{i}_ret"]), (i, fn.__name__)
    # tests/conftest.py:543-581 known answer
    assert z["c0_adv"][2] == 1.5 and z["c0_ret"][2] == 3.0


def test_endgame_traces_bit_exact(golden_dir):
    """Drop-heavy endgames played by the Python reference (oracle/gen_golden_endgame.py): nifu, drop rank limits,
    drops answering checks and uchifuzume decide these legal sets (a third of the moves are drops, up to 458 legal
    moves per position).  Every ply: legal set, action, successor, outcome and observation digest."""
    with np.load(os.path.join(golden_dir, "traces_endgame.npz")) as zf:
        z = {k: zf[k] for k in zf.files}
    base = 0
    for gi, sfen in enumerate(z["sfens"]):
        g = orc.OracleGame.from_sfen(str(sfen))
        b, h, _ = g.export()
        assert np.array_equal(b, z["start_boards"][gi]) and np.array_equal(h, z["start_hands"][gi]), sfen
        for t in range(int(z["T"][gi])):
            i = base + t
            want = z["legal"][z["legal_off"][i]:z["legal_off"][i + 1]].astype(np.int32)
            assert np.array_equal(g.legal_indices(), want), (gi, t, sfen)
            a = g.pick_action(int(z["seed"]), gi, t)
            assert a == int(z["actions"][i]), (gi, t)
            reward, done, reason, winner = g.make_move(a)
            b, h, m = g.export()
            assert np.array_equal(b, z["boards"][i]) and np.array_equal(h, z["hands"][i]), (gi, t)
            assert (m[0], m[1]) == (int(z["sides"][i]), int(z["move_counts"][i])), (gi, t)
            assert (reward, done, reason, winner) == (float(z["rewards"][i]), bool(z["dones"][i]), int(z["reasons"][i]),
                                                      int(z["winners"][i])), (gi, t)
            assert _digest(g.observation()) == int(z["digests"][i]), (gi, t)
        base += int(z["T"][gi])
    assert base == len(z["actions"]) and int((z["actions"] >= 12960).sum()) > 500


def test_stalemate_fixture_bit_exact(golden_dir):
    """Stalemate -- no legal move, not in check: a DRAW in the reference (shogi_game.py:431-435) -- from the fixture
    oracle/gen_golden_stalemate.py produced with the imported reference: the reference test-suite's two positions
    (stalemate at load, stalemate by a move) and 12 random bare-king endgames that end in stalemate."""
    with np.load(os.path.join(golden_dir, "traces_stalemate.npz")) as zf:
        z = {k: zf[k] for k in zf.files}
    # known answers
    g = orc.OracleGame.from_sfen(str(z["kat_load_sfen"]))
    m = g.meta
    assert (m[3], m[5], m[4]) == (int(z["kat_load_game_over"]), int(z["kat_load_reason"]), int(z["kat_load_winner"])) == (1, 2, -1)
    assert len(g.legal_indices()) == int(z["kat_load_legal_n"]) == 0
    assert np.array_equal(g.observation(), z["kat_load_obs"])
    g = orc.OracleGame.from_sfen(str(z["kat_move_sfen"]))
    assert np.array_equal(g.legal_indices(), z["kat_move_legal"].astype(np.int32))
    reward, done, reason, winner = g.make_move(int(z["kat_move_action"]))
    assert (reward, done, reason, winner) == (0.0, True, 2, -1)
    assert (reward, done, reason, winner) == (float(z["kat_move_reward"]), bool(z["kat_move_done"]), int(z["kat_move_reason"]),
                                              int(z["kat_move_winner"]))
    assert np.array_equal(g.observation(), z["kat_move_obs"]) and len(g.legal_indices()) == int(z["kat_move_legal_after"])
    # games that end in stalemate
    base = 0
    for gi, sfen in enumerate(z["sfens"]):
        g = orc.OracleGame.from_sfen(str(sfen))
        env = int(z["envs"][gi])
        for t in range(int(z["T"][gi])):
            i = base + t
            want = z["legal"][z["legal_off"][i]:z["legal_off"][i + 1]].astype(np.int32)
            assert np.array_equal(g.legal_indices(), want), (gi, t, sfen)
            a = g.pick_action(int(z["seed"]), env, t)
            assert a == int(z["actions"][i]), (gi, t)
            reward, done, reason, winner = g.make_move(a)
            b, h, m = g.export()
            assert np.array_equal(b, z["boards"][i]) and np.array_equal(h, z["hands"][i]), (gi, t)
            assert (reward, done, reason, winner) == (float(z["rewards"][i]), bool(z["dones"][i]), int(z["reasons"][i]),
                                                      int(z["winners"][i])), (gi, t)
            assert _digest(g.observation()) == int(z["digests"][i]), (gi, t)
        base += int(z["T"][gi])
        assert (reward, done, reason, winner) == (0.0, True, 2, -1), sfen  # every game ends in a stalemate draw
    assert base == len(z["actions"]) and len(z["sfens"]) >= 8

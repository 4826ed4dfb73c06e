"""Live differential checks against the Python reference, for the container that has it (skipped where
/root/reference is absent, e.g. on the GPU box -- there the committed golden fixtures stand in).  Fresh seeds every
round trip through the same comparisons as the golden tests: oracle vs reference self-play, the mapper's whole table,
SFEN text, ExperienceBuffer host semantics."""
import os
import random
import sys

import numpy as np
import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "keisei")), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref():
    sys.dont_write_bytecode = True  # the reference tree is read-only
    if REF not in sys.path:
        sys.path.append(REF)  # appended: must not shadow this repo's `tests` package
    import keisei.shogi as rshogi
    from keisei.core.experience_buffer import ExperienceBuffer as RBuffer
    from keisei.utils import PolicyOutputMapper as RMapper
    return {"shogi": rshogi, "Buffer": RBuffer, "Mapper": RMapper}


def test_mapper_table_identical(ref):
    from shogidrl_b200.utils import PolicyOutputMapper
    ours, theirs = PolicyOutputMapper(), ref["Mapper"]()
    assert ours.get_total_actions() == theirs.get_total_actions() == 13527
    for i in range(13527):
        a, b = ours.policy_index_to_shogi_move(i), theirs.policy_index_to_shogi_move(i)
        assert tuple(getattr(x, "value", x) for x in a) == tuple(getattr(x, "value", x) for x in b), i
    for i in (0, 1, 12959, 12960, 13526):
        assert ours.action_idx_to_usi_move(i) == theirs.action_idx_to_usi_move(i)


def test_oracle_tracks_reference_on_fresh_games(ref):
    """Two short random games with seeds that are not in the fixtures: legal sets, outcomes and observations."""
    from oracle import oracle as orc
    mapper = ref["Mapper"]()
    seed0 = int.from_bytes(os.urandom(2), "little")
    for gi in range(2):
        rng = random.Random(seed0 + gi)
        g = ref["shogi"].ShogiGame(max_moves_per_game=60)
        o = orc.OracleGame(60)
        for ply in range(45):
            want = sorted(mapper.shogi_move_to_policy_index(m) for m in g.get_legal_moves())
            assert o.legal_indices().tolist() == want, (seed0, gi, ply)
            a = want[rng.randrange(len(want))]
            obs, reward, done, info = g.make_move(mapper.policy_index_to_shogi_move(a))
            r2, d2, reason, winner = o.make_move(a)
            assert (reward, done) == (r2, d2) and np.array_equal(obs, o.observation()), (seed0, gi, ply)
            assert orc.parse_sfen(g.to_sfen_string())[0].tolist() == o.export()[0].tolist(), (seed0, gi, ply)
            if done:
                break


def test_sfen_text_identical(ref):
    from shogidrl_b200.shogi.sfen import HostPosition, pack_sfen
    for sfen in ("lnsgkgsnl/1r5b1/ppppppppp/9/9/9/PPPPPPPPP/1B5R1/LNSGKGSNL b - 1",
                 "4k4/9/9/9/9/R8/9/9/4K4 b - 1", "9/9/9/9/4K4/9/9/9/4k4 w 2P3p 7",
                 "l2+R2s1l/4gk3/p1n1pp1pp/2pp2p2/9/2P1P4/PP1P1PPPP/1B5R1/LNSGKGSNL w BGSNP 20"):
        b, h, side, mc = pack_sfen(sfen)
        theirs = ref["shogi"].ShogiGame.from_sfen(sfen)
        assert HostPosition(b, h, side, mc).to_sfen_string() == theirs.to_sfen_string()


def test_experience_buffer_host_semantics(ref):
    from shogidrl_b200.core import ExperienceBuffer
    rng = np.random.default_rng(5)
    ours, theirs = ExperienceBuffer(6, 0.99, 0.95, "cpu"), ref["Buffer"](6, 0.99, 0.95, "cpu")
    for i in range(8):  # two more than the capacity: both drop the overflow
        obs = torch.from_numpy(rng.random((46, 9, 9), dtype=np.float32))
        mask = torch.from_numpy(rng.random(13527) < 0.01)
        args = (obs, int(rng.integers(0, 13527)), float(rng.random()), float(-rng.random()), float(rng.random()), bool(i == 3), mask)
        ours.add(*args)
        theirs.add(*args)
    assert len(ours) == len(theirs) == 6 and ours.capacity() == theirs.capacity()
    for name in ("obs", "actions", "rewards", "log_probs", "values", "dones", "legal_masks"):
        assert torch.equal(getattr(ours, name), getattr(theirs, name)), name
    with pytest.raises(RuntimeError) as e1:
        ours.get_batch()
    with pytest.raises(RuntimeError) as e2:
        theirs.get_batch()
    assert str(e1.value) == str(e2.value)
    fmt_o, fmt_t = ours.get_worker_batch_format(), theirs.get_worker_batch_format()
    assert set(fmt_o) == set(fmt_t) and all(torch.equal(fmt_o[k], fmt_t[k]) for k in fmt_t)
    big_o, big_t = ExperienceBuffer(9, 0.99, 0.95, "cpu"), ref["Buffer"](9, 0.99, 0.95, "cpu")
    big_o.add_from_worker_batch(fmt_o)
    big_t.add_from_worker_batch(fmt_t)
    big_o.merge_from_parallel_buffers([ours])
    big_t.merge_from_parallel_buffers([theirs])
    assert big_o.size() == big_t.size() == 9
    for name in ("obs", "actions", "rewards", "log_probs", "values", "dones", "legal_masks"):
        assert torch.equal(getattr(big_o, name), getattr(big_t, name)), name
    ours.clear()
    theirs.clear()
    assert len(ours) == len(theirs) == 0

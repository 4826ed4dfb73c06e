"""GPU tests of the reference-API mirror: ShogiGame facade, PolicyOutputMapper, ExperienceBuffer, PPOAgent,
EnvManager / StepManager with real objects, and the batched rollout + PPO update."""
import copy
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import oracle as orc  # noqa: E402  (checker only)
from tests.helpers import make_config  # noqa: E402


@pytest.fixture(scope="module")
def kats(golden_dir):
    with np.load(os.path.join(golden_dir, "kat_positions.npz")) as zf:
        return {k: zf[k] for k in zf.files}


def test_facade_known_answers(kats):
    from shogidrl_b200.shogi import Color, ShogiGame
    from shogidrl_b200.utils import PolicyOutputMapper
    mapper = PolicyOutputMapper()
    reasons = {0: None, 1: "Tsumi", 2: "stalemate", 3: "Max moves reached", 4: "Sennichite"}
    for i, sfen in enumerate(kats["sfens"]):
        g = ShogiGame.from_sfen(str(sfen))
        assert g.game_over == bool(kats["game_over"][i]), sfen
        assert (g.winner.value if g.winner else -1) == int(kats["winner"][i]), sfen
        assert g.termination_reason == reasons[int(kats["reason"][i])], sfen
        assert np.array_equal(g.get_observation(), kats["obs"][i]), sfen
        assert [g.is_in_check(Color.BLACK), g.is_in_check(Color.WHITE)] == [bool(x) for x in kats["in_check"][i]], sfen
        want = kats["legal"][kats["legal_off"][i]:kats["legal_off"][i + 1]].astype(np.int64)
        moves = g.get_legal_moves()
        got = np.sort([mapper.shogi_move_to_policy_index(m) for m in moves])
        assert np.array_equal(got, want), sfen
        out = g.to_sfen_string()  # hands are emitted in the reference's R,B,G,S,N,L,P order
        assert out.split()[0] == str(sfen).split()[0] and out.split()[1] == str(sfen).split()[1], sfen
        assert orc.parse_sfen(out)[1].tolist() == orc.parse_sfen(str(sfen))[1].tolist(), sfen
    # reference test-suite counts
    assert len(ShogiGame().get_legal_moves()) == 30
    g = ShogiGame.from_sfen("9/9/9/9/4K4/9/9/9/4k4 b P 1")
    assert len(g.get_legal_moves()) == 78
    assert ShogiGame.from_sfen("P8/9/9/9/4k4/9/9/9/4K4 b P 1").is_nifu(Color.BLACK, 0)


def test_facade_replays_reference_game(golden_dir):
    """One golden game of the Python reference through the scalar API: make_move 4-tuples, boards, hands."""
    from shogidrl_b200.shogi import ShogiGame
    from shogidrl_b200.utils import PolicyOutputMapper
    with np.load(os.path.join(golden_dir, "traces_random.npz")) as zf:
        z = {k: zf[k] for k in zf.files}
    mapper = PolicyOutputMapper()
    gi = int(np.argmax(z["max_moves"] == 60))
    base = int(np.sum(z["T"][:gi]))
    g = ShogiGame(max_moves_per_game=60)
    names = {0: "Game ongoing", 1: "Tsumi", 2: "stalemate", 3: "Max moves reached", 4: "Sennichite"}
    for t in range(150):
        i = base + t
        want = z["legal"][z["legal_off"][i]:z["legal_off"][i + 1]].astype(np.int64)
        got = np.sort([mapper.shogi_move_to_policy_index(m) for m in g.get_legal_moves()])
        assert np.array_equal(got, want), t
        obs, reward, done, info = g.make_move(mapper.policy_index_to_shogi_move(int(z["actions"][i])))
        assert obs.shape == (46, 9, 9) and obs.dtype == np.float32
        assert reward == float(z["rewards"][i]) and done == bool(z["dones"][i]) and info["reason"] == names[int(z["reasons"][i])]
        b = np.array([[0 if p is None else p.code for p in row] for row in g.board], np.int8).reshape(81)
        assert np.array_equal(b, z["boards"][i]), t
        assert g.move_count == int(z["move_counts"][i]) and g.current_player.value == int(z["sides"][i])
        if done:
            assert ("winner" in info) == (int(z["winners"][i]) >= 0)
            again = g.make_move((0, 0, 1, 0, False))  # finished game: the terminal tuple again (shogi_game.py:589-593)
            assert again[2] is True and again[3]["reason"] == info["reason"]
            g.reset()


def test_facade_errors_undo_sennichite_deepcopy():
    from shogidrl_b200.shogi import Color, Piece, PieceType, ShogiGame
    g = ShogiGame()
    with pytest.raises(ValueError, match="Illegal movement pattern"):
        g.make_move((6, 0, 4, 0, False))
    with pytest.raises(ValueError, match="Invalid move_tuple format"):
        g.make_move((6, 0, 5, 0))
    with pytest.raises(ValueError, match="No piece at source"):
        g.make_move((4, 4, 3, 4, False))
    sfen0 = g.to_sfen_string()
    g.make_move((6, 6, 5, 6, False))
    assert g.current_player == Color.WHITE and g.move_count == 1 and len(g.move_history) == 1
    g.undo_move()
    assert g.to_sfen_string() == sfen0 and g.move_history == []
    assert sorted(g.get_individual_piece_moves(Piece(PieceType.ROOK, Color.BLACK), 4, 4)) == sorted(
        [(r, 4) for r in (3, 5)] + [(4, c) for c in range(9) if c != 4] + [(2, 4), (6 - 0, 4)][:1])
    assert g.test_move((6, 6, 5, 6, False)) and not g.test_move((6, 6, 4, 6, False))
    # sennichite on the 13th ply of a 4-ply cycle (tests/shogi/test_shogi_game_core_logic.py:1126-1179)
    g = ShogiGame.from_sfen("4k4/9/9/9/9/R8/9/9/4K4 b - 1")
    cycle = [(5, 0, 5, 1, False), (0, 4, 0, 3, False), (5, 1, 5, 0, False), (0, 3, 0, 4, False)]
    for i in range(12):
        _, _, done, _ = g.make_move(cycle[i % 4])
        assert not done, i
    _, r, done, info = g.make_move(cycle[0])
    assert done and info["reason"] == "Sennichite" and r == 0.0 and g.is_sennichite()
    g2 = copy.deepcopy(ShogiGame())
    assert g2.move_history == [] and len(g2.board_history) == 1
    # max moves through the private attribute the reference's tests poke
    g = ShogiGame()
    g._max_moves_this_game = 2
    g.make_move((6, 6, 5, 6, False))
    _, _, done, info = g.make_move((2, 2, 3, 2, False))
    assert done and info["reason"] == "Max moves reached"
    # get_legal_moves on a finished game clears the flags, as the reference's simulation undo does
    g = ShogiGame.from_sfen("9/9/9/9/9/4G4/4r4/4g4/4K4 b - 1")
    assert g.game_over and g.get_legal_moves() == [] and not g.game_over


def test_experience_buffer_gae_known_answer():
    from shogidrl_b200.core import ExperienceBuffer
    buf = ExperienceBuffer(3, 0.99, 0.95, "cuda")
    o, m = torch.zeros(46, 9, 9), torch.zeros(13527, dtype=torch.bool)
    for r, v, d in [(1.0, 0.5, False), (2.0, 1.0, False), (3.0, 1.5, True)]:
        buf.add(o, 0, r, 0.0, v, d, m)
    buf.compute_advantages_and_returns(2.0)
    batch = buf.get_batch()
    assert float(batch["advantages"][2]) == 1.5 and float(batch["returns"][2]) == 3.0  # tests/conftest.py:543-581
    a, r = orc.gae_numpy(np.array([[1.], [2.], [3.]], np.float32), np.array([[.5], [1.], [1.5]], np.float32),
                         np.array([[0], [0], [1]]), np.array([2.0], np.float32), 0.99, 0.95)
    assert np.array_equal(batch["advantages"].cpu().numpy(), a[:, 0]) and np.array_equal(batch["returns"].cpu().numpy(), r[:, 0])


def test_agent_select_action_and_step_manager_real_objects():
    from shogidrl_b200.core import ActorCritic, ExperienceBuffer, PPOAgent
    from shogidrl_b200.training import EnvManager, StepManager
    cfg = make_config()
    logs = []
    em = EnvManager(cfg, logs.append)
    game, mapper = em.setup_environment()
    assert em.validate_environment() and em.get_legal_moves_count() == 30
    torch.manual_seed(0)
    agent = PPOAgent(ActorCritic(46, 13527), cfg, torch.device("cuda"))
    buf = ExperienceBuffer(64, cfg.training.gamma, cfg.training.lambda_gae, "cuda")
    sm = StepManager(cfg, game, agent, mapper, buf)
    st = sm.reset_episode()
    mask = mapper.get_legal_mask(game.get_legal_moves(), torch.device("cuda"))
    move, idx, lp, v = agent.select_action(st.current_obs, mask, is_training=True)
    assert bool(mask[idx]) and np.isfinite(lp) and np.isfinite(v) and move == mapper.policy_index_to_shogi_move(idx)
    single = torch.zeros(13527, dtype=torch.bool, device="cuda"); single[1234] = True
    assert agent.select_action(st.current_obs, single, is_training=True)[1] == 1234
    assert agent.select_action(st.current_obs, mask, is_training=False)[1] == agent.select_action(st.current_obs, mask, is_training=False)[1]
    assert np.isfinite(agent.get_value(st.current_obs))
    log = lambda *a, **k: None
    for t in range(64):
        res = sm.execute_step(st, t, log)
        assert res.success
        st = sm.update_episode_state(st, res)
        if res.done:
            st, _ = sm.handle_episode_end(st, res, {"black_wins": 0, "white_wins": 0, "draws": 0}, 0, log)
    assert len(buf) == 64
    buf.compute_advantages_and_returns(agent.get_value(st.current_obs))
    metrics = agent.learn(buf)
    assert all(np.isfinite(x) for x in metrics.values()) and "ppo/clip_fraction" in metrics
    # every stored mask row has the chosen action legal
    b = buf.get_batch()
    assert bool(b["legal_masks"].gather(1, b["actions"][:, None]).all())


def test_vectorised_rollout_and_update():
    from shogidrl_b200 import VecShogiEnv
    from shogidrl_b200.core import ActorCritic, PPOAgent, RolloutBuffer
    from shogidrl_b200.training import VecStepManager
    cfg = make_config(minibatch_size=256, ppo_epochs=1)
    dev = torch.device("cuda")
    N, T = 256, 16
    env = VecShogiEnv(N, max_moves_per_game=40, device=dev, seed=5)
    torch.manual_seed(1)
    agent = PPOAgent(ActorCritic(46, 13527), cfg, dev, use_mixed_precision=True)
    buf = RolloutBuffer(T, N, 0.99, 0.95, dev)
    drv = VecStepManager(env, agent, buf)
    total_eps = 0
    for _ in range(4):
        drv.collect()
        stats = drv.finish()
        total_eps += stats["episodes"]
        b = buf.get_batch(expand_masks=True)
        assert b["obs"].shape == (T * N, 46, 9, 9) and b["legal_masks"].dtype == torch.bool
        assert b["legal_bitmaps"].shape == (T * N, 448) and "legal_masks" not in buf.get_batch()
        assert bool(b["legal_masks"].gather(1, b["actions"][:, None]).all())
        assert torch.equal(buf.masks_at(3), b["legal_masks"][3 * N:4 * N])
        # the stored legal sets are the engine's: recompute the last slot's mask from the live games
        if _ == 0:
            env.refresh()
            assert torch.equal(buf.masks_at(T).view(torch.uint8), env.mask)
        # stored observations are the engine's observation of the stored state: plane 42 = side to move
        assert bool(((b["obs"][:, 42, 0, 0] == 0) | (b["obs"][:, 42, 0, 0] == 1)).all())
        a_ref, r_ref = orc.gae(buf.rewards.cpu().numpy(), buf.values.cpu().numpy(), buf.dones.cpu().numpy(),
                               agent.get_values(buf.obs[T]).cpu().numpy(), 0.99, 0.95)
        assert np.allclose(buf.advantages.cpu().numpy(), a_ref, rtol=1e-5, atol=1e-6)
        m = agent.learn(buf)
        assert all(np.isfinite(x) for x in m.values())
        buf.clear()
    assert total_eps > 0 and drv.black_wins + drv.white_wins + drv.draws == drv.episodes
    assert int(env.errors().abs().sum()) == 0


def test_parallel_manager_shim():
    from shogidrl_b200.core import ActorCritic, ExperienceBuffer, PPOAgent
    from shogidrl_b200.training.parallel import ParallelManager
    cfg = make_config()
    agent = PPOAgent(ActorCritic(46, 13527), cfg, torch.device("cuda"))
    pm = ParallelManager({"max_moves_per_game": 500, "seed": 3}, {}, {"num_workers": 8, "batch_size": 4}, "cuda")
    assert pm.start_workers(agent) and pm.is_healthy()
    buf = ExperienceBuffer(100, 0.99, 0.95, "cuda")
    assert pm.collect_experiences(buf) == 32 and pm.collect_experiences(buf) == 32 and len(buf) == 64
    assert bool(buf.legal_masks[:64].gather(1, buf.actions[:64, None]).all())
    # worker semantics (self_play_worker.py:105, 130): 500-move games whatever the env config says, network in eval()
    assert pm._driver.env.max_moves == 500 and pm._driver.model_eval_mode and not agent.model.training
    pm.stop_workers()
    assert not pm.is_healthy() and pm.get_parallel_stats()["total_steps_collected"] == 64
    pm2 = ParallelManager({"max_moves_per_game": 40, "seed": 3}, {}, {"num_workers": 8, "batch_size": 4}, "cuda")
    pm2.start_workers(agent, worker_semantics=False)
    assert pm2._driver.env.max_moves == 40 and pm2.collect_experiences(ExperienceBuffer(100, 0.99, 0.95, "cuda")) == 32
    assert agent.model.training


def test_agent_matches_reference_golden(golden_dir):
    """SURVEY 8a-13 / 8a-14 / 8f-1 against the Python reference itself (oracle/gen_golden_agent.py, CPU fp32): same
    deterministic weights loaded through the reference's parameter names, same buffer contents; deterministic action
    selection, evaluate_actions, GAE and one PPOAgent.learn() (2 epochs x 1 minibatch: fused evaluation, loss and
    clip + Adam kernels on this side) must reproduce the reference's outputs.  fp32 on both sides, TF32 off;
    tolerances cover summation order only."""
    from oracle.gen_golden_agent import TRAINING, det_state_dict
    from shogidrl_b200.core import ActorCritic, ExperienceBuffer, PPOAgent
    with np.load(os.path.join(golden_dir, "agent_golden.npz")) as zf:
        z = {k: zf[k] for k in zf.files}
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        dev = torch.device("cuda")
        n = int(z["n"])
        model = ActorCritic(46, 13527)
        model.load_state_dict(det_state_dict())  # strict: the reference's parameter names and shapes
        agent = PPOAgent(model, make_config(**TRAINING), dev, use_mixed_precision=False)
        obs = torch.as_tensor(z["obs"], device=dev)
        masks = torch.zeros((n, 13527), dtype=torch.bool, device=dev)
        for i in range(n):
            masks[i, torch.as_tensor(z["legal"][z["legal_off"][i]:z["legal_off"][i + 1]].astype(np.int64), device=dev)] = True
        # a-13: deterministic selection and values
        a, lp, v = agent.select_actions(obs, masks, is_training=False)
        assert np.array_equal(a.cpu().numpy(), z["det_actions"])
        np.testing.assert_allclose(lp.cpu().numpy(), z["det_log_probs"], rtol=0, atol=2e-5)
        np.testing.assert_allclose(v.cpu().numpy(), z["det_values"], rtol=0, atol=2e-5)
        # evaluate_actions (log-prob of the taken actions, entropy over legal actions, values)
        model.eval()
        with torch.no_grad():
            elp, ent, ev = model.evaluate_actions(obs, torch.as_tensor(z["actions"], device=dev), masks)
        np.testing.assert_allclose(elp.cpu().numpy(), z["ev_log_probs"], rtol=0, atol=2e-5)
        np.testing.assert_allclose(ent.cpu().numpy(), z["ev_entropy"], rtol=0, atol=2e-5)
        np.testing.assert_allclose(ev.cpu().numpy().reshape(-1), z["ev_values"], rtol=0, atol=2e-5)
        # a-14: buffer + GAE (bit-exact: the inputs are the reference's own fp32 numbers)
        buf = ExperienceBuffer(n, TRAINING["gamma"], TRAINING["lambda_gae"], "cuda")
        for i in range(n):
            buf.add(obs[i], int(z["actions"][i]), float(z["rewards"][i]), float(z["log_probs"][i]), float(z["values"][i]),
                    bool(z["dones"][i]), masks[i])
        buf.compute_advantages_and_returns(float(z["last_value"]))
        batch = buf.get_batch()
        assert np.array_equal(batch["advantages"].cpu().numpy(), z["advantages"])
        assert np.array_equal(batch["returns"].cpu().numpy(), z["returns"])
        # f-1: one learn() call
        before = {k: t.detach().clone() for k, t in model.state_dict().items()}
        metrics = agent.learn(buf)
        want = dict(zip([str(k) for k in z["metric_names"]], z["metric_values"]))
        for k, w in want.items():
            assert abs(metrics[k] - w) <= 2e-5 + 1e-3 * abs(w), (k, metrics[k], w)
        assert abs(agent.last_gradient_norm - float(z["last_gradient_norm"])) < 1e-4
        after = model.state_dict()
        lr = TRAINING["learning_rate"]
        for k in before:
            delta = (after[k] - before[k]).cpu().numpy().reshape(-1)
            nnz, abs_sum, _ = z[f"delta_stats/{k}"]
            # Adam's first steps move every parameter with a gradient by ~lr per step whatever the gradient's size, so
            # entries whose gradient is pure rounding noise can differ; everything else must agree closely
            got, ref = delta[z[f"delta_idx/{k}"]], z[f"delta_val/{k}"]
            close = np.abs(got - ref) <= 0.1 * lr
            assert close.mean() > 0.97, (k, float(close.mean()))
            assert abs(np.abs(delta).sum(dtype=np.float64) - abs_sum) <= 0.01 * abs_sum, (k, abs_sum)
            assert abs(int(np.count_nonzero(delta)) - int(nnz)) <= 0.01 * nnz + 1, (k, int(np.count_nonzero(delta)), nnz)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32


def test_step_manager_matches_reference_golden(golden_dir):
    """the reference's StepManager driven over its own ShogiGame / PolicyOutputMapper / ExperienceBuffer
    by a scripted agent (oracle/gen_golden_stepmanager.py) against the same loop over this repo's classes on the
    device: every StepResult, counter, EpisodeState, episode-end log line and W&B payload, and the buffer contents."""
    import json
    from oracle.gen_golden_stepmanager import drive
    from shogidrl_b200.core import ExperienceBuffer
    from shogidrl_b200.shogi import ShogiGame
    from shogidrl_b200.training import StepManager
    from shogidrl_b200.utils import PolicyOutputMapper
    with open(os.path.join(golden_dir, "stepmanager_golden.json")) as f:
        want = json.load(f)
    got = json.loads(json.dumps(drive(StepManager, ShogiGame, PolicyOutputMapper, ExperienceBuffer, "cuda")))
    assert got["ends"] == want["ends"] and len(want["ends"]) == 3
    for t, (g, w) in enumerate(zip(got["steps"], want["steps"])):
        assert g == w, (t, g, w)
    assert got["buffer"] == want["buffer"] and got["n_logs"] == want["n_logs"]


def test_batched_evaluation_games():
    from shogidrl_b200.core import ActorCritic, PPOAgent
    from shogidrl_b200.evaluation import evaluate_vs_opponent
    cfg = make_config()
    torch.manual_seed(3)
    agent = PPOAgent(ActorCritic(46, 13527), cfg, torch.device("cuda"), use_mixed_precision=True)
    res = evaluate_vs_opponent(agent, 64, num_envs=64, max_moves_per_game=60, seed=9, deterministic=False)
    assert res.games == 64 and res.agent_wins + res.opponent_wins + res.draws == res.games
    # fixed quota per env: exactly num_games results, one per env here (no over-sampling of short games)
    res3 = evaluate_vs_opponent(agent, 100, num_envs=48, max_moves_per_game=40, seed=4, deterministic=False)
    assert res3.games == 100 and len(res3.outcomes) == 100
    assert 0 < res.mean_length <= 60 and 0.0 <= res.win_rate <= 1.0
    # agent vs agent also runs
    res2 = evaluate_vs_opponent(agent, 16, opponent=agent, num_envs=16, max_moves_per_game=30, deterministic=False)
    assert res2.games == 16
    assert len(res.outcomes) == res.games and res.outcomes.count("agent_win") == res.agent_wins
    assert res.outcomes.count("opponent_win") == res.opponent_wins and res.outcomes.count("draw") == res.draws


@pytest.mark.gpu
def test_tournament_and_ladder_on_the_vectorised_engine():
    """f-2: tournament standings and a ladder pass whose games run batched on the device (host arithmetic is pinned to
    the reference in tests/test_evaluation_cpu.py)."""
    from shogidrl_b200.core import ActorCritic, PPOAgent
    from shogidrl_b200.evaluation import EloTracker, evaluate_ladder, evaluate_tournament
    cfg = make_config()
    torch.manual_seed(5)
    dev = torch.device("cuda")
    agent = PPOAgent(ActorCritic(46, 13527), cfg, dev, use_mixed_precision=True)
    other = PPOAgent(ActorCritic(46, 13527), cfg, dev, use_mixed_precision=True)
    kw = dict(num_envs=32, max_moves_per_game=40, seed=3, deterministic=False)
    standings, results = evaluate_tournament(agent, {"random": None, "other": other}, 32, **kw)
    o = standings["overall_tournament_stats"]
    assert set(standings["per_opponent_results"]) == {"random", "other"}
    assert o["total_games"] == sum(r.games for r in results.values()) >= 64
    assert o["agent_total_wins"] + o["agent_total_losses"] + o["agent_total_draws"] == o["total_games"]
    for name, r in results.items():
        row = standings["per_opponent_results"][name]
        assert (row["played"], row["wins"], row["losses"], row["draws"]) == (r.games, r.agent_wins, r.opponent_wins, r.draws)
    from shogidrl_b200.evaluation import evaluate_benchmark
    perf, cases = evaluate_benchmark(agent, {"benchmark_random": None, "benchmark_other": other}, 16, **kw)
    assert set(perf["per_benchmark_case_results"]) == {"benchmark_random", "benchmark_other"}
    for name, r in cases.items():
        row = perf["per_benchmark_case_results"][name]
        assert (row["played"], row["wins_or_passes"]) == (16, r.agent_wins) and row["pass_rate"] == r.agent_wins / 16
    assert perf["overall_benchmark_pass_rate"] == sum(r.agent_wins for r in cases.values()) / 32
    tracker = EloTracker()
    tracker.ratings.update({"random": 1400.0, "other": 1500.0, "far": 2200.0})
    snap, played = evaluate_ladder(agent, "agent", {"random": None, "other": other, "far": other}, tracker,
                                   num_games_per_match=8, **kw)
    assert list(played) == ["random", "other"]  # "far" is outside the +-400 window; ascending by rating
    assert snap["far"] == 2200.0 and set(snap) == {"agent", "random", "other", "far"}
    expect = EloTracker()
    expect.ratings.update({"random": 1400.0, "other": 1500.0, "far": 2200.0})
    for name in played:
        expect.update_ratings("agent", name, played[name].outcomes)
    assert snap == expect.get_elo_snapshot()
    total = sum(snap.values())
    assert abs(total - (1500.0 + 1400.0 + 1500.0 + 2200.0)) < 1e-6  # Elo updates are zero-sum


def test_device_batch_sfen_dump_load_and_kif(golden_dir):
    """f-3: device games -> SFEN / text / KIF and back.  A batch replays golden traces on the device; its dumps
    equal the reference's strings ply by ply, and reloading the dumped SFENs reproduces the legal masks."""
    import json
    from shogidrl_b200.shogi import ShogiGame
    from shogidrl_b200.shogi.kif import game_to_kif
    from shogidrl_b200.utils import PolicyOutputMapper
    from shogidrl_b200.vec_env import VecShogiEnv
    with open(os.path.join(golden_dir, "io_golden.json"), encoding="utf-8") as f:
        games = json.load(f)["games"]
    dev = torch.device("cuda:0")
    n = len(games)
    T = min(len(g["plies"]) for g in games) - 1   # stop before any game ends: no auto-reset in the comparison
    env = VecShogiEnv(n, max_moves_per_game=500, device=dev)
    env.reset()
    for t in range(T):
        env.step(torch.tensor([g["plies"][t]["action"] for g in games], device=dev))
        if t % 9 == 0 or t == T - 1:
            assert env.to_sfen() == [g["plies"][t]["sfen"] for g in games]
            assert [p.to_string() for p in env.to_games()] == [g["plies"][t]["text"] for g in games]
    mask_before = env.mask.clone()
    env2 = VecShogiEnv(n, max_moves_per_game=500, device=dev)
    env2.load_sfens(env.to_sfen())
    assert torch.equal(env2.mask, mask_before)
    assert int(env.errors().abs().sum()) == 0
    # scalar facade: the KIF of a replayed game equals the reference's export
    mapper = PolicyOutputMapper()
    g0 = games[0]
    game = ShogiGame(max_moves_per_game=g0["max_moves"], device=dev)
    for t, ply in enumerate(g0["plies"]):
        game.make_move(mapper.policy_index_to_shogi_move(ply["action"]))
        if str(t) in g0["kif"]:
            kif = game_to_kif(game, sente_player_name="A", gote_player_name="B")
            kif = "\n".join("*Date: X" if ln.startswith("*Date:") else ln for ln in kif.split("\n"))
            assert kif == g0["kif"][str(t)]
        if t >= 45:
            break


def test_graphed_ppo_update_matches_eager():
    """PPOAgent.learn with the minibatch update replayed from a CUDA graph (default on CUDA for the fused model)
    against the same update launched eagerly: same metrics and parameters after two learn() calls."""
    from types import SimpleNamespace
    from shogidrl_b200.core import ActorCritic, PPOAgent
    dev = torch.device("cuda:0")
    B, mbs = 2048, 256

    class Buf:
        def __init__(self):
            g = torch.Generator(device="cpu").manual_seed(0)
            mask = torch.rand(B, 13536, generator=g) < 0.004
            mask[:, 0] = True
            self.mask = mask.to(dev)
            obs = torch.rand(B, 46, 9, 9, generator=g)
            self.batch = {"obs": obs.to(dev), "actions": torch.zeros(B, dtype=torch.int64, device=dev),
                          "log_probs": torch.full((B,), -3.0, device=dev), "values": torch.zeros(B, device=dev),
                          "advantages": torch.randn(B, generator=g).to(dev), "returns": torch.randn(B, generator=g).to(dev),
                          "legal_masks": self.mask[:, :13527]}

        def get_batch(self):
            return self.batch

    def run(graph):
        cfg = make_config(device="cuda", ppo_epochs=2, minibatch_size=mbs, steps_per_epoch=B)
        cfg.training.cuda_graph_update = graph
        torch.manual_seed(3)
        agent = PPOAgent(ActorCritic(46, 13527), cfg, dev, use_mixed_precision=True)
        buf = Buf()
        m1 = agent.learn(buf)
        m2 = agent.learn(buf)
        assert (agent._graph is not None) == graph
        return m1, m2, [p.detach().clone() for p in agent.model.parameters()], agent.last_gradient_norm

    ma1, ma2, pa, gna = run(True)
    mb1, mb2, pb, gnb = run(False)
    for ma, mb_ in ((ma1, mb1), (ma2, mb2)):
        for k in ma:
            assert abs(ma[k] - mb_[k]) <= 1e-3 * max(1.0, abs(mb_[k])), (k, ma[k], mb_[k])
    for x, y in zip(pa, pb):
        assert torch.allclose(x, y, rtol=1e-3, atol=1e-4), float((x - y).abs().max())
    assert abs(gna - gnb) <= 1e-3 * max(1.0, gnb)


def test_captured_update_is_dropped_when_what_it_baked_in_changes(tmp_path):
    """The captured minibatch update holds raw pointers to the optimizer's state tensors and bakes lr / clip /
    coefficients in as kernel scalars.  load_model() swaps the state tensors and a callback may anneal a coefficient:
    both must invalidate the graph, and training after a load must continue exactly like an agent that never captured."""
    from shogidrl_b200.core import ActorCritic, PPOAgent
    dev = torch.device("cuda:0")
    B, mbs = 1024, 256
    g = torch.Generator(device="cpu").manual_seed(0)
    mask = (torch.rand(B, 13536, generator=g) < 0.004)
    mask[:, 0] = True
    batch = {"obs": torch.rand(B, 46, 9, 9, generator=g).to(dev), "actions": torch.zeros(B, dtype=torch.int64, device=dev),
             "log_probs": torch.full((B,), -3.0, device=dev), "values": torch.zeros(B, device=dev),
             "advantages": torch.randn(B, generator=g).to(dev), "returns": torch.randn(B, generator=g).to(dev),
             "legal_masks": mask.to(dev)[:, :13527]}

    class Buf:
        def get_batch(self):
            return batch

    def make(graph):
        cfg = make_config(device="cuda", ppo_epochs=2, minibatch_size=mbs, steps_per_epoch=B)
        cfg.training.cuda_graph_update = graph
        torch.manual_seed(3)
        return PPOAgent(ActorCritic(46, 13527), cfg, dev, use_mixed_precision=True)

    a, b = make(True), make(False)
    for ag in (a, b):
        ag.learn(Buf()); ag.learn(Buf())
    assert a._graph is not None
    ck = str(tmp_path / "ck.pth")
    a.save_model(ck, 10, 1)
    # the file is what the reference writes: non-capturable Adam state with a CPU fp32 step counter
    sd = torch.load(ck, map_location="cpu", weights_only=False)["optimizer_state_dict"]
    assert sd["param_groups"][0]["capturable"] is False
    assert all(st["step"].device.type == "cpu" and st["step"].dtype == torch.float32 for st in sd["state"].values())
    old_state = [t.data_ptr() for st in a.optimizer.state.values() for t in st.values() if torch.is_tensor(t)]
    for ag in (a, b):
        assert "error" not in ag.load_model(ck)
    assert a._graph is None  # dropped: its kernels pointed at the replaced exp_avg / exp_avg_sq / step tensors
    new_state = [t.data_ptr() for st in a.optimizer.state.values() for t in st.values() if torch.is_tensor(t)]
    assert old_state != new_state
    ma, mb_ = a.learn(Buf()), b.learn(Buf())
    a.learn(Buf()); b.learn(Buf())
    assert a._graph is not None
    for k in ma:
        assert abs(ma[k] - mb_[k]) <= 1e-3 * max(1.0, abs(mb_[k])), (k, ma[k], mb_[k])
    for x, y in zip(a.model.parameters(), b.model.parameters()):
        assert torch.allclose(x, y, rtol=1e-3, atol=1e-4)
    # the moments the checkpoint now holds are the ones the updates since the load have advanced
    for (pa, sa), (pb, sb) in zip(a.optimizer.state.items(), b.optimizer.state.items()):
        assert float(sa["step"]) == float(sb["step"]) and torch.allclose(sa["exp_avg"], sb["exp_avg"], rtol=1e-3, atol=1e-6)
    # an annealed coefficient takes effect on the next learn() (it is a baked scalar of the captured kernels)
    a.entropy_coef = b.entropy_coef = 0.5
    key_before = a._graph_key
    ma, mb_ = a.learn(Buf()), b.learn(Buf())
    assert a._graph_key != key_before
    for x, y in zip(a.model.parameters(), b.model.parameters()):
        assert torch.allclose(x, y, rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("model_kind", ["cnn", "resnet"])
def test_fused_optimizer_tail_matches_torch_optimizer(model_kind):
    """PPOAgent.learn with the fused clip + Adam tail (kz_adam_clip_step on the optimizer's own state) against
    clip_grad_norm_ + torch.optim.Adam.step(): same metrics, gradient norm, parameters and optimizer state, and a
    save_model / load_model round trip of that state into an agent that uses the stock optimizer."""
    from shogidrl_b200.core import ActorCritic, ActorCriticResTower, PPOAgent
    dev = torch.device("cuda:0")
    B, mbs = 512, 128
    g = torch.Generator(device="cpu").manual_seed(2)
    mask = torch.rand(B, 13536, generator=g) < 0.004
    mask[:, 0] = True
    mask = mask.to(dev)
    batch = {"obs": torch.rand(B, 46, 9, 9, generator=g).to(dev), "actions": torch.zeros(B, dtype=torch.int64, device=dev),
             "log_probs": torch.full((B,), -3.0, device=dev), "values": torch.zeros(B, device=dev),
             "advantages": torch.randn(B, generator=g).to(dev), "returns": torch.randn(B, generator=g).to(dev),
             "legal_masks": mask[:, :13527]}

    class Buf:
        def get_batch(self):
            return batch

    def run(fused):
        cfg = make_config(device="cuda", ppo_epochs=2, minibatch_size=mbs, steps_per_epoch=B)
        cfg.training.fused_optimizer = fused
        cfg.training.cuda_graph_update = False
        torch.manual_seed(7)
        model = ActorCritic(46, 13527) if model_kind == "cnn" else ActorCriticResTower(46, 13527, 2, 32, 0.25)
        agent = PPOAgent(model, cfg, dev, use_mixed_precision=True)
        m = agent.learn(Buf())
        assert agent._fused_optimizer == fused
        return agent, m

    a, ma = run(True)
    b, mb_ = run(False)
    # The CNN path is deterministic up to fp32 rounding, so eight updates stay together tightly.  The ResNet's cuDNN
    # weight gradients and batch-norm statistics are not run-to-run reproducible in bf16: two runs of the SAME optimizer
    # drift by ~0.5 % in the gradient norm over eight updates, so that case only has to stay within a few percent (the
    # exact single-step comparison is tests/test_gpu_rl.py::test_fused_clip_adam_matches_torch).
    # Parameters whose true gradient is zero (a conv bias in front of batch norm) receive rounding noise, which Adam
    # normalises to +-lr per step, so ResNet parameters are only bounded by the eight steps they can have moved apart.
    tol, atol = (2e-3, 2e-4) if model_kind == "cnn" else (5e-2, 2 * 8 * 3e-4 + 1e-4)
    for k in ma:
        assert abs(ma[k] - mb_[k]) <= tol * max(1.0, abs(mb_[k])), (k, ma[k], mb_[k])
    assert abs(a.last_gradient_norm - b.last_gradient_norm) <= tol * max(1.0, b.last_gradient_norm)
    for (k, x), y in zip(a.model.named_parameters(), b.model.parameters()):
        assert torch.allclose(x.detach(), y.detach(), rtol=tol, atol=atol), (k, float((x - y).detach().abs().max()))
        sa, sb = a.optimizer.state[x], b.optimizer.state[y]
        assert float(sa["step"]) == float(sb["step"]) == 8
        if model_kind == "cnn":
            assert torch.allclose(sa["exp_avg"], sb["exp_avg"], rtol=5e-2, atol=1e-5), k
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "ck.pth")
        a.save_model(path, global_timestep=123)
        out = b.load_model(path)
        assert out["global_timestep"] == 123 and "error" not in out
        for x, y in zip(a.model.parameters(), b.model.parameters()):
            assert torch.equal(x, y) and torch.equal(a.optimizer.state[x]["exp_avg_sq"], b.optimizer.state[y]["exp_avg_sq"])
        assert all(np.isfinite(v) for v in b.learn(Buf()).values())  # the stock optimizer continues from the fused state


def test_selfplay_trainer_run_loop_and_checkpoint(tmp_path):
    """SelfPlayTrainer.run: the batched TrainingLoopManager.run -- epochs until total_timesteps, no PPO update after the
    epoch that reaches the target (training_loop_manager.py:125-133), callbacks per epoch, a reference-format checkpoint
    with the cumulative game counters, and resume from it."""
    from shogidrl_b200.core import ActorCritic
    from shogidrl_b200.training.selfplay import SelfPlayTrainer
    N, T = 64, 8
    cfg = make_config(device="cuda", minibatch_size=128, ppo_epochs=1, steps_per_epoch=N * T, total_timesteps=3 * N * T)
    cfg.env.max_moves_per_game = 12  # short games: episodes finish inside the run
    torch.manual_seed(0)
    tr = SelfPlayTrainer(ActorCritic(46, 13527), cfg, N, T, "cuda")
    seen, logs = [], []
    path = str(tmp_path / "ck.pth")
    last = tr.run(log=logs.append, callbacks=[lambda t, m: seen.append(dict(m))], checkpoint_path=path,
                  checkpoint_interval_timesteps=N * T)
    assert tr.global_timestep == 3 * N * T and tr.current_epoch == 3 and len(seen) == 3
    assert "ppo/policy_loss" in seen[0] and "ppo/policy_loss" in seen[1] and "ppo/policy_loss" not in seen[2]
    assert any("Target timesteps" in m for m in logs) and last["speed/sps"] > 0
    assert tr.driver.episodes > 0 and sum(m["episodes/episodes"] for m in seen) == tr.driver.episodes
    ck = torch.load(path, map_location="cpu", weights_only=False)
    assert set(ck) >= {"model_state_dict", "optimizer_state_dict", "global_timestep", "total_episodes_completed",
                       "black_wins", "white_wins", "draws"}
    assert ck["global_timestep"] == 3 * N * T and ck["total_episodes_completed"] == tr.driver.episodes
    assert set(ck["model_state_dict"]) == {"conv.weight", "conv.bias", "policy_head.weight", "policy_head.bias",
                                           "value_head.weight", "value_head.bias"}  # keisei/core/neural_network.py names
    tr2 = SelfPlayTrainer(ActorCritic(46, 13527), cfg, N, T, "cuda")
    out = tr2.load_checkpoint(path)
    assert "error" not in out and tr2.global_timestep == 3 * N * T and tr2.driver.episodes == tr.driver.episodes
    for x, y in zip(tr.agent.model.parameters(), tr2.agent.model.parameters()):
        assert torch.equal(x, y)
    tr2.run(total_timesteps=5 * N * T)  # continues: two more epochs, the first of them with an update
    assert tr2.global_timestep == 5 * N * T and tr2.current_epoch == 2


def test_host_pipelined_env_equals_one_batch():
    """HostPipelinedEnv (2 and 4 groups, host actions, own streams) plays exactly the games a single VecShogiEnv
    plays with the same seed: group g's games carry the RNG streams of envs [g n/G, (g+1) n/G)."""
    from shogidrl_b200.host_env import HostPipelinedEnv
    from shogidrl_b200.vec_env import VecShogiEnv
    dev = torch.device("cuda:0")
    n, T = 512, 60
    ref = VecShogiEnv(n, max_moves_per_game=40, device=dev, seed=77, auto_reset=True)
    ref.refresh(random_actions=True)
    acts = [ref.next_actions.clone(), torch.empty_like(ref.next_actions)]
    rewards = []
    for t in range(T):
        out = ref.step(acts[t & 1], random_actions=True, next_out=acts[(t + 1) & 1])
        rewards.append(out["reward"].clone())
    rb, rh, rm = [x.cpu() for x in ref.export()]
    for G in (2, 4):
        pipe = HostPipelinedEnv(n, groups=G, max_moves_per_game=40, device=dev, seed=77, auto_reset=True)
        pipe.prime(random_actions=True)
        got = [[] for _ in range(G)]
        for t in range(T):
            for g in range(G):
                nxt, rew, done, reason, winner = pipe.wait(g)
                if t > 0:
                    got[g].append(rew.clone())
                pipe.h_actions[g].copy_(nxt)          # the host "policy": play the pre-selected legal action
                pipe.submit(g, random_actions=True)
        for g in range(G):
            got[g].append(pipe.wait(g)[1].clone())
        pipe.synchronize()
        b = torch.cat([e.export()[0].cpu() for e in pipe.envs]); h = torch.cat([e.export()[1].cpu() for e in pipe.envs])
        m = torch.cat([e.export()[2].cpu() for e in pipe.envs])
        assert torch.equal(b, rb) and torch.equal(h, rh) and torch.equal(m[:, :5], rm[:, :5])
        for t in range(T):
            assert torch.equal(torch.cat([got[g][t] for g in range(G)]), rewards[t].cpu())
        assert all(int(e.errors().abs().sum()) == 0 for e in pipe.envs)


def test_rule_queries_match_oracle_on_random_endgames():
    """The facade's rule queries, which the hot loop never calls but displays and tests do -- generate_piece_potential_moves
    (kz_piece_targets), ShogiGame.is_uchi_fu_zume, can_drop_piece, get_king_legal_moves (shogi_game.py:208-260;
    shogi_rules_logic.py:82-208, 275-359, 424-483) -- against the oracle's restatement of the same reference functions:
    pseudo-legal targets of EVERY piece of 4,096 random drop-heavy endgames in batch, and the three scalar queries on 256
    of them (every pawn-drop square in front of a king, sampled drops of every type in hand, both kings)."""
    from tests.helpers import random_endgames
    from shogidrl_b200 import VecShogiEnv
    from shogidrl_b200.shogi import Color, Piece, PieceType, ShogiGame

    dev = torch.device("cuda:0")
    n = 4096
    boards, hands, sides, mcs = random_endgames(n, 2024)
    games = [orc.OracleGame.from_arrays(boards[i], hands[i], int(sides[i]), 0, 500, evaluate_termination=False) for i in range(n)]
    env = VecShogiEnv(n, device=dev, auto_reset=False)
    env.load_positions(boards, hands, sides, mcs, eval_termination=False)
    checked = 0
    for sq in range(81):
        occ = np.nonzero(boards[:, sq])[0]
        if len(occ) == 0:
            continue
        w = env.piece_targets(np.full(n, sq, np.int32)).cpu().numpy().astype(np.uint32)
        got = ((w[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(n, 96)[:, :81].astype(np.uint8)
        for i in occ:
            want = games[i].piece_targets(sq)
            assert np.array_equal(got[i], want), (i, sq, int(boards[i, sq]))
            checked += 1
    assert checked > 20000
    # scalar facade queries
    rng = np.random.default_rng(5)
    game = ShogiGame()
    ufz_true = drops = 0
    for i in range(256):
        game.board = [[Piece.from_code(int(boards[i, r * 9 + c])) for c in range(9)] for r in range(9)]
        for color in (0, 1):
            for t in range(7):
                game.hands[color][PieceType(t)] = int(hands[i, color * 7 + t])
        game.current_player = Color(int(sides[i]))
        game.move_count = 0
        game.game_over, game.winner, game.termination_reason = False, None, None
        o = games[i]
        for color in (Color.BLACK, Color.WHITE):
            assert game.get_king_legal_moves(color) == o.king_legal_moves(color.value), (i, color)
            # the square in front of the enemy king (where a pawn drop can give check) + two random squares
            ek = game.find_king(color.opponent())
            cand = [int(x) for x in rng.integers(0, 81, 2)]
            if ek is not None:
                fr = ek[0] + (1 if color == Color.BLACK else -1)
                if 0 <= fr < 9:
                    cand.append(fr * 9 + ek[1])
            for sq in cand:
                want = o.is_uchi_fu_zume(sq, color.value)
                assert game.is_uchi_fu_zume(sq // 9, sq % 9, color) == want, (i, color, sq)
                ufz_true += int(want)
            for t in range(7):
                for sq in [int(x) for x in rng.integers(0, 81, 2)] + cand[-1:]:
                    want = o.can_drop(t, sq, color.value)
                    assert game.can_drop_piece(PieceType(t), sq // 9, sq % 9, color) == want, (i, color, t, sq)
                    drops += int(want)
    assert drops > 500


def test_no_legal_action_rows_fall_back_to_uniform_and_say_so(capsys):
    """A legal mask without any legal action: the reference masks every logit to -inf, gets NaN probabilities, reports
    them on stderr and continues with a uniform distribution (base_actor_critic.py:92-101, 166-174; ppo_agent.py:160-168).
    The kernels take the same fallback; the facade prints the same lines."""
    import math
    from shogidrl_b200.core import ActorCritic, PPOAgent
    cfg = make_config()
    torch.manual_seed(3)
    agent = PPOAgent(ActorCritic(46, 13527), cfg, torch.device("cuda"))
    obs = np.zeros((46, 9, 9), np.float32)
    none_legal = torch.zeros(13527, dtype=torch.bool, device="cuda")
    move, idx, lp, v = agent.select_action(obs, none_legal, is_training=True)
    err = capsys.readouterr().err
    assert "[PPOAgent] ERROR: select_action called with no legal moves (based on input legal_mask)" in err
    assert "[ActorCritic] ERROR: NaNs in probabilities in get_action_and_value. Check legal_mask and logits. Defaulting to uniform." in err
    assert 0 <= idx < 13527 and abs(lp - math.log(1.0 / 13527)) < 1e-4 and np.isfinite(v) and move is not None
    masks = torch.zeros((3, 13527), dtype=torch.bool, device="cuda")
    masks[1, 100:140] = True
    acts = torch.tensor([5, 120, 13000], device="cuda")
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logp, ent, val = (t.detach() for t in agent.model.evaluate_actions(torch.zeros((3, 46, 9, 9), device="cuda"), acts, masks))
    err = capsys.readouterr().err
    assert "[ActorCritic] ERROR: NaNs in probabilities in evaluate_actions. Check legal_mask and logits. Defaulting to uniform for affected rows." in err
    assert torch.isfinite(logp).all() and torch.isfinite(ent).all() and torch.isfinite(val).all()
    for r in (0, 2):
        assert abs(float(logp[r]) - math.log(1.0 / 13527)) < 1e-4 and abs(float(ent[r]) - math.log(13527)) < 1e-3
    assert float(ent[1]) <= math.log(40) + 1e-4
    agent.model.evaluate_actions(torch.zeros((1, 46, 9, 9), device="cuda"), acts[1:2], masks[1:2])
    assert "NaNs" not in capsys.readouterr().err

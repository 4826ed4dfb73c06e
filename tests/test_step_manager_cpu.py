"""StepManager / EnvManager / buffer host logic with injected mocks, the way the reference tests its StepManager
(tests/training/test_step_manager.py:37-62): CPU only, no engine calls."""
from unittest.mock import MagicMock, Mock

import numpy as np
import pytest
import torch

from shogidrl_b200.core.experience_buffer import Experience, ExperienceBuffer
from shogidrl_b200.shogi.definitions import Color, PieceType
from shogidrl_b200.training.step_manager import EpisodeState, StepManager, StepResult
from tests.helpers import make_config


def _mk(move=(6, 0, 5, 0, False), done=False, info=None):
    cfg = make_config(device="cpu")
    game = Mock()
    game.current_player = Color.BLACK
    game.get_legal_moves.return_value = [move]
    game.make_move.return_value = (np.ones((46, 9, 9), np.float32), 0.5, done, info or {"reason": "Game ongoing"})
    game.reset.return_value = np.zeros((46, 9, 9), np.float32)
    agent = Mock()
    agent.select_action.return_value = (move, 7, -0.25, 0.125)
    mapper = Mock()
    mapper.get_legal_mask.return_value = torch.zeros(13527, dtype=torch.bool)
    buf = Mock()
    sm = StepManager(cfg, game, agent, mapper, buf)
    obs = np.zeros((46, 9, 9), np.float32)
    st = EpisodeState(obs, torch.zeros(1, 46, 9, 9), 0.0, 0)
    return sm, game, agent, mapper, buf, st


def test_execute_step_call_order_and_buffer_add():
    sm, game, agent, mapper, buf, st = _mk()
    res = sm.execute_step(st, 3, Mock())
    assert res.success and res.policy_index == 7 and res.reward == 0.5 and not res.done
    game.get_legal_moves.assert_called_once()
    mapper.get_legal_mask.assert_called_once()
    assert mapper.get_legal_mask.call_args[0][0] == [(6, 0, 5, 0, False)]
    agent.select_action.assert_called_once()
    assert agent.select_action.call_args[1] == {"is_training": True}
    game.make_move.assert_called_once_with((6, 0, 5, 0, False))
    args = buf.add.call_args[0]
    assert args[1:6] == (7, 0.5, -0.25, 0.125, False) and args[0].shape == (46, 9, 9)
    assert res.next_obs_tensor.shape == (1, 46, 9, 9)


def test_no_legal_moves_resets_and_reports_terminal():
    sm, game, *_, st = _mk()
    game.get_legal_moves.return_value = []
    log = Mock()
    res = sm.execute_step(st, 0, log)
    assert not res.success and res.done and res.info == {"terminal_reason": "no_legal_moves"}
    game.reset.assert_called_once()
    assert "TERMINAL" in log.call_args[0][0]


def test_agent_returns_none_and_value_error_paths():
    sm, game, agent, *_ , st = _mk()
    agent.select_action.return_value = (None, -1, 0.0, 0.0)
    res = sm.execute_step(st, 0, Mock())
    assert not res.success and not res.done and "failed to select" in res.error_message
    sm, game, agent, mapper, buf, st = _mk()
    game.make_move.side_effect = ValueError("Illegal movement pattern: boom")
    res = sm.execute_step(st, 0, Mock())
    assert not res.success and "Illegal movement pattern" in res.error_message
    game.reset.assert_called_once()
    buf.add.assert_not_called()
    game.reset.side_effect = RuntimeError("reset broke")
    res = sm.execute_step(st, 0, Mock())
    assert res.done and "Reset also failed" in res.error_message


def test_counters_and_episode_end():
    sm, game, agent, mapper, buf, st = _mk(move=(None, None, 4, 4, PieceType.PAWN), done=True,
                                           info={"reason": "Tsumi", "winner": "BLACK", "captured_piece_type": "PROMOTED_ROOK"})
    res = sm.execute_step(st, 0, Mock())
    assert sm.sente_drop_count == 1 and sm.sente_capture_count == 1 and sm.sente_best_capture == "Rook"
    log = Mock()
    new_state, winner = sm.handle_episode_end(sm.update_episode_state(st, res), res, {"black_wins": 1, "white_wins": 0, "draws": 1}, 2, log)
    assert winner == "black" and new_state.episode_length == 0 and sm.sente_drop_count == 0
    kw = log.call_args[1]
    assert kw["wandb_data"]["black_wins_total"] == 2 and abs(kw["wandb_data"]["black_win_rate"] - 2 / 3) < 1e-9
    assert "Sente wins by Tsumi." in log.call_args[0][0]
    st2 = sm.update_episode_state(st, res)
    assert st2.episode_length == 1 and st2.episode_reward == 0.5


def test_experience_buffer_contract_cpu(capsys):
    buf = ExperienceBuffer(3, 0.99, 0.95, "cpu")
    assert buf.obs.shape == (3, 46, 9, 9) and buf.legal_masks.dtype == torch.bool and buf.legal_masks.is_contiguous()
    assert buf.actions.dtype == torch.int64 and buf.dones.dtype == torch.bool and buf.get_batch() == {}
    o, m = torch.zeros(46, 9, 9), torch.zeros(13527, dtype=torch.bool)
    for i in range(3):
        buf.add(o, i, 1.0, 0.0, 0.5, False, m)
    with pytest.raises(RuntimeError, match="compute_advantages_and_returns"):
        buf.get_batch()
    buf.add(o, 9, 1.0, 0.0, 0.5, False, m)  # full: dropped with a warning, no raise
    assert "Buffer is full. Cannot add new experience." in capsys.readouterr().err and len(buf) == 3
    with pytest.raises(RuntimeError, match="The expanded size of the tensor"):
        ExperienceBuffer(1, 0.99, 0.95).add(torch.zeros(3, 9, 9), 0, 0.0, 0.0, 0.0, False, m)
    other = ExperienceBuffer(8, 0.99, 0.95)
    other.merge_from_parallel_buffers([buf, buf])
    assert other.size() == 6 and other.capacity() == 8 and other.actions[:6].tolist() == [0, 1, 2, 0, 1, 2]
    other.add_batch([Experience(o, 5, 0.0, 0.0, 0.0, True, m)] * 5)
    assert other.size() == 8
    buf.clear()
    assert len(buf) == 0
    from shogidrl_b200 import NativeError
    buf.add(o, 0, 1.0, 0.0, 0.5, False, m)
    if not torch.cuda.is_available():
        with pytest.raises(NativeError):  # GAE has no CPU fallback
            buf.compute_advantages_and_returns(0.0)


def test_env_manager_action_space_mismatch():
    from shogidrl_b200.training.env_manager import EnvManager
    cfg = make_config(device="cuda")
    cfg.env.num_actions_total = 13000
    em = EnvManager(cfg, Mock())
    with pytest.raises(RuntimeError, match="Action space mismatch"):
        em.setup_environment()
    cfg.env.num_actions_total = 13527
    game, mapper = EnvManager(cfg).setup_environment()  # constructing the facade does not touch the GPU
    assert mapper.get_total_actions() == 13527 and game.max_moves_per_game == 500
    assert game.to_sfen_string() == "lnsgkgsnl/1r5b1/ppppppppp/9/9/9/PPPPPPPPP/1B5R1/LNSGKGSNL b - 1"


def test_distributed_sharding_math():
    from shogidrl_b200.training.distributed import shard_envs
    for total, world in [(65536, 8), (10, 3), (7, 8)]:
        spans = [shard_envs(total, r, world) for r in range(world)]
        assert sum(n for _, n in spans) == total
        assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(world - 1)) and spans[0][0] == 0

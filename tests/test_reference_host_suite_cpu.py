"""The reference's own mock-driven tests of the two training-side callers of the hot path -- tests/training/
test_step_manager.py (35 tests: execute_step, episode end, demo-mode log lines, winner / reason strings) and
test_env_manager.py (19 tests: setup, validation, seeding, their log messages) -- and the two CPU-only files of its
tests/shogi/ (test_move_formatting.py, 48 tests of the demo-log move descriptions; test_shogi_core_definitions.py, 34 tests of
Color / PieceType / Piece / the observation-plane constants) run UNMODIFIED against
shogidrl_b200.training.StepManager / EnvManager, shogidrl_b200.utils.move_formatting and shogidrl_b200.shogi.definitions through the hybrid `keisei` alias (tests/ref_alias: hot-path modules are
this repository's, config_schema and the rest the reference install under baseline/_ref).  They mock the game and the
agent, so they need no GPU.  Skipped where the untracked copies (tests/fetch_reference_tests.py) or baseline/_ref are absent."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_TESTS = os.path.join(ROOT, "baseline", "_ref_tests", "host")
REF_PKG = os.path.join(ROOT, "baseline", "_ref", "keisei")


@pytest.mark.skipif(not (os.path.isdir(HOST_TESTS) and os.path.isdir(REF_PKG)),
                    reason="baseline/_ref_tests/host or baseline/_ref absent (tests/fetch_reference_tests.py, DESIGN.md section 7)")
def test_reference_step_and_env_manager_tests_pass_against_this_repository():
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "tests", "ref_alias"), ROOT]),
               PYTHONDONTWRITEBYTECODE="1", WANDB_DISABLED="true", WANDB_MODE="disabled")
    out = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "--rootdir", HOST_TESTS, HOST_TESTS,
                          "--tb=short"], capture_output=True, text=True, env=env, cwd=HOST_TESTS, timeout=900)
    failed = sorted(set(re.findall(r"^(?:FAILED|ERROR) (\S+)", out.stdout, flags=re.M)))
    m = re.search(r"(\d+) passed", out.stdout)
    assert not failed and out.returncode == 0, out.stdout[-6000:] + out.stderr[-2000:]
    assert m and int(m.group(1)) >= 136, out.stdout[-2000:]
    # the classes under test were this repository's, not the reference's
    probe = subprocess.run([sys.executable, "-c", "import keisei.training.step_manager as s, keisei.training.env_manager as e, keisei.utils as u,"
                            "keisei.shogi.shogi_core_definitions as d;"
                            "print(s.StepManager.__module__, e.EnvManager.__module__, u.format_move_with_description.__module__,"
                            "d.PieceType.__module__)"], capture_output=True, text=True, env=env)
    assert probe.stdout.split() == ["shogidrl_b200.training.step_manager", "shogidrl_b200.training.env_manager",
                                    "shogidrl_b200.utils.move_formatting", "shogidrl_b200.shogi.definitions"], probe

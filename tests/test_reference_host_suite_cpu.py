"""The reference's own mock-driven tests of the two training-side callers of the hot path -- tests/training/
test_step_manager.py (35 tests: execute_step, episode end, demo-mode log lines, winner / reason strings) and
test_env_manager.py (19 tests: setup, validation, seeding, their log messages) -- run UNMODIFIED against
shogidrl_b200.training.StepManager / EnvManager through the hybrid `keisei` alias (tests/ref_alias: hot-path modules are
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
    assert m and int(m.group(1)) >= 54, out.stdout[-2000:]
    # the classes under test were this repository's, not the reference's
    probe = subprocess.run([sys.executable, "-c", "import keisei.training.step_manager as s, keisei.training.env_manager as e;"
                            "print(s.StepManager.__module__, e.EnvManager.__module__)"], capture_output=True, text=True, env=env)
    assert probe.stdout.split() == ["shogidrl_b200.training.step_manager", "shogidrl_b200.training.env_manager"], probe

"""Build-container only (needs /root/reference): the restated third-party pieces the oracle
depends on (oracle/emcee_port.py, oracle/skopt_port.py) are pinned by running the UNMODIFIED
reference sources, and the reference's own golden tests, on top of them.

The RNG-stream-dependent golden values of the reference (8 acquisition argmaxes in
tests/test_acquisition.py:42-70) only reproduce if the emcee restatement consumes the numpy
RandomState exactly like emcee 3.1.6 does, so this is a sharp check of oracle/emcee_port.py."""
import os
import subprocess
import sys

import pytest

from oracle.ref_loader import reference_available

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference is not present")


def run_reference_tests(files, extra=()):
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1", PYTHONPATH=REPO)
    cmd = [sys.executable, "-m", "pytest", "-q", "-p", "oracle.ref_pytest_plugin", "-p", "no:cacheprovider",
           "--rootdir", "/tmp", *extra, *[os.path.join("/root/reference/tests", f) for f in files]]
    return subprocess.run(cmd, cwd=REPO, env=env, capture_output=True, text=True)


def test_reference_golden_acquisition_and_gpr_tests_pass_on_the_restated_stack():
    res = run_reference_tests(["test_acquisition.py", "test_bayesgpr.py", "test_utils.py", "test_priors.py"])
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
    assert "18 passed" in res.stdout, res.stdout[-500:]


def test_reference_optimizer_tests_pass_on_the_restated_stack():
    res = run_reference_tests(["test_optimizer.py"],
                              extra=("-k", "not optimum_intervals and not optimality and not expected_optimality"))
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]

"""Runs the C++ drop-in test of corbo::SolverB200Lm (tests/adapter/dropin_test.cpp) -- the reference's own
PredictiveController / StructuredOptimalControlProblem stack with the plugin solver against the same stack with
LevenbergMarquardtSparse.  The binary links the compiled reference, so it is built where /root/reference exists
(make -C tests/adapter) and travels to the GPU box with the snapshot."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

BIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "adapter", "_build", "dropin_test")


def test_dropin_under_reference_controller_stack():
    if not os.path.exists(BIN):
        pytest.skip("tests/adapter/_build/dropin_test not built (needs /root/reference at build time)")
    out = subprocess.run([BIN], capture_output=True, text=True, timeout=600)
    print(out.stdout[-3000:])
    print(out.stderr[-2000:])
    assert out.returncode == 0 and "DROP-IN TEST PASSED" in out.stdout

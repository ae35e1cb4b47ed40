"""Row a15 of SURVEY.md section 8: numerical linearisation of the system dynamics (SystemDynamicsInterface::getLinearA/getLinearB over
numerics ForwardDifferences / CentralDifferences).  The fixture tests/golden/linearize.npz holds the compiled reference's own
matrices; the oracle is pinned against it on CPU, the device path (b200sqp_linearize_dynamics) against both on the GPU."""
import os
import sys

import numpy as np
import pytest

from control_box_rst_b200 import solver

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import cases  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "linearize.npz"))


def _tol(ref):
    # one ulp of a transcendental amplified by 1/delta = 1e9 (forward) on entries of size ~10
    return 4e-6 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("method", ["forward", "central"])
@pytest.mark.parametrize("name", list(cases.LINEARIZE_MODELS))
def test_oracle_matches_reference_fixture(oracle, name, method):
    make, polynomial = cases.LINEARIZE_MODELS[name]
    ocp = make()
    xs, us = GOLD[f"{name}_x"], GOLD[f"{name}_u"]
    assert np.array_equal(xs, cases.linearize_points(ocp)[0])
    for i in range(len(xs)):
        A, B = oracle.linearize(ocp, xs[i], us[i], method)
        A_ref, B_ref = GOLD[f"{name}_{method}_A"][i], GOLD[f"{name}_{method}_B"][i]
        if polynomial:
            assert np.array_equal(A, A_ref) and np.array_equal(B, B_ref)
        else:
            np.testing.assert_allclose(A, A_ref, rtol=0, atol=_tol(A_ref))
            np.testing.assert_allclose(B, B_ref, rtol=0, atol=_tol(B_ref))


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["forward", "central"])
@pytest.mark.parametrize("name", list(cases.LINEARIZE_MODELS))
def test_device_matches_reference_fixture_and_oracle(oracle, name, method):
    make, polynomial = cases.LINEARIZE_MODELS[name]
    ocp = make()
    xs, us = GOLD[f"{name}_x"], GOLD[f"{name}_u"]
    A, B = solver.linearize_dynamics(ocp.dynamics, list(ocp.dyn_params), xs, us, method)
    A_ref, B_ref = GOLD[f"{name}_{method}_A"], GOLD[f"{name}_{method}_B"]
    if polynomial:
        assert np.array_equal(A, A_ref) and np.array_equal(B, B_ref)  # bit-exact: same IEEE expressions, no FMA contraction
    else:
        np.testing.assert_allclose(A, A_ref, rtol=0, atol=_tol(A_ref))
        np.testing.assert_allclose(B, B_ref, rtol=0, atol=_tol(B_ref))
    for i in range(len(xs)):
        Ao, Bo = oracle.linearize(ocp, xs[i], us[i], method)
        np.testing.assert_allclose(A[i], Ao, rtol=0, atol=0 if polynomial else _tol(Ao))
        np.testing.assert_allclose(B[i], Bo, rtol=0, atol=0 if polynomial else _tol(Bo))


@pytest.mark.gpu
def test_device_linearisation_large_batch_is_pointwise():
    """size-independent property at a large batch: every point is linearised independently of its neighbours"""
    make, _ = cases.LINEARIZE_MODELS["van_der_pol"]
    ocp = make()
    rng = np.random.default_rng(5)
    xs, us = rng.uniform(-2, 2, (65536, 2)), rng.uniform(-1, 1, (65536, 1))
    A, B = solver.linearize_dynamics(ocp.dynamics, list(ocp.dyn_params), xs, us, "central")
    idx = rng.integers(0, 65536, 64)
    A2, B2 = solver.linearize_dynamics(ocp.dynamics, list(ocp.dyn_params), xs[idx], us[idx], "central")
    assert np.array_equal(A[idx], A2) and np.array_equal(B[idx], B2)
    # analytic Jacobian of the Van der Pol oscillator within the FD noise
    a = ocp.dyn_params[0]
    exact = np.zeros_like(A)
    exact[:, 0, 1] = 1.0
    exact[:, 1, 0] = -2 * a * xs[:, 0] * xs[:, 1] - 1.0
    exact[:, 1, 1] = -a * (xs[:, 0] ** 2 - 1.0)
    assert np.abs(A - exact).max() <= 1e-5 and np.abs(B[:, 1, 0] - 1.0).max() <= 1e-6


# ---- the finite-difference Hessians of numerics/finite_differences (delta = 1e-5), applied to the dynamics as a function of [x; u] ----
def _htol(ref):
    # one ulp of a transcendental (1e-16 on values of size ~10) amplified by 1/delta^2 = 1e10, a few evaluations per entry
    return 4e-5 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("weighted", [False, True], ids=["plain_sum", "multipliers"])
@pytest.mark.parametrize("method", ["forward", "central"])
@pytest.mark.parametrize("name", list(cases.LINEARIZE_MODELS))
def test_oracle_hessian_matches_reference_fixture(oracle, name, method, weighted):
    make, polynomial = cases.LINEARIZE_MODELS[name]
    ocp = make()
    xs, us = GOLD[f"{name}_x"], GOLD[f"{name}_u"]
    mult = cases.hessian_multipliers(ocp)
    ref = GOLD[f"{name}_{method}_{'Hm' if weighted else 'H'}"]
    for i in range(len(xs)):
        H = oracle.dynamics_hessian(ocp, xs[i], us[i], mult[i] if weighted else None, method)
        if polynomial:
            assert np.array_equal(H, ref[i])
        else:
            np.testing.assert_allclose(H, ref[i], rtol=0, atol=_htol(ref[i]))


@pytest.mark.gpu
@pytest.mark.parametrize("weighted", [False, True], ids=["plain_sum", "multipliers"])
@pytest.mark.parametrize("method", ["forward", "central"])
@pytest.mark.parametrize("name", list(cases.LINEARIZE_MODELS))
def test_device_hessian_matches_reference_fixture_and_oracle(oracle, name, method, weighted):
    make, polynomial = cases.LINEARIZE_MODELS[name]
    ocp = make()
    xs, us = GOLD[f"{name}_x"], GOLD[f"{name}_u"]
    mult = cases.hessian_multipliers(ocp) if weighted else None
    H = solver.dynamics_hessian(ocp.dynamics, list(ocp.dyn_params), xs, us, mult, method)
    ref = GOLD[f"{name}_{method}_{'Hm' if weighted else 'H'}"]
    if polynomial:
        assert np.array_equal(H, ref)  # bit-exact: same IEEE expressions and increment order, no FMA contraction
    else:
        np.testing.assert_allclose(H, ref, rtol=0, atol=_htol(ref))
    for i in range(len(xs)):
        Ho = oracle.dynamics_hessian(ocp, xs[i], us[i], None if mult is None else mult[i], method)
        np.testing.assert_allclose(H[i], Ho, rtol=0, atol=0 if polynomial else _htol(Ho))


@pytest.mark.gpu
def test_device_hessian_large_batch_is_pointwise_and_close_to_analytic():
    """size-independent properties at a large batch: points are independent of their neighbours; the Van der Pol Hessian
    sum_v m_v d2 f_v / dz dz has the closed form m_1 * [[-2 a x1, -2 a x0, 0], [-2 a x0, 0, 0], [0, 0, 0]]"""
    make, _ = cases.LINEARIZE_MODELS["van_der_pol"]
    ocp = make()
    rng = np.random.default_rng(6)
    n = 16384
    xs, us, m = rng.uniform(-2, 2, (n, 2)), rng.uniform(-1, 1, (n, 1)), rng.uniform(-2, 2, (n, 2))
    H = solver.dynamics_hessian(ocp.dynamics, list(ocp.dyn_params), xs, us, m, "central")
    idx = rng.integers(0, n, 64)
    H2 = solver.dynamics_hessian(ocp.dynamics, list(ocp.dyn_params), xs[idx], us[idx], m[idx], "central")
    assert np.array_equal(H[idx], H2)
    a = ocp.dyn_params[0]
    exact = np.zeros_like(H)
    exact[:, 0, 0] = -2 * a * xs[:, 1] * m[:, 1]
    exact[:, 0, 1] = exact[:, 1, 0] = -2 * a * xs[:, 0] * m[:, 1]
    assert np.abs(H - exact).max() <= 1e-3  # rounding noise of delta = 1e-5 second differences (1.1e-4 observed over 16384 points)

"""Live cross-check of the oracle against the compiled reference on fresh random instances (skipped where oracle/_ref was not
built, e.g. on the GPU box if /root/reference never existed)."""
import numpy as np
import pytest

from control_box_rst_b200 import _abi as abi
from control_box_rst_b200 import problems


@pytest.mark.parametrize("make,weights,tol", [
    (lambda: problems.van_der_pol(50), (2.0, 2.0, 2.0), 1e-5),
    (lambda: problems.van_der_pol(13, collocation=abi.COLL_MIDPOINT), (2.0, 5.0, 7.0), 1e-5),
    (lambda: problems.unicycle_time_optimal(20), (2.0, 2.0, 2.0), 1e-3),
    (lambda: problems.cart_pole_shooting(25), (10.0, 10.0, 10.0), 5e-3),
    (lambda: problems.free_space_rocket_time_optimal(16), (2.0, 3.0, 4.0), 1e-4),
    (lambda: problems.make_ocp(grid=abi.GRID_MULTIPLE_SHOOTING, dynamics=abi.DYN_TOY_EXAMPLE, n_grid=15, dt=0.05, q=(1, 1), r=(0.1,), qf=(1, 1),
                               u_lb=(-2.0,), u_ub=(2.0,), dyn_params=(0.3,), integrator=abi.INT_EULER), (2.0, 2.0, 2.0), 1e-4),
], ids=["vdp50", "vdp13_midpoint", "unicycle20", "cartpole25", "rocket16_timeopt", "toy15_ms_euler"])
def test_oracle_matches_compiled_reference(oracle, reference, make, weights, tol):
    ocp = make()
    B = 24
    x0, xref = problems.instance_data(ocp, B, seed=77)
    # Jacobian level: bit-exact for the polynomial model, last-ulp libm differences amplified by 1/(2 delta) otherwise
    v_r, J_r, P_r, a_r = reference.evaluate(ocp, x0[0], xref[0], None, weights)
    v_o, J_o, P_o, a_o = oracle.evaluate(ocp, x0[0], xref[0], None, weights)
    assert np.array_equal(P_r, P_o)
    if ocp.dynamics in (abi.DYN_VAN_DER_POL, abi.DYN_FREE_SPACE_ROCKET, abi.DYN_TOY_EXAMPLE):  # + - * / only
        assert np.array_equal(v_r, v_o) and np.array_equal(J_r, J_o) and np.array_equal(a_r, a_o)
    else:
        np.testing.assert_allclose(v_o, v_r, rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(J_o, J_r, rtol=0, atol=2e-6 * max(1.0, np.abs(J_r).max()))
    opts = abi.LmOptions.defaults(iterations=10, weights=weights)
    p_r, c_r, s_r, _ = reference.solve_batch(ocp, opts, x0, xref, threads=4)
    p_o, c_o, s_o, _ = oracle.solve_batch(ocp, opts, x0, xref, threads=4)
    err = np.abs(p_o - p_r).max(axis=1) / np.maximum(1.0, np.abs(p_r).max(axis=1))
    assert err.max() <= tol, err
    assert np.array_equal(s_r, s_o)


def test_weight_adaptation_and_warm_start_sequence(oracle, reference):
    """new_run = false adapts the penalty weights (levenberg_marquardt_sparse.cpp:83-86,270-287): compare a warm-started
    sequence of solves through the traces of both checkers (single instance)."""
    ocp = problems.van_der_pol(20)
    opts = abi.LmOptions.defaults(iterations=3, weights=(2.0, 2.0, 2.0), factors=(2.0, 2.0, 2.0), maxima=(10.0, 10.0, 10.0))
    x0 = np.array([1.0, 0.5])
    tr_r = reference.trace(ocp, opts, x0)
    tr_o = oracle.trace(ocp, opts, x0)
    assert [e[0] for e in tr_r["events"]] == [e[0] for e in tr_o["events"]]
    np.testing.assert_allclose(tr_o["params"], tr_r["params"], rtol=0, atol=1e-6)

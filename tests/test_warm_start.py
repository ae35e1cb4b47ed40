"""the moving-horizon warm start between MPC steps (FullDiscretizationGridBase::findNearestState +
warmStartShifting, full_discretization_grid_base.cpp:230-318).  tests/golden/warm_start_shift.npz holds the compiled reference's
shifted parameter vectors and a 15-step closed loop with the warm start active; the numpy restatement (oracle/warm_start.py) is pinned
against it on CPU, the device path (b200sqp_warm_start_shift) against both on the GPU."""
import os

import numpy as np
import pytest

from control_box_rst_b200 import problems, solver
from oracle import warm_start

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "warm_start_shift.npz"))


def _split(ocp, p, x0):
    """reference parameter vector -> x_seq [N, nx] (incl. start and xf), u_seq [N-1, nu]"""
    x_idx, u_idx, _ = solver.vertex_indices(ocp)
    N = ocp.n_grid
    x = np.zeros((N, ocp.nx))
    x[0] = x0
    for k in range(1, N):
        x[k] = p[x_idx[k]:x_idx[k] + ocp.nx]
    u = np.stack([p[u_idx[k]:u_idx[k] + ocp.nu] for k in range(N - 1)])
    return x, u


def _join(ocp, x, u):
    x_idx, u_idx, _ = solver.vertex_indices(ocp)
    p = np.zeros(solver.dims_of(ocp).n_params)
    for k in range(1, ocp.n_grid):
        p[x_idx[k]:x_idx[k] + ocp.nx] = x[k]
    for k in range(ocp.n_grid - 1):
        p[u_idx[k]:u_idx[k] + ocp.nu] = u[k]
    return p


# FiniteDifferencesGrid (FullDiscretizationGridBase) and MultipleShootingGrid (ShootingGridBase::findNearestShootingInterval +
# warmStartShifting, shooting_grid_base.cpp:292-381): fixture key prefix, descriptor
GRIDS = [("", lambda n: problems.van_der_pol(n)), ("ms_", lambda n: problems.van_der_pol_shooting(n))]


@pytest.mark.parametrize("prefix,make", GRIDS, ids=["fd_grid", "shooting_grid"])
def test_restatement_matches_reference_fixture(prefix, make):
    ocp = make(12)
    shifts = []
    for xo, xn, p_in, p_out in zip(GOLD[prefix + "x0_old"], GOLD[prefix + "x0_new"], GOLD[prefix + "p_in"], GOLD[prefix + "p_out"]):
        x, u = _split(ocp, p_in, xo)
        xs, us, s = warm_start.warm_start_shift(x, u, xn)
        shifts.append(s)
        assert np.array_equal(_join(ocp, xs, us), p_out)  # pure copies and a + 2 (b - a): bit-exact
    assert set(shifts) >= {0, 1, 2, 3}, shifts  # the fixture exercises no shift, single and multiple shifts


@pytest.mark.gpu
@pytest.mark.parametrize("prefix,make", GRIDS, ids=["fd_grid", "shooting_grid"])
def test_device_shift_matches_reference_fixture(prefix, make):
    ocp = make(12)
    x0_old, x0_new, p_in, p_out = (GOLD[prefix + k] for k in ("x0_old", "x0_new", "p_in", "p_out"))
    B = len(x0_old)
    lm = solver.BatchedLevenbergMarquardt(ocp, B)
    lm.set_problem_data(x0_old, None)
    lm.initialize_trajectories()
    lm.set_params(p_in)
    shifts = lm.warm_start_shift(x0_new)
    assert np.array_equal(lm.get_params(), p_out)
    want = [warm_start.find_nearest_state(_split(ocp, p, xo)[0], xn) for xo, xn, p in zip(x0_old, x0_new, p_in)]
    assert np.array_equal(shifts, np.array(want, np.int32))
    lm.clear()


@pytest.mark.gpu
def test_shooting_grid_closed_loop_with_shift_matches_reference_controller_and_plant():
    """MultipleShootingGrid with setWarmStart(true) under the reference's PredictiveController + SimulatedPlant (RK4) against one
    b200sqp_closed_loop call in shift mode."""
    ocp = problems.van_der_pol_shooting(20)
    lm = solver.BatchedLevenbergMarquardt(ocp, 1)
    lm.setIterations(10)
    u, x, chi2, status = lm.closed_loop(np.array([[1.0, 0.5]]), 15, mode=solver.BatchedLevenbergMarquardt.MPC_SHIFT, integrator="rk4")
    np.testing.assert_allclose(u[:, 0], GOLD["ms_loop_u"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(x[:, 0], GOLD["ms_loop_x"], rtol=0, atol=2e-6)
    lm.clear()


@pytest.mark.gpu
def test_shift_is_refused_where_the_reference_never_shifts():
    """NonUniformFiniteDifferencesVariableGrid::isMovingHorizonWarmStartActive() is false: no silent approximation"""
    ocp = problems.unicycle_time_optimal(12)
    lm = solver.BatchedLevenbergMarquardt(ocp, 4)
    with pytest.raises(solver.B200SqpError) as e:
        lm.warm_start_shift(np.zeros((4, 3)))
    assert e.value.code == -2
    lm.clear()


@pytest.mark.gpu
def test_closed_loop_with_device_warm_start_matches_reference_controller():
    """15 MPC steps with setWarmStart(true) in the reference (shift, then solve from the shifted trajectory) against
    b200sqp_warm_start_shift + b200sqp_step(cold_start=0); plant = one RK4 step of the same dynamics."""
    ocp = problems.van_der_pol(20)
    lm = solver.BatchedLevenbergMarquardt(ocp, 1)
    lm.setIterations(10)

    def f(x, u):
        return np.array([x[1], -1.0 * (x[0] * x[0] - 1) * x[1] - x[0] + u[0]])

    x = np.array([1.0, 0.5])
    dt = ocp.dt_ref
    total_shift = 0
    for s in range(15):
        if s > 0:
            total_shift += int(lm.warm_start_shift(x[None, :])[0])
        lm.step(x[None, :], None, cold_start=(s == 0))
        u = lm.get_first_controls()[0]
        np.testing.assert_allclose(u, GOLD["loop_u"][s], rtol=0, atol=2e-6)
        k1 = f(x, u) * dt
        k2 = f(x + k1 / 2.0, u) * dt
        k3 = f(x + k2 / 2.0, u) * dt
        k4 = f(x + k3, u) * dt
        x = x + (k1 + 2.0 * k2 + 2.0 * k3 + k4) / 6.0
        np.testing.assert_allclose(x, GOLD["loop_x"][s + 1], rtol=0, atol=2e-6)
    assert total_shift >= 10  # the horizon really moved
    lm.clear()


@pytest.mark.gpu
def test_device_shift_large_batch_is_per_instance():
    ocp = problems.van_der_pol(50)
    B = 4096
    x0, _ = problems.instance_data(ocp, B, seed=9)
    lm = solver.BatchedLevenbergMarquardt(ocp, B)
    lm.setIterations(10)
    lm.set_problem_data(x0, None)
    lm.initialize_trajectories()
    lm.solve(new_run=True)
    p = lm.get_params()
    x_idx, _, _ = solver.vertex_indices(ocp)
    x1 = p[:, x_idx[1]:x_idx[1] + 2]
    x0_new = np.where((np.arange(B) % 2 == 0)[:, None], x1, x0)  # every other instance moved exactly one state ahead
    shifts = lm.warm_start_shift(x0_new)
    assert np.array_equal(shifts[1::2], np.zeros(B // 2, np.int32)) and shifts[0::2].min() >= 1
    after = lm.get_params()
    assert np.array_equal(after[1::2], p[1::2])  # untouched where the start did not move
    idx = np.arange(0, B, 2)[:16]
    for i in idx:
        xs, us, s = warm_start.warm_start_shift(*_split(ocp, p[i], x0[i]), x0_new[i])
        assert s == shifts[i] and np.array_equal(_join(ocp, xs, us), after[i])
    lm.clear()


@pytest.mark.gpu
def test_mpc_step_front_end_matches_reference_closed_loops():
    """b200sqp_mpc_step (x0 in, u0 out) in its keep and shift modes against the reference's controller loops without / with
    setWarmStart(true) (tests/golden/vdp20_closed_loop.npz, warm_start_shift.npz)."""
    plain = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vdp20_closed_loop.npz"))
    ocp = problems.van_der_pol(20)

    def f(x, u):
        return np.array([x[1], -1.0 * (x[0] * x[0] - 1) * x[1] - x[0] + u[0]])

    for mode, gold_u, gold_x in ((solver.BatchedLevenbergMarquardt.MPC_KEEP, plain["u"], plain["x"]),
                                 (solver.BatchedLevenbergMarquardt.MPC_SHIFT, GOLD["loop_u"], GOLD["loop_x"])):
        lm = solver.BatchedLevenbergMarquardt(ocp, 1)
        lm.setIterations(10)
        x = np.array([1.0, 0.5])
        for s in range(15):
            u, chi2, status = lm.mpc_step(x[None, :], None, mode=solver.BatchedLevenbergMarquardt.MPC_COLD if s == 0 else mode)
            u = u[0]
            np.testing.assert_allclose(u, gold_u[s], rtol=0, atol=2e-6)
            k1 = f(x, u) * ocp.dt_ref
            k2 = f(x + k1 / 2.0, u) * ocp.dt_ref
            k3 = f(x + k2 / 2.0, u) * ocp.dt_ref
            k4 = f(x + k3, u) * ocp.dt_ref
            x = x + (k1 + 2.0 * k2 + 2.0 * k3 + k4) / 6.0
            np.testing.assert_allclose(x, gold_x[s + 1], rtol=0, atol=2e-6)
        lm.clear()


@pytest.mark.gpu
def test_pinned_and_pageable_host_buffers_give_the_same_bits():
    """b200sqp_step / b200sqp_mpc_step read pinned start states in place and write the small results straight into pinned buffers
    (ingest / export kernels); pageable buffers take the cudaMemcpyAsync path.  Both must deliver identical results."""
    import torch

    ocp = problems.van_der_pol(30)
    B = 200  # not a multiple of the tile or the block size
    x0, _ = problems.instance_data(ocp, B, seed=21)
    xref = np.tile(np.array([0.2, -0.1]), (B, 1))
    lm = solver.BatchedLevenbergMarquardt(ocp, B)
    lm.setIterations(6)
    n = lm.dims.n_params
    # pageable (numpy) path
    p_a, chi2_a, st_a = lm.step(x0, xref, cold_start=True)
    u_a, chi2_ma, st_ma = lm.mpc_step(x0, xref, mode=solver.BatchedLevenbergMarquardt.MPC_COLD)
    # pinned path
    h_x0, h_xref = torch.from_numpy(x0.copy()).pin_memory(), torch.from_numpy(xref.copy()).pin_memory()
    h_p = torch.zeros((B, n), dtype=torch.float64).pin_memory()
    h_chi2, h_st = torch.zeros(B, dtype=torch.float64).pin_memory(), torch.zeros(B, dtype=torch.int32).pin_memory()
    h_u = torch.zeros((B, ocp.nu), dtype=torch.float64).pin_memory()
    lm.step_raw(h_x0.data_ptr(), h_xref.data_ptr(), h_p.data_ptr(), h_chi2.data_ptr(), h_st.data_ptr(), cold_start=True)
    assert np.array_equal(h_p.numpy(), p_a) and np.array_equal(h_chi2.numpy(), chi2_a) and np.array_equal(h_st.numpy(), st_a)
    h_chi2.zero_()
    h_st.zero_()
    lm.mpc_step_raw(0, h_x0.data_ptr(), h_xref.data_ptr(), h_u.data_ptr(), h_chi2.data_ptr(), h_st.data_ptr())
    assert np.array_equal(h_u.numpy(), u_a) and np.array_equal(h_chi2.numpy(), chi2_ma) and np.array_equal(h_st.numpy(), st_ma)
    # zero reference through the pinned path (xref = NULL)
    u_b, chi2_b, _ = lm.mpc_step(x0, None, mode=0)
    lm.mpc_step_raw(0, h_x0.data_ptr(), 0, h_u.data_ptr(), h_chi2.data_ptr(), h_st.data_ptr())
    assert np.array_equal(h_u.numpy(), u_b) and np.array_equal(h_chi2.numpy(), chi2_b)
    lm.clear()

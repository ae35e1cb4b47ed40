"""plant simulation (SimulatedPlant::control over the reference's explicit integrators) and the whole
closed loop of ClosedLoopControlTask (PredictiveController + SimulatedPlant) on the device.  tests/golden/plant.npz holds the compiled
reference's own outputs (tests/golden/make_golden.py plant); the oracle is pinned against it on CPU, the device path
(b200sqp_plant_step, b200sqp_closed_loop) against both on the GPU."""
import os
import sys

import numpy as np
import pytest

from control_box_rst_b200 import _abi as abi
from control_box_rst_b200 import problems, solver

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import cases  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "plant.npz"))
LOOPS = [("euler", "keep", 1), ("rk4", "keep", 1), ("rk4", "shift", 2)]
# closed-loop trajectories: the FD-noise floor of the solver's central differences, as in test_gpu_parity.py
LOOP_ATOL = 2e-6


def _ulp_tol(ref):
    # a few ulps of sin/cos (device libm vs glibc) through at most four dynamics evaluations
    return 2e-14 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("integrator", ["euler", "rk4"])
@pytest.mark.parametrize("name", list(cases.LINEARIZE_MODELS))
def test_oracle_plant_step_matches_reference_fixture(oracle, name, integrator):
    make, polynomial = cases.LINEARIZE_MODELS[name]
    ocp = make()
    xs, us = cases.linearize_points(ocp)
    xn = oracle.plant_step(ocp, xs, us, cases.PLANT_DT, integrator)
    ref = GOLD[f"{name}_{integrator}"]
    if polynomial:
        assert np.array_equal(xn, ref)
    else:
        np.testing.assert_allclose(xn, ref, rtol=0, atol=_ulp_tol(ref))


@pytest.mark.parametrize("integrator,kind,mode", LOOPS)
def test_reference_plant_integrates_over_its_nanosecond_clock_interval(oracle, integrator, kind, mode):
    """The reference's closed-loop log is reproduced bit for bit from its own (x, u) pairs only if the plant interval is
    (t + dt) - t on the integer-nanosecond clock (TimeValueBuffer::getValues) -- pins oracle.plant_interval."""
    ocp = problems.van_der_pol(20)
    u, x = GOLD[f"loop_{integrator}_{kind}_u"], GOLD[f"loop_{integrator}_{kind}_x"]
    assert np.array_equal(x[0], cases.closed_loop_starts())
    intervals = [oracle.plant_interval(ocp.dt_ref, s) for s in range(len(u))]
    assert any(dt != ocp.dt_ref for dt in intervals) and max(abs(dt - ocp.dt_ref) for dt in intervals) < 1e-15
    for s in range(len(u)):
        assert np.array_equal(oracle.plant_step(ocp, x[s], u[s], intervals[s], integrator), x[s + 1])


def test_plant_entry_points_reject_bad_arguments_without_a_device():
    lib = solver.load_library()
    import ctypes as C
    z = (C.c_double * 8)()
    assert lib.b200sqp_plant_step(C.c_int32(abi.DYN_VAN_DER_POL), z, C.c_int32(2), C.c_double(0.1), C.c_int32(1), z, z, z, C.c_int32(0)) == -1
    assert lib.b200sqp_plant_step(C.c_int32(999), z, C.c_int32(0), C.c_double(0.1), C.c_int32(1), z, z, z, C.c_int32(0)) == -2
    assert lib.b200sqp_closed_loop(None, None, C.c_int32(1), C.c_int32(0), C.c_double(0.1), C.c_int32(1), z, None, None, None, None, None) == -1


@pytest.mark.gpu
@pytest.mark.parametrize("integrator", ["euler", "rk4"])
@pytest.mark.parametrize("name", list(cases.LINEARIZE_MODELS))
def test_device_plant_step_matches_reference_fixture_and_oracle(oracle, name, integrator):
    make, polynomial = cases.LINEARIZE_MODELS[name]
    ocp = make()
    xs, us = cases.linearize_points(ocp)
    xn = solver.plant_step(ocp.dynamics, list(ocp.dyn_params), xs, us, cases.PLANT_DT, integrator)
    ref = GOLD[f"{name}_{integrator}"]
    xo = oracle.plant_step(ocp, xs, us, cases.PLANT_DT, integrator)
    if polynomial:
        assert np.array_equal(xn, ref) and np.array_equal(xn, xo)  # bit-exact: same IEEE expressions, no FMA contraction
    else:
        np.testing.assert_allclose(xn, ref, rtol=0, atol=_ulp_tol(ref))
        np.testing.assert_allclose(xn, xo, rtol=0, atol=_ulp_tol(ref))


@pytest.mark.gpu
@pytest.mark.parametrize("integrator,kind,mode", LOOPS)
def test_device_closed_loop_matches_reference_controller_and_plant(oracle, integrator, kind, mode):
    """ClosedLoopControlTask's loop of the compiled reference (PredictiveController::step + SimulatedPlant::control) for six start
    states against one b200sqp_closed_loop call; the log must also be self-consistent bit for bit with the oracle's plant."""
    ocp = problems.van_der_pol(20)
    x0 = GOLD["loop_x0"]
    u_ref, x_ref = GOLD[f"loop_{integrator}_{kind}_u"], GOLD[f"loop_{integrator}_{kind}_x"]
    steps = len(u_ref)
    lm = solver.BatchedLevenbergMarquardt(ocp, len(x0))
    lm.setIterations(10)
    u, x, chi2, status = lm.closed_loop(x0, steps, mode=mode, integrator=integrator)
    np.testing.assert_allclose(u, u_ref, rtol=0, atol=LOOP_ATOL)
    np.testing.assert_allclose(x, x_ref, rtol=0, atol=LOOP_ATOL)
    assert np.array_equal(x[0], x0) and np.isfinite(chi2).all() and (status >= 0).all()
    for s in range(steps):
        assert np.array_equal(oracle.plant_step(ocp, x[s], u[s], oracle.plant_interval(ocp.dt_ref, s), integrator), x[s + 1])
    # the same loop driven from the host through b200sqp_mpc_step + b200sqp_plant_step gives the same bits
    xs = x0.copy()
    for s in range(steps):
        us, chi2_s, _ = lm.mpc_step(xs, None, mode=(0 if s == 0 else mode))
        assert np.array_equal(us, u[s]) and np.array_equal(chi2_s, chi2[s])
        xs = solver.plant_step(ocp.dynamics, list(ocp.dyn_params), xs, us, oracle.plant_interval(ocp.dt_ref, s), integrator)
        assert np.array_equal(xs, x[s + 1])
    lm.clear()


@pytest.mark.gpu
def test_device_closed_loop_monte_carlo_batch_regulates_and_is_instancewise():
    """BenchmarkTaskVaryingInitialState-style study at BASELINE's batch: 4096 start states, 40 closed-loop steps wholly on the device.
    Size-independent properties: every instance is regulated towards the origin, and an instance's
    closed loop does not depend on its neighbours (a re-run of a subset gives the same bits)."""
    ocp = problems.van_der_pol(50)
    B, steps = 4096, 40
    x0, _ = problems.instance_data(ocp, B)
    lm = solver.BatchedLevenbergMarquardt(ocp, B)
    lm.setIterations(5)
    u, x, chi2, status = lm.closed_loop(x0, steps, mode=2, integrator="rk4")
    lm.clear()
    assert np.isfinite(x).all() and np.isfinite(u).all() and (status >= 0).all()
    # (|u| <= 1 is a quadratic penalty of weight 2 in this solver, not a hard bound, so no assertion on it)
    n0, n1 = np.linalg.norm(x[0], axis=1), np.linalg.norm(x[-1], axis=1)
    print(f"closed loop 4096 x 40: median |x| {np.median(n0):.3f} -> {np.median(n1):.3f}, max {n0.max():.3f} -> {n1.max():.3f}, max|u| {np.abs(u).max():.3f}")
    assert np.median(n1) < 0.5 * np.median(n0)
    idx = np.arange(0, B, 64)
    lm2 = solver.BatchedLevenbergMarquardt(ocp, len(idx))
    lm2.setIterations(5)
    u2, x2, chi2_2, _ = lm2.closed_loop(x0[idx], steps, mode=2, integrator="rk4")
    lm2.clear()
    assert np.array_equal(u2, u[:, idx]) and np.array_equal(x2, x[:, idx]) and np.array_equal(chi2_2, chi2[:, idx])


@pytest.mark.gpu
@pytest.mark.parametrize("name,make,mode,integrator", [
    ("quadrotor_pipeline", lambda: problems.quadrotor(12), 2, "rk4"),          # warp-cooperative pipeline inside the loop
    ("cartpole_shooting", lambda: problems.cart_pole_shooting(20), 2, "euler"),  # shooting grid, shift mode, trigonometric model
    ("unicycle_timeopt", lambda: problems.unicycle_time_optimal(15), 1, "rk4"),  # non-uniform grid: keep mode (the reference never shifts it)
], ids=lambda v: v if isinstance(v, str) else None)
def test_device_closed_loop_equals_host_driven_loop_on_other_models(oracle, name, make, mode, integrator):
    """b200sqp_closed_loop against the same loop driven from the host through b200sqp_mpc_step + b200sqp_plant_step: identical bits
    for every model family / solver path, and plant steps consistent with the oracle's restatement of SimulatedPlant::control."""
    ocp = make()
    B, steps = 5, 4
    x0, xref = problems.instance_data(ocp, B, seed=31)
    lm = solver.BatchedLevenbergMarquardt(ocp, B)
    lm.setIterations(4)
    u, x, chi2, status = lm.closed_loop(x0, steps, xref=xref, mode=mode, integrator=integrator)
    assert np.isfinite(u).all() and np.isfinite(x).all() and (status >= 0).all()
    xs = x0.copy()
    for s in range(steps):
        us, chi2_s, _ = lm.mpc_step(xs, xref, mode=(0 if s == 0 else mode))
        assert np.array_equal(us, u[s]) and np.array_equal(chi2_s, chi2[s])
        dt_s = oracle.plant_interval(ocp.dt_ref, s)
        xs = solver.plant_step(ocp.dynamics, list(ocp.dyn_params), xs, us, dt_s, integrator)
        assert np.array_equal(xs, x[s + 1])
        xo = oracle.plant_step(ocp, x[s], u[s], dt_s, integrator)
        np.testing.assert_allclose(x[s + 1], xo, rtol=0, atol=_ulp_tol(xo))
    # a longer run on the same handle (the log buffers grow) starts cold again and reproduces the shorter one as its prefix
    u2, x2, chi2_2, _ = lm.closed_loop(x0, steps + 3, xref=xref, mode=mode, integrator=integrator)
    assert np.array_equal(u2[:steps], u) and np.array_equal(x2[:steps + 1], x) and np.array_equal(chi2_2[:steps], chi2)
    lm.clear()

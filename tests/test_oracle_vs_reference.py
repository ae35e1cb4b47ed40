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


def _random_ocp(rng):
    """A random member of the supported family: model, grid, defect, horizon, step, weights of the cost, bounds, goal handling."""
    poly = [(abi.DYN_VAN_DER_POL, (1.0 + rng.uniform(-0.5, 0.5),)), (abi.DYN_DUFFING, (1.0, -1.0, 1.0)), (abi.DYN_TOY_EXAMPLE, (rng.uniform(0.2, 0.8),)),
            (abi.DYN_ARTSTEINS_CIRCLE, ()), (abi.DYN_DOUBLE_INTEGRATOR, (rng.uniform(0.5, 2.0),)), (abi.DYN_TRIPLE_INTEGRATOR, (1.5,)),
            (abi.DYN_LINEAR_2X1, (-0.3, -2.0, 1.1, -0.5, 0.2, 1.5)), (abi.DYN_LINEAR_3X1, (-0.5, 0.2, -0.1, 1.0, -0.4, 0.3, 0.1, 1.2, -0.8, 0.3, -0.2, 1.0))]
    dynamics, params = poly[rng.integers(len(poly))]
    nx, nu = abi.DYN_DIMS[dynamics]
    grid = int(rng.integers(3))
    n_grid = int(rng.integers(3, 24))
    dt = float(rng.choice([0.05, 0.1, 0.2]))
    kw = dict(grid=grid, dynamics=dynamics, n_grid=n_grid, dt=dt, dyn_params=params, u_lb=(-1.0,) * nu, u_ub=(1.5,) * nu)
    if grid == abi.GRID_FD_NONUNIFORM_VARDT:
        kw.update(stage_cost=abi.COST_MINIMUM_TIME_LSQ, xf_fixed=(1,) * nx, dt_lb=0.0, dt_ub=1.0, collocation=int(rng.integers(4)))
    else:
        kw.update(q=tuple(rng.uniform(0.5, 2.0, nx)), r=tuple(rng.uniform(0.05, 0.5, nu)))
        if rng.random() < 0.7:
            kw.update(qf=tuple(rng.uniform(0.5, 5.0, nx)))
        if rng.random() < 0.4:
            kw.update(x_lb=(-3.0,) * nx, x_ub=(2.5,) * nx)
        if grid == abi.GRID_FD_UNIFORM:
            kw.update(collocation=int(rng.integers(4)))
            if rng.random() < 0.3:
                kw.update(xf_fixed=tuple(int(b) for b in rng.integers(0, 2, nx)))
        else:
            kw.update(integrator=int(rng.choice([abi.INT_EULER, abi.INT_RK4])))
    return problems.make_ocp(**kw)


@pytest.mark.parametrize("seed", range(16))
def test_random_structures_match_compiled_reference(oracle, reference, seed):
    """Seeded random OCPs over the polynomial / rational models: dimensions, vertex and edge indices, the initial guess, the value
    vector, the combined Jacobian (pattern and values) and the parameter drift of the in-place differences equal the compiled
    reference's bit for bit, at a perturbed point with random penalty weights."""
    rng = np.random.default_rng(1000 + seed)
    ocp = _random_ocp(rng)
    d_r, d_o = reference.dims(ocp), oracle.dims(ocp)
    for f in ("n_params", "m_lsq", "m_eq", "m_ineq", "m_bounds", "nnz_jacobian", "nnz_hessian_upper", "algorithmic_bytes_per_iteration"):
        assert getattr(d_r, f) == getattr(d_o, f), f
    for a, b in zip(reference.vertex_indices(ocp), oracle.vertex_indices(ocp)):
        assert np.array_equal(a, b)
    for cat in range(3):
        assert np.array_equal(reference.edge_table(ocp, cat), oracle.edge_table(ocp, cat))
    x0, xref = problems.instance_data(ocp, 1, seed=seed)
    p_r, p_o = reference.initial_params(ocp, x0[0], xref[0]), oracle.initial_params(ocp, x0[0], xref[0])
    np.testing.assert_allclose(p_o, p_r, rtol=0, atol=1e-15)
    p = p_r + rng.uniform(-0.3, 0.3, p_r.shape)
    if ocp.grid == abi.GRID_FD_NONUNIFORM_VARDT:
        dt_idx = reference.vertex_indices(ocp)[2]
        p[dt_idx] = np.abs(p[dt_idx]) + 0.05
    weights = tuple(rng.uniform(1.0, 10.0, 3))
    v_r, J_r, P_r, a_r = reference.evaluate(ocp, x0[0], xref[0], p, weights)
    v_o, J_o, P_o, a_o = oracle.evaluate(ocp, x0[0], xref[0], p, weights)
    assert np.array_equal(P_r, P_o)
    assert np.array_equal(v_r, v_o) and np.array_equal(J_r, J_o) and np.array_equal(a_r, a_o)


@pytest.mark.parametrize("seed", range(16))
def test_random_structures_first_lm_iterations_match_compiled_reference(oracle, reference, seed):
    """The same random OCPs through the solver: three LM iterations from the reference's initial guess for six random start states
    (before the finite-difference noise floor is reached) -- the event sequence of instance 0 is identical, trajectories agree to
    1e-5 relative and chi2 to 1e-6 (worst observed 2.3e-6 / 1.2e-7, on the time-optimal grids)."""
    rng = np.random.default_rng(1000 + seed)
    ocp = _random_ocp(rng)
    weights = tuple(rng.uniform(1.0, 10.0, 3))
    x0, xref = problems.instance_data(ocp, 6, seed=seed)
    opts = abi.LmOptions.defaults(iterations=3, weights=weights)
    p_r, c_r, s_r, _ = reference.solve_batch(ocp, opts, x0, xref, threads=2)
    p_o, c_o, s_o, _ = oracle.solve_batch(ocp, opts, x0, xref, threads=2)
    err = np.abs(p_o - p_r).max(axis=1) / np.maximum(1.0, np.abs(p_r).max(axis=1))
    assert err.max() <= 1e-5, err
    np.testing.assert_allclose(c_o, c_r, rtol=1e-6)
    assert np.array_equal(s_r, s_o)
    tr_r, tr_o = reference.trace(ocp, opts, x0[0], xref[0]), oracle.trace(ocp, opts, x0[0], xref[0])
    assert [e[0] for e in tr_r["events"]] == [e[0] for e in tr_o["events"]]


@pytest.mark.parametrize("make", [lambda: problems.van_der_pol(12), lambda: problems.van_der_pol(11, collocation=abi.COLL_FORWARD, final_cost=False),
                                  lambda: problems.make_ocp(grid=abi.GRID_FD_UNIFORM, dynamics=abi.DYN_DUFFING, n_grid=9, dt=0.1, q=(1.0, 2.0), r=(0.3,),
                                                            qf=(2.0, 1.0), u_lb=(-1.0,), u_ub=(1.0,), dyn_params=(1.0, -1.0, 1.0))],
                         ids=["vdp12_fd", "vdp11_forward_nofinal", "duffing9_fd"])
def test_time_varying_reference_oracle_matches_compiled_reference(oracle, reference, make):
    """A non-static state reference (ReferenceTrajectoryInterface::isStatic() == false, core/reference_trajectory.h:60-95): the cost edge of
    grid point k measures x_k against getReferenceCached(k), the initial guess is the reference trajectory itself
    (full_discretization_grid_base.cpp:181-228), the goal is its last point.  Values, Jacobian and drift bit-identical; solves agree.
    (Full-discretisation grids only: on a cold start the shooting grids put xref(0) into the FIRST shooting node and fix it there,
    shooting_grid_base.cpp:259,278, i.e. they ignore the measured state -- the device refuses that combination instead of mirroring it.)"""
    ocp = make()
    N, nx = ocp.n_grid, ocp.nx
    rng = np.random.default_rng(5)
    B = 6
    x0, _ = problems.instance_data(ocp, B, seed=3)
    t = np.linspace(0.0, 1.0, N)
    xref = np.stack([np.stack([0.5 * np.sin(2.0 * t + i) + 0.1 * j for j in range(nx)], axis=1) for i in range(B)])  # [B, N, nx]
    for chk in (oracle, reference):
        chk.set_xref_points(N)
    try:
        p_r, p_o = reference.initial_params(ocp, x0[0], xref[0]), oracle.initial_params(ocp, x0[0], xref[0])
        assert np.array_equal(p_r, p_o)
        p = p_r + rng.uniform(-0.2, 0.2, p_r.shape)
        v_r, J_r, P_r, a_r = reference.evaluate(ocp, x0[0], xref[0], p, (2.0, 3.0, 4.0))
        v_o, J_o, P_o, a_o = oracle.evaluate(ocp, x0[0], xref[0], p, (2.0, 3.0, 4.0))
        assert np.array_equal(P_r, P_o) and np.array_equal(v_r, v_o) and np.array_equal(J_r, J_o) and np.array_equal(a_r, a_o)
        opts = abi.LmOptions.defaults(iterations=6)
        pr, cr, sr, _ = reference.solve_batch(ocp, opts, x0, xref, threads=2)
        po, co, so, _ = oracle.solve_batch(ocp, opts, x0, xref, threads=2)
        err = np.abs(po - pr).max(axis=1) / np.maximum(1.0, np.abs(pr).max(axis=1))
        assert err.max() <= 1e-5, err
        np.testing.assert_allclose(co, cr, rtol=1e-6)
        # and the trajectory really matters: a static reference at the last point gives another optimum
        for chk in (oracle, reference):
            chk.set_xref_points(0)
        ps, _, _, _ = oracle.solve_batch(ocp, opts, x0, np.ascontiguousarray(xref[:, -1, :]), threads=2)
        assert np.abs(ps - po).max() > 1e-3
    finally:
        for chk in (oracle, reference):
            chk.set_xref_points(0)


def _dense_weight_cases():
    Q2 = np.array([[2.0, 0.3], [0.3, 1.0]])
    Qf2 = np.array([[3.0, -0.4], [-0.4, 2.0]])
    Q3 = np.array([[2.0, 0.3, -0.2], [0.3, 1.0, 0.1], [-0.2, 0.1, 1.5]])
    R2 = np.array([[0.5, 0.1], [0.1, 0.3]])
    return {
        "vdp10_full_q_qf": (lambda: problems.van_der_pol(10, q_full=Q2, qf_full=Qf2), (0.2, -0.1), True),
        "vdp9_ms_full_q": (lambda: problems.van_der_pol_shooting(9, q_full=Q2), (0.1, 0.3), True),
        "rocket8_full_q3": (lambda: problems.make_ocp(grid=abi.GRID_FD_UNIFORM, dynamics=abi.DYN_FREE_SPACE_ROCKET, n_grid=8, dt=0.1, q=(1, 1, 1), r=(0.1,),
                                                      qf=(1, 1, 1), u_lb=(-1.1,), u_ub=(1.1,), q_full=Q3, qf_full=2.0 * Q3), (0.1, 0.2, 0.8), True),
        "unicycle8_full_q3_r2": (lambda: problems.make_ocp(grid=abi.GRID_FD_UNIFORM, dynamics=abi.DYN_UNICYCLE, n_grid=8, dt=0.1, q=(1, 1, 1), r=(0.1, 0.1),
                                                           qf=(1, 1, 1), u_lb=(-1, -1), u_ub=(1, 1), q_full=Q3, r_full=R2), (0.1, 0.2, 0.3), False),
        "vdp10_full_but_diagonal": (lambda: problems.van_der_pol(10, q_full=np.diag([2.0, 0.5]), qf_full=np.diag([1.5, 1.0])), (0.2, -0.1), True),
    }


DENSE_CASES = _dense_weight_cases()


@pytest.mark.parametrize("name", list(DENSE_CASES))
def test_full_weight_matrices_oracle_matches_compiled_reference(oracle, reference, name):
    """Non-diagonal Q / R / Qf: QuadraticFormCost / QuadraticFinalStateCost take the upper Cholesky factor as square root
    (quadratic_cost.cpp:32-96, final_state_cost.cpp:38-58) and the cost edges get dense Jacobian blocks.  Values, Jacobian (pattern and
    entries) and drift bit-identical for up to three states (Eigen's sums are sequential there); a matrix handed over in full but
    diagonal takes the reference's diagonal branch."""
    make, xr, exact = DENSE_CASES[name]
    ocp = make()
    B = 5
    x0, _ = problems.instance_data(ocp, B, seed=2)
    xref = np.tile(np.array(xr, dtype=float), (B, 1))
    rng = np.random.default_rng(1)
    p_r = reference.initial_params(ocp, x0[0], xref[0])
    p = p_r + rng.uniform(-0.2, 0.2, p_r.shape)
    v_r, J_r, P_r, a_r = reference.evaluate(ocp, x0[0], xref[0], p, (2.0, 3.0, 4.0))
    v_o, J_o, P_o, a_o = oracle.evaluate(ocp, x0[0], xref[0], p, (2.0, 3.0, 4.0))
    assert np.array_equal(P_r, P_o)
    if exact:
        assert np.array_equal(v_r, v_o) and np.array_equal(J_r, J_o) and np.array_equal(a_r, a_o)
    else:
        np.testing.assert_allclose(v_o, v_r, rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(J_o, J_r, rtol=0, atol=2e-6 * max(1.0, np.abs(J_r).max()))
    opts = abi.LmOptions.defaults(iterations=6)
    pr, cr, sr, _ = reference.solve_batch(ocp, opts, x0, xref, threads=2)
    po, co, so, _ = oracle.solve_batch(ocp, opts, x0, xref, threads=2)
    err = np.abs(po - pr).max(axis=1) / np.maximum(1.0, np.abs(pr).max(axis=1))
    assert err.max() <= (1e-5 if exact else 1e-3), err
    np.testing.assert_allclose(co, cr, rtol=1e-6)


DT_EQ_CASES = {
    "unicycle12_timeopt_dteq": lambda: problems.unicycle_time_optimal(12, dt_eq_constraint=True),
    "vdp10_timeopt_dteq": lambda: problems.make_ocp(grid=abi.GRID_FD_NONUNIFORM_VARDT, dynamics=abi.DYN_VAN_DER_POL, n_grid=10, dt=0.1,
                                                     stage_cost=abi.COST_MINIMUM_TIME_LSQ, u_lb=(-1.0,), u_ub=(1.0,), xf_fixed=(1, 1), dt_lb=0.0, dt_ub=1.0,
                                                     dyn_params=(1.0,), dt_eq_constraint=True),
    "rocket9_timeopt_dteq": lambda: _with_dt_eq(problems.free_space_rocket_time_optimal(9)),
    "dint3_timeopt_dteq": lambda: problems.make_ocp(grid=abi.GRID_FD_NONUNIFORM_VARDT, dynamics=abi.DYN_DOUBLE_INTEGRATOR, n_grid=3, dt=0.2,
                                                     stage_cost=abi.COST_MINIMUM_TIME_LSQ, u_lb=(-1.0,), u_ub=(1.0,), xf_fixed=(1, 1), dt_lb=0.01, dt_ub=1.0,
                                                     dyn_params=(1.0,), dt_eq_constraint=True),
}


def _with_dt_eq(ocp):
    ocp.dt_eq_constraint = 1
    return ocp


@pytest.mark.parametrize("name", list(DT_EQ_CASES))
def test_dt_equality_edges_oracle_matches_compiled_reference(oracle, reference, name):
    """NonUniformFiniteDifferencesVariableGrid::setDtEqConstraint(true): a TwoScalarEqualEdge (edges/misc_edges.h:40-67) after the dynamics
    edge of every interval k >= 1 ties dt_k to dt_{k-1}.  Dimensions, edge tables (row offsets, vertex indices), values, the Jacobian and
    the parameter drift equal the compiled reference's bit for bit (polynomial models); solves agree."""
    ocp = DT_EQ_CASES[name]()
    d_r, d_o = reference.dims(ocp), oracle.dims(ocp)
    for f in ("n_params", "m_lsq", "m_eq", "m_ineq", "m_bounds", "nnz_jacobian", "nnz_hessian_upper", "algorithmic_bytes_per_iteration"):
        assert getattr(d_r, f) == getattr(d_o, f), f
    assert d_r.m_eq == (ocp.n_grid - 1) * ocp.nx + (ocp.n_grid - 2)
    for cat in range(3):
        assert np.array_equal(reference.edge_table(ocp, cat), oracle.edge_table(ocp, cat))
    B = 6
    x0, xref = problems.instance_data(ocp, B, seed=4)
    rng = np.random.default_rng(2)
    p_r = reference.initial_params(ocp, x0[0], xref[0])
    p = p_r + rng.uniform(-0.2, 0.2, p_r.shape)
    dt_idx = reference.vertex_indices(ocp)[2]
    p[dt_idx] = np.abs(p[dt_idx]) + 0.05
    w = (2.0, 3.0, 4.0)
    v_r, J_r, P_r, a_r = reference.evaluate(ocp, x0[0], xref[0], p, w)
    v_o, J_o, P_o, a_o = oracle.evaluate(ocp, x0[0], xref[0], p, w)
    assert np.array_equal(P_r, P_o)
    if ocp.dynamics != abi.DYN_UNICYCLE:
        assert np.array_equal(v_r, v_o) and np.array_equal(J_r, J_o) and np.array_equal(a_r, a_o)
    else:
        np.testing.assert_allclose(v_o, v_r, rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(J_o, J_r, rtol=0, atol=2e-6 * max(1.0, np.abs(J_r).max()))
    opts = abi.LmOptions.defaults(iterations=5, weights=w)
    pr, cr, sr, _ = reference.solve_batch(ocp, opts, x0, xref, threads=2)
    po, co, so, _ = oracle.solve_batch(ocp, opts, x0, xref, threads=2)
    err = np.abs(po - pr).max(axis=1) / np.maximum(1.0, np.abs(pr).max(axis=1))
    assert err.max() <= 1e-4, err
    np.testing.assert_allclose(co, cr, rtol=1e-5)

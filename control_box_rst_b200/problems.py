"""OCP descriptors for BASELINE.json's configs and seeded synthetic instance data.

Each builder returns a `b200sqp_ocp` (control_box_rst_b200._abi.Ocp) describing what the reference would assemble out of
StructuredOptimalControlProblem + a discretization grid + stage functions; `instance_data` draws the per-instance start states.
"""
import math

import numpy as np

from . import _abi as abi


def _fill(arr, values, default=0.0):
    for i in range(len(arr)):
        arr[i] = values[i] if i < len(values) else default


def make_ocp(*, grid, dynamics, n_grid, dt, collocation=abi.COLL_CRANK_NICOLSON, integrator=abi.INT_RK4, stage_cost=abi.COST_QUADRATIC_LSQ,
             q=(), r=(), qf=None, x_lb=None, x_ub=None, u_lb=None, u_ub=None, xf_fixed=None, dt_lb=0.0, dt_ub=abi.CORBO_INF_DBL,
             dyn_params=(), terminal_equality=None, terminal_ball=None, q_full=None, r_full=None, qf_full=None, dt_eq_constraint=False):
    nx, nu = abi.DYN_DIMS[dynamics]
    d = abi.Ocp()
    d.grid, d.dynamics, d.collocation, d.integrator = grid, dynamics, collocation, integrator
    d.n_grid, d.nx, d.nu = n_grid, nx, nu
    d.stage_cost = stage_cost
    d.final_cost = 1 if qf is not None else 0
    d.zero_x_ref, d.zero_u_ref = 0, 1
    _fill(d.xf_fixed, [int(b) for b in (xf_fixed or [])], 0)
    d.dt_ref, d.dt_lb, d.dt_ub = dt, dt_lb, dt_ub
    d.dt_eq_constraint = 1 if dt_eq_constraint else 0
    _fill(d.dyn_params, list(dyn_params))
    _fill(d.q_diag, list(q))
    _fill(d.r_diag, list(r))
    _fill(d.qf_diag, list(qf or []))
    inf = abi.CORBO_INF_DBL
    _fill(d.x_lb, list(x_lb) if x_lb is not None else [-inf] * nx, -inf)
    _fill(d.x_ub, list(x_ub) if x_ub is not None else [inf] * nx, inf)
    _fill(d.u_lb, list(u_lb) if u_lb is not None else [-inf] * nu, -inf)
    _fill(d.u_ub, list(u_ub) if u_ub is not None else [inf] * nu, inf)
    # full (non-diagonal) weight matrices: row-major nx x nx / nu x nu (quadratic_cost.cpp:32-76: Cholesky-upper square root)
    for name, mat, dim in (("q", q_full, nx), ("r", r_full, nu), ("qf", qf_full, nx)):
        if mat is not None:
            m = np.asarray(mat, np.float64).reshape(dim, dim)
            setattr(d, name + "_dense", 1)
            _fill(getattr(d, name + "_full"), list(m.reshape(-1)))
            if name == "qf":
                d.final_cost = 1
    # final-stage constraint: TerminalEqualityConstraint(xref) or TerminalBall(diag S, gamma)
    if terminal_equality is not None:
        d.final_constraint = abi.FINAL_CONSTRAINT_EQUALITY
        _fill(d.term_xref, list(terminal_equality))
    elif terminal_ball is not None:
        d.final_constraint = abi.FINAL_CONSTRAINT_BALL
        _fill(d.term_s_diag, list(terminal_ball[0]))
        d.term_gamma = float(terminal_ball[1])
    return d


def van_der_pol(n_grid=50, dt=0.1, final_cost=True, collocation=abi.COLL_CRANK_NICOLSON, a=1.0, **kw):
    """configs[0] (N=20) / configs[1] (N=50): VdP, FiniteDifferencesGrid, CN collocation, Q=I, R=0.1, |u|<=1."""
    return make_ocp(grid=abi.GRID_FD_UNIFORM, dynamics=abi.DYN_VAN_DER_POL, n_grid=n_grid, dt=dt, collocation=collocation,
                    q=(1.0, 1.0), r=(0.1,), qf=(1.0, 1.0) if final_cost else None, u_lb=(-1.0,), u_ub=(1.0,), dyn_params=(a,), **kw)


def van_der_pol_shooting(n_grid=20, dt=0.1, integrator=abi.INT_RK4, a=1.0, **kw):
    """Van der Pol on a MultipleShootingGrid (one control per interval): the polynomial shooting case of the parity tests."""
    return make_ocp(grid=abi.GRID_MULTIPLE_SHOOTING, dynamics=abi.DYN_VAN_DER_POL, n_grid=n_grid, dt=dt, integrator=integrator,
                    q=(1.0, 1.0), r=(0.1,), qf=(1.0, 1.0), u_lb=(-1.0,), u_ub=(1.0,), dyn_params=(a,), **kw)


def unicycle_time_optimal(n_grid=50, dt=0.1, **kw):
    """configs[2]: unicycle, NonUniformFiniteDifferencesVariableGrid, MinimumTime(lsq), xf fixed, |v|,|w|<=1, dt in [0,1]."""
    return make_ocp(grid=abi.GRID_FD_NONUNIFORM_VARDT, dynamics=abi.DYN_UNICYCLE, n_grid=n_grid, dt=dt,
                    stage_cost=abi.COST_MINIMUM_TIME_LSQ, u_lb=(-1.0, -1.0), u_ub=(1.0, 1.0), xf_fixed=(1, 1, 1), dt_lb=0.0, dt_ub=1.0, **kw)


def cart_pole_shooting(n_grid=100, dt=0.02, **kw):
    """configs[3]: CartPole defaults, MultipleShootingGrid, RK4, Q=I4, R=0.01, Qf=10 I4, |u|<=20 (weights 10 set on the solver)."""
    return make_ocp(grid=abi.GRID_MULTIPLE_SHOOTING, dynamics=abi.DYN_CART_POLE, n_grid=n_grid, dt=dt, integrator=abi.INT_RK4,
                    q=(1.0,) * 4, r=(0.01,), qf=(10.0,) * 4, u_lb=(-20.0,), u_ub=(20.0,), dyn_params=(1.0, 0.3, 0.5, 9.81), **kw)


def free_space_rocket(n_grid=20, dt=0.1, **kw):
    """FreeSpaceRocket (nonlinear_benchmark_systems.h:154-184; states s, v, m) on the fixed-dt FiniteDifferencesGrid."""
    return make_ocp(grid=abi.GRID_FD_UNIFORM, dynamics=abi.DYN_FREE_SPACE_ROCKET, n_grid=n_grid, dt=dt, q=(1.0, 1.0, 1.0), r=(0.1,),
                    qf=(1.0, 1.0, 1.0), u_lb=(-1.1,), u_ub=(1.1,), **kw)


def free_space_rocket_time_optimal(n_grid=20, dt=0.1):
    """The reference's classic time-optimal rocket: MinimumTime(lsq) on the NonUniformFiniteDifferencesVariableGrid, position and
    velocity of the final state fixed, final mass free, |u| <= 1.1, v in [-0.5, 1.7], m >= 0."""
    inf = abi.CORBO_INF_DBL
    return make_ocp(grid=abi.GRID_FD_NONUNIFORM_VARDT, dynamics=abi.DYN_FREE_SPACE_ROCKET, n_grid=n_grid, dt=dt,
                    stage_cost=abi.COST_MINIMUM_TIME_LSQ, u_lb=(-1.1,), u_ub=(1.1,), x_lb=(-inf, -0.5, 0.0), x_ub=(inf, 1.7, inf),
                    xf_fixed=(1, 1, 0), dt_lb=0.0, dt_ub=1.0)


def quadrotor(n_grid=60, dt=0.05, **kw):
    """configs[4]: 12-state quadrotor, FiniteDifferencesGrid, quadratic lsq cost, thrust/torque bounds."""
    m, g = 1.0, 9.81
    return make_ocp(grid=abi.GRID_FD_UNIFORM, dynamics=abi.DYN_QUADROTOR, n_grid=n_grid, dt=dt,
                    q=(1.0,) * 12, r=(0.1,) * 4, qf=(1.0,) * 12, u_lb=(0.0, -1.0, -1.0, -1.0), u_ub=(2.0 * m * g, 1.0, 1.0, 1.0),
                    dyn_params=(m, g, 0.01, 0.01, 0.02), **kw)


def config(index):
    """BASELINE.json `configs[index]` -> (ocp, lm option kwargs, default batch)"""
    if index == 0:
        return van_der_pol(20), dict(iterations=10, weights=(2.0, 2.0, 2.0)), 1
    if index == 1:
        return van_der_pol(50), dict(iterations=10, weights=(2.0, 2.0, 2.0)), 1024
    if index == 2:
        return unicycle_time_optimal(50), dict(iterations=10, weights=(2.0, 2.0, 2.0)), 4096
    if index == 3:
        return cart_pole_shooting(100), dict(iterations=10, weights=(10.0, 10.0, 10.0)), 16384
    if index == 4:
        return quadrotor(60), dict(iterations=10, weights=(2.0, 2.0, 2.0)), 65536
    raise ValueError(index)


def instance_data(ocp, batch, seed=1234, offset=0):
    """Seeded per-instance start states x0 [batch, nx] and static references xref [batch, nx] (zeros).

    Instance i of a job always gets the same x0 whatever the rank / shard (`offset` = first global instance index), so that
    a sharded multi-GPU run solves exactly the instances a single-GPU run of the same global batch does.
    """
    nx = ocp.nx
    rng = np.random.Generator(np.random.PCG64(seed))
    total = offset + batch
    if ocp.dynamics == abi.DYN_CART_POLE:
        x0 = np.zeros((total, nx))
        x0[:, 1] = math.pi + rng.uniform(-0.3, 0.3, total)
    elif ocp.dynamics == abi.DYN_UNICYCLE:
        x0 = np.concatenate([rng.uniform(-2.0, 2.0, (total, 2)), rng.uniform(-math.pi, math.pi, (total, 1))], axis=1)
    elif ocp.dynamics == abi.DYN_QUADROTOR:
        x0 = np.zeros((total, nx))
        x0[:, 0:3] = rng.uniform(-1.0, 1.0, (total, 3))
        x0[:, 3:6] = rng.uniform(-0.2, 0.2, (total, 3))
    elif ocp.dynamics == abi.DYN_FREE_SPACE_ROCKET:  # (s, v, m): the mass stays away from the pole of (u - 0.02 v^2) / m
        x0 = np.stack([rng.uniform(-1.0, 1.0, total), rng.uniform(-0.5, 0.5, total), rng.uniform(0.9, 1.3, total)], axis=1)
    else:
        x0 = rng.uniform(-2.0, 2.0, (total, nx))
    x0 = np.ascontiguousarray(x0[offset:], dtype=np.float64)
    xref = np.zeros_like(x0)
    if ocp.dynamics == abi.DYN_FREE_SPACE_ROCKET:
        xref[:, 2] = 0.8  # reference (and, on a grid with a fixed goal, final) mass
    return x0, xref

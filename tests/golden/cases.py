"""Shared case table for the golden fixtures (tests/golden/*.npz) -- used by make_golden.py (generation from the compiled
reference) and by the tests that replay them."""
import numpy as np

from control_box_rst_b200 import _abi as abi
from control_box_rst_b200 import problems

def _timeopt(dynamics, n_grid, dt, u_lb, u_ub, dyn_params):
    """MinimumTime(lsq) on the NonUniformFiniteDifferencesVariableGrid, goal state fixed, dt in [0, 1]."""
    nx, _ = abi.DYN_DIMS[dynamics]
    return problems.make_ocp(grid=abi.GRID_FD_NONUNIFORM_VARDT, dynamics=dynamics, n_grid=n_grid, dt=dt, stage_cost=abi.COST_MINIMUM_TIME_LSQ,
                             u_lb=u_lb, u_ub=u_ub, xf_fixed=(1,) * nx, dt_lb=0.0, dt_ub=1.0, dyn_params=dyn_params)


def _quadratic(grid, dynamics, n_grid, dt, u_lb, u_ub, dyn_params, **kw):
    """Quadratic lsq stage cost (Q = I, R = 0.1 I) and final cost (Qf = 2 I) with control bounds."""
    nx, nu = abi.DYN_DIMS[dynamics]
    return problems.make_ocp(grid=grid, dynamics=dynamics, n_grid=n_grid, dt=dt, q=(1.0,) * nx, r=(0.1,) * nu, qf=(2.0,) * nx, u_lb=u_lb, u_ub=u_ub,
                             dyn_params=dyn_params, **kw)


LINEAR_AB = (-0.3, -2.0, 1.1, -0.5, 0.2, 1.5)  # A = [[-0.3, 1.1], [-2.0, -0.5]] column-major, B = [0.2, 1.5]

# a chain of three lags with feedback, and a two-mass oscillator (positions, velocities) driven at one / both masses
LINEAR_AB_3X1 = (-0.5, 0.2, -0.1, 1.0, -0.4, 0.3, 0.1, 1.2, -0.8) + (0.3, -0.2, 1.0)
_A4 = (0.0, 0.0, -2.0, 1.0, 0.0, 0.0, 1.0, -1.5, 1.0, 0.0, -0.3, 0.1, 0.0, 1.0, 0.1, -0.2)  # column-major
LINEAR_AB_4X1 = _A4 + (0.0, 0.0, 1.0, 0.25)
LINEAR_AB_4X2 = _A4 + (0.0, 0.0, 1.0, 0.25) + (0.1, 0.0, -0.2, 0.8)

# name -> (ocp builder, LM weights, instances in the fixture)
CASES = {
    "vdp20_cn": (lambda: problems.van_der_pol(20), (2.0, 2.0, 2.0), 8),
    "vdp50_cn": (lambda: problems.van_der_pol(50), (2.0, 2.0, 2.0), 16),
    "vdp50_cn_nofinal": (lambda: problems.van_der_pol(50, final_cost=False), (2.0, 2.0, 2.0), 4),
    "vdp30_forward": (lambda: problems.van_der_pol(30, collocation=abi.COLL_FORWARD), (2.0, 2.0, 2.0), 4),
    "vdp30_backward": (lambda: problems.van_der_pol(30, collocation=abi.COLL_BACKWARD), (2.0, 2.0, 2.0), 4),
    "vdp30_midpoint": (lambda: problems.van_der_pol(30, collocation=abi.COLL_MIDPOINT), (2.0, 2.0, 2.0), 4),
    "vdp2_minimal": (lambda: problems.van_der_pol(2), (2.0, 2.0, 2.0), 4),  # smallest legal grid: one interval
    "unicycle30_timeopt": (lambda: problems.unicycle_time_optimal(30), (2.0, 2.0, 2.0), 4),
    "cartpole40_rk4": (lambda: problems.cart_pole_shooting(40), (10.0, 10.0, 10.0), 4),
    "quadrotor12_cn": (lambda: problems.quadrotor(12), (2.0, 2.0, 2.0), 2),
    # final-stage constraints (functions/final_state_constraints.h): TerminalEqualityConstraint -> equality edge on xf,
    # TerminalBall -> the path's only inequality edge (active-set rows); nx = 2, 4, 12 exercise Eigen's reduction order
    "vdp30_terminal_eq": (lambda: problems.van_der_pol(30, terminal_equality=(0.1, -0.05)), (2.0, 2.0, 2.0), 4),
    "vdp30_terminal_ball": (lambda: problems.van_der_pol(30, terminal_ball=((2.0, 0.5), 0.01)), (2.0, 3.0, 2.0), 4),
    "vdp20_terminal_ball_xf_partly_fixed": (lambda: problems.van_der_pol(20, terminal_ball=((1.0, 1.0), 0.04), xf_fixed=(1, 0)),
                                            (2.0, 2.0, 2.0), 4),
    "cartpole20_terminal_ball": (lambda: problems.cart_pole_shooting(20, terminal_ball=((1.0, 2.0, 0.5, 0.25), 0.05)), (10.0, 10.0, 10.0), 4),
    "quadrotor8_terminal_ball": (lambda: problems.quadrotor(8, terminal_ball=(tuple(0.5 + 0.1 * i for i in range(12)), 0.02)), (2.0, 2.0, 2.0), 2),
    # one reference-pinned fixture for every remaining (dynamics, defect, grid) combination the library compiles (kernels_*.cu)
    "duffing20_cn": (lambda: problems.make_ocp(grid=abi.GRID_FD_UNIFORM, dynamics=abi.DYN_DUFFING, n_grid=20, dt=0.1, q=(1, 1), r=(0.1,), qf=(1, 1),
                                               u_lb=(-1.5,), u_ub=(1.5,), dyn_params=(1.0, -1.0, 1.0)), (2.0, 2.0, 2.0), 3),
    "pendulum20_cn": (lambda: problems.make_ocp(grid=abi.GRID_FD_UNIFORM, dynamics=abi.DYN_SIMPLE_PENDULUM, n_grid=20, dt=0.1, q=(1, 1), r=(0.1,),
                                                qf=(1, 1), u_lb=(-2.0,), u_ub=(2.0,), dyn_params=(0.205, 0.34, 9.81, 0.25)), (2.0, 2.0, 2.0), 3),
    "dint20_cn": (lambda: problems.make_ocp(grid=abi.GRID_FD_UNIFORM, dynamics=abi.DYN_DOUBLE_INTEGRATOR, n_grid=20, dt=0.1, q=(1, 1), r=(0.1,),
                                            qf=(1, 1), u_lb=(-1.0,), u_ub=(1.0,), dyn_params=(2.0,)), (2.0, 2.0, 2.0), 3),
    "dint20_forward": (lambda: problems.make_ocp(grid=abi.GRID_FD_UNIFORM, dynamics=abi.DYN_DOUBLE_INTEGRATOR, n_grid=20, dt=0.1,
                                                 collocation=abi.COLL_FORWARD, q=(1, 1), r=(0.1,), qf=(1, 1), u_lb=(-1.0,), u_ub=(1.0,),
                                                 dyn_params=(2.0,)), (2.0, 2.0, 2.0), 3),
    "cartpole20_cn_fd_grid": (lambda: problems.make_ocp(grid=abi.GRID_FD_UNIFORM, dynamics=abi.DYN_CART_POLE, n_grid=20, dt=0.05, q=(1.0,) * 4,
                                                        r=(0.01,), qf=(10.0,) * 4, u_lb=(-20.0,), u_ub=(20.0,),
                                                        dyn_params=(1.0, 0.3, 0.5, 9.81)), (10.0, 10.0, 10.0), 3),
    "unicycle20_cn_fixed_dt": (lambda: problems.make_ocp(grid=abi.GRID_FD_UNIFORM, dynamics=abi.DYN_UNICYCLE, n_grid=20, dt=0.1, q=(1.0,) * 3,
                                                         r=(0.1, 0.1), qf=(5.0,) * 3, u_lb=(-1.0, -1.0), u_ub=(1.0, 1.0)), (2.0, 2.0, 2.0), 3),
    "vdp20_timeopt": (lambda: problems.make_ocp(grid=abi.GRID_FD_NONUNIFORM_VARDT, dynamics=abi.DYN_VAN_DER_POL, n_grid=20, dt=0.1,
                                                stage_cost=abi.COST_MINIMUM_TIME_LSQ, u_lb=(-1.0,), u_ub=(1.0,), xf_fixed=(1, 1), dt_lb=0.0,
                                                dt_ub=1.0, dyn_params=(1.0,)), (2.0, 2.0, 2.0), 3),
    # the remaining systems of nonlinear_benchmark_systems.h (rocket, massless pendulum, toy example, Artstein's circle)
    "rocket20_cn": (lambda: problems.free_space_rocket(20), (2.0, 2.0, 2.0), 3),
    "rocket20_timeopt": (lambda: problems.free_space_rocket_time_optimal(20), (2.0, 2.0, 2.0), 3),
    "massless_pendulum20_cn": (lambda: problems.make_ocp(grid=abi.GRID_FD_UNIFORM, dynamics=abi.DYN_MASSLESS_PENDULUM, n_grid=20, dt=0.1, q=(1, 1),
                                                         r=(0.1,), qf=(1, 1), u_lb=(-2.0,), u_ub=(2.0,), dyn_params=(1.5,)), (2.0, 2.0, 2.0), 3),
    "toy20_cn": (lambda: problems.make_ocp(grid=abi.GRID_FD_UNIFORM, dynamics=abi.DYN_TOY_EXAMPLE, n_grid=20, dt=0.05, q=(1, 1), r=(0.1,), qf=(1, 1),
                                           u_lb=(-2.0,), u_ub=(2.0,), dyn_params=(0.5,)), (2.0, 2.0, 2.0), 3),
    "artstein20_cn": (lambda: problems.make_ocp(grid=abi.GRID_FD_UNIFORM, dynamics=abi.DYN_ARTSTEINS_CIRCLE, n_grid=20, dt=0.1, q=(1, 1), r=(0.1,),
                                                qf=(1, 1), u_lb=(-1.0,), u_ub=(1.0,)), (2.0, 2.0, 2.0), 3),
    # further compiled combinations (kernels_combos_fd.cu, kernels_combos_ms.cu)
    "dint20_timeopt": (lambda: _timeopt(abi.DYN_DOUBLE_INTEGRATOR, 20, 0.1, (-1.0,), (1.0,), (1.0,)), (2.0, 2.0, 2.0), 3),
    "pendulum20_timeopt": (lambda: _timeopt(abi.DYN_SIMPLE_PENDULUM, 20, 0.1, (-2.0,), (2.0,), (0.205, 0.34, 9.81, 0.25)), (2.0, 2.0, 2.0), 3),
    "cartpole20_timeopt": (lambda: _timeopt(abi.DYN_CART_POLE, 20, 0.05, (-20.0,), (20.0,), (1.0, 0.3, 0.5, 9.81)), (10.0, 10.0, 10.0), 3),
    "pendulum20_midpoint": (lambda: _quadratic(abi.GRID_FD_UNIFORM, abi.DYN_SIMPLE_PENDULUM, 20, 0.1, (-2.0,), (2.0,), (0.205, 0.34, 9.81, 0.25),
                                               collocation=abi.COLL_MIDPOINT), (2.0, 2.0, 2.0), 3),
    "cartpole20_forward": (lambda: _quadratic(abi.GRID_FD_UNIFORM, abi.DYN_CART_POLE, 20, 0.05, (-20.0,), (20.0,), (1.0, 0.3, 0.5, 9.81),
                                              collocation=abi.COLL_FORWARD), (10.0, 10.0, 10.0), 3),
    "unicycle20_backward": (lambda: _quadratic(abi.GRID_FD_UNIFORM, abi.DYN_UNICYCLE, 20, 0.1, (-1.0, -1.0), (1.0, 1.0), (),
                                               collocation=abi.COLL_BACKWARD), (2.0, 2.0, 2.0), 3),
    "cartpole20_ms_euler": (lambda: _quadratic(abi.GRID_MULTIPLE_SHOOTING, abi.DYN_CART_POLE, 20, 0.02, (-20.0,), (20.0,), (1.0, 0.3, 0.5, 9.81),
                                               integrator=abi.INT_EULER), (10.0, 10.0, 10.0), 3),
    "unicycle20_ms_rk4": (lambda: _quadratic(abi.GRID_MULTIPLE_SHOOTING, abi.DYN_UNICYCLE, 20, 0.1, (-1.0, -1.0), (1.0, 1.0), ()), (2.0, 2.0, 2.0), 3),
    "duffing20_ms_rk4": (lambda: _quadratic(abi.GRID_MULTIPLE_SHOOTING, abi.DYN_DUFFING, 20, 0.1, (-1.5,), (1.5,), (1.0, -1.0, 1.0)), (2.0, 2.0, 2.0), 3),
    "pendulum20_ms_rk4": (lambda: _quadratic(abi.GRID_MULTIPLE_SHOOTING, abi.DYN_SIMPLE_PENDULUM, 20, 0.1, (-2.0,), (2.0,),
                                             (0.205, 0.34, 9.81, 0.25)), (2.0, 2.0, 2.0), 3),
    "dint20_ms_rk4": (lambda: _quadratic(abi.GRID_MULTIPLE_SHOOTING, abi.DYN_DOUBLE_INTEGRATOR, 20, 0.1, (-1.0,), (1.0,), (2.0,)), (2.0, 2.0, 2.0), 3),
    "dint20_ms_euler": (lambda: _quadratic(abi.GRID_MULTIPLE_SHOOTING, abi.DYN_DOUBLE_INTEGRATOR, 20, 0.1, (-1.0,), (1.0,), (2.0,),
                                           integrator=abi.INT_EULER), (2.0, 2.0, 2.0), 3),
    # LinearStateSpaceModel (linear_benchmark_systems.h:186-214) with a full 2x2 A and a 2x1 B on all three grids
    "linear20_cn": (lambda: _quadratic(abi.GRID_FD_UNIFORM, abi.DYN_LINEAR_2X1, 20, 0.1, (-1.0,), (1.0,), LINEAR_AB), (2.0, 2.0, 2.0), 3),
    "linear20_timeopt": (lambda: _timeopt(abi.DYN_LINEAR_2X1, 20, 0.1, (-1.0,), (1.0,), LINEAR_AB), (2.0, 2.0, 2.0), 3),
    "linear20_ms_rk4": (lambda: _quadratic(abi.GRID_MULTIPLE_SHOOTING, abi.DYN_LINEAR_2X1, 20, 0.1, (-1.0,), (1.0,), LINEAR_AB), (2.0, 2.0, 2.0), 3),
    "linear3_20_cn": (lambda: _quadratic(abi.GRID_FD_UNIFORM, abi.DYN_LINEAR_3X1, 20, 0.1, (-1.0,), (1.0,), LINEAR_AB_3X1), (2.0, 2.0, 2.0), 3),
    "linear3_20_timeopt": (lambda: _timeopt(abi.DYN_LINEAR_3X1, 20, 0.1, (-1.0,), (1.0,), LINEAR_AB_3X1), (2.0, 2.0, 2.0), 3),
    "linear3_20_ms_rk4": (lambda: _quadratic(abi.GRID_MULTIPLE_SHOOTING, abi.DYN_LINEAR_3X1, 20, 0.1, (-1.0,), (1.0,), LINEAR_AB_3X1),
                          (2.0, 2.0, 2.0), 3),
    "linear4_20_cn": (lambda: _quadratic(abi.GRID_FD_UNIFORM, abi.DYN_LINEAR_4X1, 20, 0.1, (-1.0,), (1.0,), LINEAR_AB_4X1), (2.0, 2.0, 2.0), 3),
    "linear4_20_ms_rk4": (lambda: _quadratic(abi.GRID_MULTIPLE_SHOOTING, abi.DYN_LINEAR_4X1, 20, 0.1, (-1.0,), (1.0,), LINEAR_AB_4X1),
                          (2.0, 2.0, 2.0), 3),
    "linear4x2_20_cn": (lambda: _quadratic(abi.GRID_FD_UNIFORM, abi.DYN_LINEAR_4X2, 20, 0.1, (-1.0, -1.0), (1.0, 1.0), LINEAR_AB_4X2),
                        (2.0, 2.0, 2.0), 3),
    "linear4x2_20_timeopt": (lambda: _timeopt(abi.DYN_LINEAR_4X2, 20, 0.1, (-1.0, -1.0), (1.0, 1.0), LINEAR_AB_4X2), (2.0, 2.0, 2.0), 3),
    "tint20_cn": (lambda: _quadratic(abi.GRID_FD_UNIFORM, abi.DYN_TRIPLE_INTEGRATOR, 20, 0.1, (-1.0,), (1.0,), (1.5,)), (2.0, 2.0, 2.0), 3),
    "tint20_timeopt": (lambda: _timeopt(abi.DYN_TRIPLE_INTEGRATOR, 20, 0.1, (-1.0,), (1.0,), (1.0,)), (2.0, 2.0, 2.0), 3),
    "qint20_cn": (lambda: _quadratic(abi.GRID_FD_UNIFORM, abi.DYN_QUAD_INTEGRATOR, 20, 0.1, (-1.0,), (1.0,), (1.5,)), (2.0, 2.0, 2.0), 3),
    "vdp20_ms_euler": (lambda: problems.van_der_pol_shooting(20, integrator=abi.INT_EULER), (2.0, 2.0, 2.0), 3),
    "vdp20_ms_rk4": (lambda: problems.van_der_pol_shooting(20), (2.0, 2.0, 2.0), 3),
}

EVAL_WEIGHTS = (2.0, 3.0, 5.0)


def perturbed_params(ocp, p_init, seed=3):
    """Initial guess pushed off the linear interpolation so that bounds are violated and every Jacobian block is exercised."""
    rng = np.random.default_rng(seed)
    p = p_init + rng.uniform(-0.3, 0.3, p_init.shape)
    return p


def dense_to_csc_values(J, col_ptr, row_idx):
    out = np.zeros(len(row_idx))
    for c in range(len(col_ptr) - 1):
        out[col_ptr[c]:col_ptr[c + 1]] = J[row_idx[col_ptr[c]:col_ptr[c + 1]], c]
    return out


# dynamics linearisation fixtures: name -> (descriptor builder, polynomial?)
LINEARIZE_MODELS = {
    "van_der_pol": (lambda: problems.van_der_pol(5, a=1.3), True),
    "duffing": (lambda: problems.make_ocp(grid=abi.GRID_FD_UNIFORM, dynamics=abi.DYN_DUFFING, n_grid=5, dt=0.1, q=(1, 1), r=(0.1,),
                                          dyn_params=(1.0, -1.0, 1.0)), True),
    "simple_pendulum": (lambda: problems.make_ocp(grid=abi.GRID_FD_UNIFORM, dynamics=abi.DYN_SIMPLE_PENDULUM, n_grid=5, dt=0.1, q=(1, 1), r=(0.1,),
                                                  dyn_params=(0.205, 0.34, 9.81, 0.25)), False),
    "cart_pole": (lambda: problems.cart_pole_shooting(5), False),
    "double_integrator": (lambda: problems.make_ocp(grid=abi.GRID_FD_UNIFORM, dynamics=abi.DYN_DOUBLE_INTEGRATOR, n_grid=5, dt=0.1, q=(1, 1), r=(0.1,),
                                                    dyn_params=(2.0,)), True),
    "unicycle": (lambda: problems.unicycle_time_optimal(5), False),
    "triple_integrator": (lambda: _quadratic(abi.GRID_FD_UNIFORM, abi.DYN_TRIPLE_INTEGRATOR, 5, 0.1, (-1.0,), (1.0,), (1.5,)), True),
    "quad_integrator": (lambda: _quadratic(abi.GRID_FD_UNIFORM, abi.DYN_QUAD_INTEGRATOR, 5, 0.1, (-1.0,), (1.0,), (1.5,)), True),
    "linear_3x1": (lambda: _quadratic(abi.GRID_FD_UNIFORM, abi.DYN_LINEAR_3X1, 5, 0.1, (-1.0,), (1.0,), LINEAR_AB_3X1), True),
    "linear_4x1": (lambda: _quadratic(abi.GRID_FD_UNIFORM, abi.DYN_LINEAR_4X1, 5, 0.1, (-1.0,), (1.0,), LINEAR_AB_4X1), False),
    "linear_4x2": (lambda: _quadratic(abi.GRID_FD_UNIFORM, abi.DYN_LINEAR_4X2, 5, 0.1, (-1.0, -1.0), (1.0, 1.0), LINEAR_AB_4X2), False),
    "linear_2x1": (lambda: _quadratic(abi.GRID_FD_UNIFORM, abi.DYN_LINEAR_2X1, 5, 0.1, (-1.0,), (1.0,), LINEAR_AB), True),
    "free_space_rocket": (lambda: problems.free_space_rocket(5), True),
    "massless_pendulum": (lambda: problems.make_ocp(grid=abi.GRID_FD_UNIFORM, dynamics=abi.DYN_MASSLESS_PENDULUM, n_grid=5, dt=0.1, q=(1, 1), r=(0.1,),
                                                    dyn_params=(1.5,)), False),
    "toy_example": (lambda: problems.make_ocp(grid=abi.GRID_FD_UNIFORM, dynamics=abi.DYN_TOY_EXAMPLE, n_grid=5, dt=0.1, q=(1, 1), r=(0.1,),
                                              dyn_params=(0.5,)), True),
    "artsteins_circle": (lambda: problems.make_ocp(grid=abi.GRID_FD_UNIFORM, dynamics=abi.DYN_ARTSTEINS_CIRCLE, n_grid=5, dt=0.1, q=(1, 1), r=(0.1,)), True),
    "quadrotor": (lambda: problems.quadrotor(5), False),
}
LINEARIZE_POINTS = 6


def linearize_points(ocp, seed=17):
    rng = np.random.default_rng(seed)
    return rng.uniform(-1.5, 1.5, (LINEARIZE_POINTS, ocp.nx)), rng.uniform(-1.0, 1.0, (LINEARIZE_POINTS, ocp.nu))


def hessian_multipliers(ocp, seed=29):
    return np.random.default_rng(seed).uniform(-2.0, 2.0, (LINEARIZE_POINTS, ocp.nx))


# plant fixtures: SimulatedPlant::control at the linearisation points, ClosedLoopControlTask's loop
PLANT_DT = 0.05
CLOSED_LOOP_STEPS = 12


def closed_loop_starts(seed=23, count=6):
    rng = np.random.default_rng(seed)
    return np.vstack([[1.0, 0.5], rng.uniform(-1.5, 1.5, (count - 1, 2))])

"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs.

Tolerances (stated per test):
  * values r: the device evaluates the same IEEE expressions without FMA contraction -> compared at 1e-13 relative
    (bit-exact for the polynomial models; libm sin/cos may differ in the last ulp for the trigonometric ones).
  * Jacobian: central differences with delta=1e-9 amplify 1 ulp to ~1e-7; identical arithmetic gives identical entries for the
    polynomial models (asserted exactly), <= 2e-6 absolute for models calling sin/cos.
  * trajectories after 10 LM iterations, err = |x_gpu - x_oracle|_inf / max(1, |x_oracle|_inf).  The reference's Jacobians are
    central differences with delta = 1e-9, so its OWN result moves under a 1-ulp change of the start state (measured with the
    oracle, DESIGN.md "FD-noise floor"): Van der Pol ~5e-8 median / 2e-7 max, unicycle time-optimal 6e-8 median / 1e-5 max,
    cart-pole shooting 5e-5 median / 7e-5 max.  The bar per case is therefore
        Van der Pol  : 95% of the instances <= 1e-6 (north_star's bar), all <= 1e-4
        unicycle     : 95% <= 2e-5, all <= 1e-3
        cart-pole    : 95% <= 5e-4, all <= 5e-3
    and chi2 relative <= 1e-6 (<= 1e-4 for cart-pole).
"""
import numpy as np
import pytest

from control_box_rst_b200 import _abi as abi
from control_box_rst_b200 import problems, solver

pytestmark = pytest.mark.gpu


def _csc_to_dense(ocp, jac_values):
    d = solver.dims_of(ocp)
    col_ptr, row_idx = solver.jacobian_pattern(ocp)
    J = np.zeros((d.m, d.n_params))
    for c in range(d.n_params):
        J[row_idx[col_ptr[c]:col_ptr[c + 1]], c] = jac_values[col_ptr[c]:col_ptr[c + 1]]
    return J


CASES = [
    ("vdp20_cn", lambda: problems.van_der_pol(20), True),
    ("vdp50_cn", lambda: problems.van_der_pol(50), True),
    ("vdp50_nofinal", lambda: problems.van_der_pol(50, final_cost=False), True),
    ("vdp30_forward", lambda: problems.van_der_pol(30, collocation=abi.COLL_FORWARD), True),
    ("vdp30_backward", lambda: problems.van_der_pol(30, collocation=abi.COLL_BACKWARD), True),
    ("vdp30_midpoint", lambda: problems.van_der_pol(30, collocation=abi.COLL_MIDPOINT), True),
    ("unicycle30_timeopt", lambda: problems.unicycle_time_optimal(30), False),
    ("cartpole40_rk4", lambda: problems.cart_pole_shooting(40), False),
    ("quadrotor12", lambda: problems.quadrotor(12), False),
    # final-stage constraints: equality edge / the inequality edge with its active-set rows (final_state_constraints.h)
    ("vdp30_terminal_eq", lambda: problems.van_der_pol(30, terminal_equality=(0.1, -0.05)), True),
    ("vdp30_terminal_ball", lambda: problems.van_der_pol(30, terminal_ball=((2.0, 0.5), 0.01)), True),
    ("vdp20_terminal_ball_xf_partly_fixed", lambda: problems.van_der_pol(20, terminal_ball=((1.0, 1.0), 0.04), xf_fixed=(1, 0)), True),
    ("cartpole20_terminal_ball", lambda: problems.cart_pole_shooting(20, terminal_ball=((1.0, 2.0, 0.5, 0.25), 0.05)), False),
]


@pytest.mark.parametrize("name,make,exact", CASES, ids=[c[0] for c in CASES])
def test_values_and_jacobian_match_oracle(oracle, name, make, exact):
    ocp = make()
    B = 8
    x0, xref = problems.instance_data(ocp, B, seed=7)
    lm = solver.BatchedLevenbergMarquardt(ocp, B)
    lm.set_problem_data(x0, xref)
    lm.initialize_trajectories()
    # perturb the initial guess so that bounds are violated and every Jacobian block is exercised
    rng = np.random.default_rng(3)
    p0 = lm.get_params()
    for i in range(B):
        np.testing.assert_allclose(p0[i], oracle.initial_params(ocp, x0[i], xref[i]), rtol=0, atol=1e-14)
    p0 = p0 + rng.uniform(-0.3, 0.3, p0.shape)
    if ocp.grid == abi.GRID_FD_NONUNIFORM_VARDT:
        _, _, dt_idx = solver.vertex_indices(ocp)
        p0[:, dt_idx] = np.abs(p0[:, dt_idx]) + 0.05
    lm.set_params(p0)
    w = (2.0, 3.0, 5.0)
    values, jac = lm.evaluate(w)
    after = lm.get_params()
    for i in range(B):
        v_o, J_o, _, after_o = oracle.evaluate(ocp, x0[i], xref[i], p0[i], w)
        J = _csc_to_dense(ocp, jac[i])
        if exact:
            assert np.array_equal(values[i], v_o), np.abs(values[i] - v_o).max()
            assert np.array_equal(J, J_o), np.abs(J - J_o).max()
            assert np.array_equal(after[i], after_o)
        else:
            np.testing.assert_allclose(values[i], v_o, rtol=1e-13, atol=1e-13)
            np.testing.assert_allclose(J, J_o, rtol=0, atol=2e-6 * max(1.0, np.abs(J_o).max()))
            np.testing.assert_allclose(after[i], after_o, rtol=0, atol=1e-15 * max(1.0, np.abs(after_o).max()) * 4)
    lm.clear()


def _traj_err(p, p_ref):
    return np.abs(p - p_ref).max(axis=1) / np.maximum(1.0, np.abs(p_ref).max(axis=1))


@pytest.mark.parametrize("threads", [1, 2, 8], ids=["T1", "T2", "T8"])
@pytest.mark.parametrize("name,make,weights,B,tol95,tolmax,tolchi2", [
    ("vdp20", lambda: problems.van_der_pol(20), (2.0, 2.0, 2.0), 64, 1e-6, 1e-4, 1e-6),
    ("vdp50", lambda: problems.van_der_pol(50), (2.0, 2.0, 2.0), 256, 1e-6, 1e-4, 1e-6),
    ("unicycle30", lambda: problems.unicycle_time_optimal(30), (2.0, 2.0, 2.0), 32, 2e-5, 1e-3, 1e-6),
    ("cartpole40", lambda: problems.cart_pole_shooting(40), (10.0, 10.0, 10.0), 32, 5e-4, 5e-3, 1e-4),
    ("vdp30_terminal_eq", lambda: problems.van_der_pol(30, terminal_equality=(0.1, -0.05)), (2.0, 2.0, 2.0), 64, 1e-6, 1e-4, 1e-6),
    ("vdp30_terminal_ball", lambda: problems.van_der_pol(30, terminal_ball=((2.0, 0.5), 0.01)), (2.0, 3.0, 2.0), 64, 1e-6, 1e-4, 1e-6),
], ids=["vdp20", "vdp50", "unicycle30", "cartpole40", "vdp30_terminal_eq", "vdp30_terminal_ball"])
def test_solve_matches_oracle(oracle, name, make, weights, B, tol95, tolmax, tolchi2, threads):
    ocp = make()
    x0, xref = problems.instance_data(ocp, B, seed=11)
    lm = solver.BatchedLevenbergMarquardt(ocp, B)
    lm.setIterations(10)
    lm.setPenaltyWeights(*weights)
    lm.set_threads_per_instance(threads)
    lm.set_problem_data(x0, xref)
    lm.initialize_trajectories()
    status, chi2 = lm.solve(new_run=True)
    p = lm.get_params()
    opts = abi.LmOptions.defaults(iterations=10, weights=weights)
    p_o, chi2_o, status_o, _ = oracle.solve_batch(ocp, opts, x0, xref, threads=8)
    err = _traj_err(p, p_o)
    print(name, "traj err percentiles 50/90/99/100:", np.percentile(err, [50, 90, 99, 100]), "status agree", (status == status_o).mean())
    assert (err <= tol95).mean() >= 0.95
    assert err.max() <= tolmax
    np.testing.assert_allclose(chi2, chi2_o, rtol=tolchi2)
    assert (status == status_o).mean() >= 0.95
    lm.clear()


# ---------------------------------------------------------------------------------------------------------------------------
# against the golden vectors generated from the compiled reference itself
# ---------------------------------------------------------------------------------------------------------------------------
import os  # noqa: E402
import sys  # noqa: E402

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import cases  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
POLYNOMIAL = {"vdp20_cn", "vdp50_cn", "vdp50_cn_nofinal", "vdp30_forward", "vdp30_backward", "vdp30_midpoint", "vdp2_minimal",
              "vdp30_terminal_eq", "vdp30_terminal_ball", "vdp20_terminal_ball_xf_partly_fixed",
              "duffing20_cn", "dint20_cn", "dint20_forward", "vdp20_timeopt", "vdp20_ms_euler", "vdp20_ms_rk4",
              "rocket20_cn", "rocket20_timeopt", "toy20_cn", "artstein20_cn",
              "dint20_timeopt", "duffing20_ms_rk4", "dint20_ms_rk4", "dint20_ms_euler",
              "linear20_cn", "linear20_timeopt", "linear20_ms_rk4", "linear3_20_cn", "linear3_20_timeopt", "linear3_20_ms_rk4",
              "tint20_cn", "tint20_timeopt", "qint20_cn"}
GOLD_TOL = {"unicycle30_timeopt": (1e-3, 1e-4), "cartpole40_rk4": (5e-3, 1e-3), "quadrotor12_cn": (1e-3, 1e-4),
            "cartpole20_terminal_ball": (5e-3, 1e-3), "quadrotor8_terminal_ball": (1e-3, 1e-4),
            "pendulum20_cn": (1e-3, 1e-4), "cartpole20_cn_fd_grid": (5e-3, 1e-3), "unicycle20_cn_fixed_dt": (1e-3, 1e-4),
            "vdp20_timeopt": (2e-5, 1e-5),
            "rocket20_cn": (2e-5, 1e-5), "rocket20_timeopt": (2e-5, 1e-5), "toy20_cn": (2e-5, 1e-5), "massless_pendulum20_cn": (1e-3, 1e-4),
            "pendulum20_timeopt": (1e-3, 1e-4), "cartpole20_timeopt": (5e-3, 1e-3), "pendulum20_midpoint": (1e-3, 1e-4),
            "cartpole20_forward": (5e-3, 1e-3), "unicycle20_backward": (1e-3, 1e-4), "cartpole20_ms_euler": (5e-3, 1e-3),
            "unicycle20_ms_rk4": (1e-3, 1e-4), "duffing20_ms_rk4": (1e-5, 1e-6), "pendulum20_ms_rk4": (1e-3, 1e-4),
            "linear20_timeopt": (2e-5, 1e-5), "linear20_ms_rk4": (1e-5, 1e-6), "linear3_20_timeopt": (2e-5, 1e-5),
            "linear4x2_20_timeopt": (2e-5, 1e-5), "linear4x2_20_cn": (1e-6, 1e-6), "linear4_20_cn": (1e-6, 1e-6),
            "linear3_20_ms_rk4": (1e-5, 1e-6), "linear4_20_ms_rk4": (1e-5, 1e-6),
            "tint20_timeopt": (2e-5, 1e-5), "tint20_cn": (1e-6, 1e-6), "qint20_cn": (1e-6, 1e-6),
            "artstein20_cn": (1e-6, 1e-6),  # same optimum to 1e-8 in chi2; the intermediate iterates of this poorly controllable system differ by 1.3e-7
            "vdp20_ms_rk4": (1e-5, 1e-6)}  # (trajectory, chi2); RK4 shooting: four nested evaluations per defect amplify the FD noise (1.5e-6 observed)


@pytest.mark.parametrize("name", list(cases.CASES))
def test_matches_reference_golden(name):
    make, weights, B = cases.CASES[name]
    ocp = make()
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    x0, xref = gold["x0"], gold["xref"]
    lm = solver.BatchedLevenbergMarquardt(ocp, B)
    lm.set_problem_data(x0, xref)
    lm.initialize_trajectories()
    np.testing.assert_allclose(lm.get_params(), gold["p_init"], rtol=0, atol=1e-14)
    p = gold["p_init"].copy()
    p[0] = gold["p_eval"]
    lm.set_params(p)
    values, jac = lm.evaluate(cases.EVAL_WEIGHTS)
    after = lm.get_params()
    if name in POLYNOMIAL:
        assert np.array_equal(values[0], gold["values"])
        assert np.array_equal(jac[0], gold["jac_values"])
        assert np.array_equal(after[0], gold["p_after"])
    else:
        np.testing.assert_allclose(values[0], gold["values"], rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(jac[0], gold["jac_values"], rtol=0, atol=2e-6 * max(1.0, np.abs(gold["jac_values"]).max()))
    lm.setIterations(10)
    lm.setPenaltyWeights(*weights)
    lm.initialize_trajectories()
    status, chi2 = lm.solve(new_run=True)
    traj_tol, chi2_tol = GOLD_TOL.get(name, (1e-6, 1e-8))
    err = _traj_err(lm.get_params(), gold["p_final"])
    assert err.max() <= traj_tol, err
    np.testing.assert_allclose(chi2, gold["chi2"], rtol=chi2_tol)
    # status = Converged iff the last trial step was rejected (rho <= 0) or ||r|| <= 1e-5 (levenberg_marquardt_sparse.cpp:218).  Once an
    # instance sits on the FD-noise floor (chi2 no longer moves in 9 digits) the sign of the last gain ratio is one realisation of the
    # Jacobian noise (DESIGN.md "FD-noise floor"), so the flag is only compared for instances that are still descending.
    full_trace = lm.chi2_trace()
    at_floor = np.abs(full_trace[:, -1] - full_trace[:, -2]) <= 1e-9 * np.abs(full_trace[:, -1])
    assert np.array_equal(status[~at_floor], gold["status"][~at_floor])
    assert name not in POLYNOMIAL or name in GOLD_TOL or np.array_equal(status, gold["status"])
    # per-iteration chi2 of instance 0 against the reference's event trace (values at every Jacobian evaluation)
    trace = full_trace[0]
    ref_chi2 = gold["trace_chi2"][gold["trace_types"] == 0]
    np.testing.assert_allclose(trace[:len(ref_chi2)], ref_chi2, rtol=max(chi2_tol, 1e-7))
    lm.clear()


def test_closed_loop_matches_reference_controller():
    """configs[0] plumbing: 15 warm-started MPC steps (structure kept, x0 replaced, weights reset each step) against the
    reference's own StructuredOptimalControlProblem::compute loop; plant = one RK4 step of the same dynamics."""
    gold = np.load(os.path.join(GOLDEN, "vdp20_closed_loop.npz"))
    ocp = problems.van_der_pol(20)
    lm = solver.BatchedLevenbergMarquardt(ocp, 1)
    lm.setIterations(10)

    def f(x, u):
        return np.array([x[1], -1.0 * (x[0] * x[0] - 1) * x[1] - x[0] + u[0]])

    x = np.array([1.0, 0.5])
    dt = ocp.dt_ref
    for s in range(15):
        lm.step(x[None, :], None, cold_start=(s == 0))
        u = lm.get_first_controls()[0]
        np.testing.assert_allclose(u, gold["u"][s], rtol=0, atol=2e-6)
        k1 = f(x, u) * dt
        k2 = f(x + k1 / 2.0, u) * dt
        k3 = f(x + k2 / 2.0, u) * dt
        k4 = f(x + k3, u) * dt
        x = x + (k1 + 2.0 * k2 + 2.0 * k3 + k4) / 6.0
        np.testing.assert_allclose(x, gold["x"][s + 1], rtol=0, atol=2e-6)
    lm.clear()


def test_partially_fixed_goal_state(oracle):
    """PartiallyFixedVectorVertex xf (setXfFixed): fixed components are no parameters, take their value from the reference and
    never move; parameter order and results match the oracle."""
    ocp = problems.van_der_pol(25)
    ocp.xf_fixed[1] = 1
    B = 16
    x0, _ = problems.instance_data(ocp, B, seed=9)
    xref = np.tile(np.array([0.3, -0.2]), (B, 1))
    lm = solver.BatchedLevenbergMarquardt(ocp, B)
    assert lm.dims.n_params == 24 * 3 - 1
    lm.setIterations(8)
    lm.set_problem_data(x0, xref)
    lm.initialize_trajectories()
    status, chi2 = lm.solve()
    p_o, chi2_o, status_o, _ = oracle.solve_batch(ocp, abi.LmOptions.defaults(iterations=8), x0, xref, threads=4)
    assert _traj_err(lm.get_params(), p_o).max() <= 1e-6
    np.testing.assert_allclose(chi2, chi2_o, rtol=1e-7)
    lm.clear()


def test_weight_adaptation_over_warm_started_solves(oracle):
    """new_run = false multiplies the penalty weights by the adaptation factors up to their maxima
    (levenberg_marquardt_sparse.cpp:83-86, 270-287)."""
    ocp = problems.van_der_pol(20)
    B = 4
    x0, xref = problems.instance_data(ocp, B, seed=13)
    lm = solver.BatchedLevenbergMarquardt(ocp, B)
    lm.setIterations(3)
    lm.setPenaltyWeights(2.0, 2.0, 2.0)
    lm.setWeightAdapation(2.0, 2.0, 3.0, 10.0, 10.0, 10.0)
    lm.set_problem_data(x0, xref)
    lm.initialize_trajectories()
    opts = abi.LmOptions.defaults(iterations=3, weights=(2.0, 2.0, 2.0), factors=(2.0, 2.0, 3.0), maxima=(10.0, 10.0, 10.0))
    chi2_seq = []
    for s in range(4):
        _, chi2 = lm.solve(new_run=(s == 0))
        chi2_seq.append(chi2)
    p = lm.get_params()
    for i in range(B):
        p_o, chi2_o = oracle.solve_sequence(ocp, opts, x0[i], xref[i], n_solves=4)
        np.testing.assert_allclose([c[i] for c in chi2_seq], chi2_o, rtol=1e-7)
        np.testing.assert_allclose(p[i], p_o, rtol=0, atol=1e-6)
    lm.clear()


def test_full_size_properties():
    """BASELINE.json configs[1] at the north-star batch (4096 x Van der Pol N=50), size-independent properties:
    determinism (bit-identical reruns), batch independence (an instance's result does not depend on its neighbours),
    monotone accepted chi2, first-order optimality of the penalty problem, agreement of the thread mappings within the FD-noise
    floor."""
    ocp, kw, _ = problems.config(1)
    B = 4096
    x0, xref = problems.instance_data(ocp, B, seed=1235)
    lm = solver.BatchedLevenbergMarquardt(ocp, B)
    lm.setIterations(10)
    lm.set_problem_data(x0, xref)

    def run(T):
        lm.set_threads_per_instance(T)
        lm.initialize_trajectories()
        status, chi2 = lm.solve()
        return lm.get_params(), chi2, status, lm.chi2_trace()

    p8, c8, s8, tr8 = run(8)
    p8b, c8b, s8b, _ = run(8)
    assert np.array_equal(p8, p8b) and np.array_equal(c8, c8b) and np.array_equal(s8, s8b)
    assert np.all(np.diff(tr8, axis=1) <= 0)  # chi2 only changes on accepted steps, which decrease it
    assert np.all((s8 == abi.STATUS_CONVERGED) | (s8 == abi.STATUS_EARLY_TERMINATED)) and np.all(np.isfinite(p8))
    p1, c1, _, _ = run(1)
    err = _traj_err(p8, p1)
    assert (err <= 1e-6).mean() >= 0.99 and err.max() <= 1e-4
    np.testing.assert_allclose(c8, c1, rtol=1e-6)
    # batch independence: the first 96 instances alone
    small = solver.BatchedLevenbergMarquardt(ocp, 96)
    small.setIterations(10)
    small.set_threads_per_instance(8)
    small.set_problem_data(x0[:96], xref[:96])
    small.initialize_trajectories()
    small.solve()
    assert np.array_equal(small.get_params(), p8[:96])
    small.clear()
    # stationarity of the converged points: the LM gradient J^T r is small (evaluate re-linearises at the solution)
    lm.set_threads_per_instance(8)
    values, jac = lm.evaluate((2.0, 2.0, 2.0))
    col_ptr, row_idx = solver.jacobian_pattern(ocp)
    grad = np.zeros((B, lm.dims.n_params))
    for c in range(lm.dims.n_params):
        sl = slice(col_ptr[c], col_ptr[c + 1])
        grad[:, c] = np.einsum("bi,bi->b", jac[:, sl], values[:, row_idx[sl]])
    assert np.percentile(np.abs(grad).max(axis=1), 95) <= 1e-3
    lm.clear()


def test_headline_batch_every_instance_against_the_checkers(oracle):
    """The benchmark's own batch -- all 4096 Van der Pol N=50 instances bench.py solves (same seed, same start states), 10 LM iterations --
    against the oracle and, where oracle/_ref/libcorbo_ref.so is present, against the compiled reference, instance by instance:
    chi2 within 1e-8, status identical, trajectories within north_star's 1e-6 relative for every instance except a listed handful
    (8 of 4096 observed, worst 7.9e-6).  Each exception must be an instance on which the CHECKER ITSELF is that sensitive: its own result
    moves by at least half the device's deviation under random changes of its start state by up to 8 ulp (central differences with delta = 1e-9
    amplify last-bit differences of the linear solver; the count of exceptions is bounded at 0.5 % of the batch)."""
    from oracle import bindings

    ocp, kw, _ = problems.config(1)
    B = 4096
    x0, xref = problems.instance_data(ocp, B, seed=1234 + 1)  # bench.py: seed = 1234 + config
    lm = solver.BatchedLevenbergMarquardt(ocp, B)
    lm.setIterations(kw["iterations"])
    lm.setPenaltyWeights(*kw["weights"])
    lm.set_problem_data(x0, xref)
    lm.initialize_trajectories()
    status, chi2 = lm.solve(new_run=True)
    p = lm.get_params()
    lm.clear()
    opts = abi.LmOptions.defaults(iterations=kw["iterations"], weights=kw["weights"])
    checkers = [("oracle", oracle)] + ([("reference", bindings.Reference())] if bindings.Reference.available() else [])
    for label, chk in checkers:
        p_c, c_c, s_c, _ = chk.solve_batch(ocp, opts, x0, xref, threads=8)
        err = _traj_err(p, p_c)
        bad = np.nonzero(err > 1e-6)[0]
        print(f"headline batch vs {label}: trajectory error median {np.median(err):.2e} p99 {np.percentile(err, 99):.2e} max {err.max():.2e}; "
              f"chi2 rel max {np.abs(chi2 / c_c - 1).max():.2e}; status agree {(status == s_c).mean():.4f}; exceptions {bad.tolist()}")
        np.testing.assert_allclose(chi2, c_c, rtol=1e-8)
        assert np.array_equal(status, s_c), label
        assert bad.size <= B // 200, (label, bad.tolist())
        if bad.size:
            # 32 random perturbations of each exceptional start state by up to +-8 ulp per component: these instances are bimodal -- the
            # checker lands on one of two outcomes a fixed distance apart (one accept/reject or bound-activation decision flips)
            rng, n_pert = np.random.default_rng(0), 32
            own = np.zeros(bad.size)
            for _ in range(n_pert):
                xk = x0[bad].copy()
                steps = rng.integers(-8, 9, xk.shape)
                for _s in range(8):
                    xk = np.where(steps > _s, np.nextafter(xk, np.inf), np.where(-steps > _s, np.nextafter(xk, -np.inf), xk))
                p_u, _, _, _ = chk.solve_batch(ocp, opts, xk, xref[bad], threads=8)
                own = np.maximum(own, _traj_err(p_u, p_c[bad]))
            print(f"  exceptions vs {label}: device deviation {err[bad]}, the checker's own movement under +-8 ulp {own}")
            assert np.all(err[bad] <= 2.0 * own), (label, bad.tolist(), err[bad].tolist(), own.tolist())


# (config index, trajectory tolerance of the oracle spot check, chi2 tolerance): the per-config FD-noise floors of the module docstring
FULL_SIZE = [(2, 1e-3, 1e-6), (3, 5e-3, 1e-4), (4, 1e-3, 1e-4)]


@pytest.mark.parametrize("index,tol_traj,tol_chi2", FULL_SIZE, ids=["unicycle_4096", "cartpole_16384", "quadrotor_65536"])
def test_full_size_properties_other_configs(oracle, index, tol_traj, tol_chi2):
    """BASELINE.json configs[2..4] at their full batch (unicycle time-optimal 4096 x N=50, cart-pole shooting 16384 x N=100, quadrotor
    65536 x N=60; fp64 -- the reference has no fp32 path): determinism, batch independence (the first 64 instances alone give the same
    bits), monotone accepted chi2, admissible status everywhere, and a spot check of 32 instances spread over the batch against
    the oracle."""
    ocp, kw, B = problems.config(index)
    iters = 10
    x0, xref = problems.instance_data(ocp, B, seed=77)
    lm = solver.BatchedLevenbergMarquardt(ocp, B)
    lm.setIterations(iters)
    lm.setPenaltyWeights(*kw["weights"])
    lm.set_problem_data(x0, xref)

    def run():
        lm.initialize_trajectories()
        status, chi2 = lm.solve(new_run=True)
        return lm.get_params(), chi2.copy(), status.copy()

    p, c, st = run()
    p2, c2, st2 = run()
    assert np.array_equal(p, p2) and np.array_equal(c, c2) and np.array_equal(st, st2)
    del p2
    tr = lm.chi2_trace()
    lm.clear()
    assert np.all(np.isfinite(p)) and np.all(np.isfinite(c))
    assert np.all(np.diff(tr, axis=1) <= 0)
    assert np.all((st == abi.STATUS_CONVERGED) | (st == abi.STATUS_EARLY_TERMINATED))
    small = solver.BatchedLevenbergMarquardt(ocp, 64)
    small.setIterations(iters)
    small.setPenaltyWeights(*kw["weights"])
    small.set_problem_data(x0[:64], xref[:64])
    small.initialize_trajectories()
    small.solve(new_run=True)
    assert np.array_equal(small.get_params(), p[:64])
    small.clear()
    idx = np.linspace(0, B - 1, 32).astype(int)
    opts = abi.LmOptions.defaults(iterations=iters, weights=kw["weights"])
    p_o, c_o, _, _ = oracle.solve_batch(ocp, opts, x0[idx], xref[idx], threads=8)
    err = _traj_err(p[idx], p_o)
    print(f"config {index}: B={B} spot-check traj err max {err.max():.2e}, chi2 rel {np.abs(c[idx] / c_o - 1).max():.2e}")
    assert err.max() <= tol_traj
    np.testing.assert_allclose(c[idx], c_o, rtol=tol_chi2)


def test_ragged_batch_sizes(oracle):
    """batches that do not fill a warp / a block, including a single instance"""
    ocp = problems.van_der_pol(10)
    for B in (1, 5, 33):
        x0, xref = problems.instance_data(ocp, B, seed=B)
        lm = solver.BatchedLevenbergMarquardt(ocp, B)
        lm.setIterations(5)
        lm.set_problem_data(x0, xref)
        lm.initialize_trajectories()
        status, chi2 = lm.solve()
        p_o, chi2_o, _, _ = oracle.solve_batch(ocp, abi.LmOptions.defaults(iterations=5), x0, xref, threads=2)
        assert _traj_err(lm.get_params(), p_o).max() <= 1e-6
        np.testing.assert_allclose(chi2, chi2_o, rtol=1e-8)
        lm.clear()


def test_large_block_pipeline_agrees_with_fused_kernel_and_oracle(oracle):
    """The 12-state quadrotor (16 x 16 stage blocks) runs the warp-per-instance pipeline (lm_pipeline.cuh); the fused kernel is kept
    behind set_threads_per_instance(-1).  Both must agree with each other and with the oracle within the FD-noise floor, the
    pipeline must be deterministic, and instances must not depend on their neighbours."""
    ocp = problems.quadrotor(16)
    B = 96
    x0, xref = problems.instance_data(ocp, B, seed=13)
    opts = abi.LmOptions.defaults(iterations=6)

    def run(T, batch=B):
        lm = solver.BatchedLevenbergMarquardt(ocp, batch)
        lm.setIterations(6)
        lm.set_threads_per_instance(T)
        lm.set_problem_data(x0[:batch], xref[:batch])
        lm.initialize_trajectories()
        status, chi2 = lm.solve(new_run=True)
        out = lm.get_params(), chi2, status, lm.chi2_trace(), lm.statistics()
        lm.clear()
        return out

    p_pipe, c_pipe, s_pipe, tr_pipe, st_pipe = run(0)
    p_pipe2, c_pipe2, _, _, _ = run(0)
    assert np.array_equal(p_pipe, p_pipe2) and np.array_equal(c_pipe, c_pipe2)
    p_small, _, _, _, _ = run(0, batch=40)
    assert np.array_equal(p_small, p_pipe[:40])
    p_fused, c_fused, _, tr_fused, st_fused = run(-1)
    assert _traj_err(p_pipe, p_fused).max() <= 1e-3
    np.testing.assert_allclose(c_pipe, c_fused, rtol=1e-4)
    np.testing.assert_allclose(tr_pipe[:, :3], tr_fused[:, :3], rtol=1e-6)  # the first iterations are not yet noise-dominated
    assert np.all(np.diff(tr_pipe, axis=1) <= 0)
    p_o, c_o, _, _ = oracle.solve_batch(ocp, opts, x0, xref, threads=8)
    assert _traj_err(p_pipe, p_o).max() <= 1e-3
    np.testing.assert_allclose(c_pipe, c_o, rtol=1e-4)
    assert st_pipe["relinearizations"].min() >= 1 and st_pipe["inner_passes"].min() >= 6


def test_quadrotor_fp32_variant_against_the_fp64_oracle(oracle):
    """BASELINE.json configs[4] names fp32.  The reduced-precision variant of the pipeline (Jacobian columns by fp32 central differences
    with delta = 2^-10, normal equations / Cholesky / substitutions in fp32; parameters, trial residuals and LM control in fp64) has no
    reference counterpart: SURVEY.md section 8d judges it against the fp64 oracle at 1e-3 relative on the trajectories.  Asserted here
    for EVERY instance, together with chi2 at 1e-3, determinism, batch independence, and the refusal of structures without such a path."""
    ocp = problems.quadrotor(16)
    B = 256
    x0, xref = problems.instance_data(ocp, B, seed=21)
    opts = abi.LmOptions.defaults(iterations=10)

    def run(precision, batch=B):
        lm = solver.BatchedLevenbergMarquardt(ocp, batch)
        lm.setIterations(10)
        lm.set_precision(precision)
        lm.set_problem_data(x0[:batch], xref[:batch])
        lm.initialize_trajectories()
        status, chi2 = lm.solve(new_run=True)
        out = lm.get_params(), chi2, status, lm.chi2_trace()
        lm.clear()
        return out

    p32, c32, s32, tr32 = run("f32")
    p32b, c32b, _, _ = run("f32")
    assert np.array_equal(p32, p32b) and np.array_equal(c32, c32b)
    p_small, _, _, _ = run("f32", batch=40)
    assert np.array_equal(p_small, p32[:40])
    assert np.all(np.isfinite(p32)) and np.all(np.diff(tr32, axis=1) <= 0)
    p_o, c_o, _, _ = oracle.solve_batch(ocp, opts, x0, xref, threads=8)
    err = _traj_err(p32, p_o)
    p64, c64, _, _ = run("f64")
    print(f"quadrotor fp32 vs fp64 oracle: trajectory error median {np.median(err):.2e} p99 {np.percentile(err, 99):.2e} max {err.max():.2e}; "
          f"chi2 rel max {np.abs(c32 / c_o - 1).max():.2e}; fp64 device vs oracle max {_traj_err(p64, p_o).max():.2e}")
    assert err.max() <= 1e-3
    np.testing.assert_allclose(c32, c_o, rtol=1e-3)
    # structures that do not run the warp-cooperative pipeline have no fp32 path: refused, never silently fp64
    vdp = solver.BatchedLevenbergMarquardt(problems.van_der_pol(20), 8)
    with pytest.raises(solver.B200SqpError) as info:
        vdp.set_precision("f32")
    assert info.value.code == abi.ERR_UNSUPPORTED
    vdp.clear()


@pytest.mark.parametrize("make", [lambda: problems.van_der_pol(12), lambda: problems.van_der_pol(11, collocation=abi.COLL_FORWARD, final_cost=False),
                                  lambda: problems.make_ocp(grid=abi.GRID_FD_UNIFORM, dynamics=abi.DYN_DUFFING, n_grid=9, dt=0.1, q=(1.0, 2.0), r=(0.3,),
                                                            qf=(2.0, 1.0), u_lb=(-1.0,), u_ub=(1.0,), dyn_params=(1.0, -1.0, 1.0)),
                                  lambda: problems.van_der_pol(50)],
                         ids=["vdp12_fd", "vdp11_forward_nofinal", "duffing9_fd", "vdp50_fd"])
def test_time_varying_reference_matches_the_checkers(oracle, make):
    """Non-static state reference (b200sqp_set_reference_trajectory; ReferenceTrajectoryInterface::isStatic() == false): the cost edge of
    grid point k measures x_k against getReferenceCached(k), the cold start is the reference trajectory itself.  Initial guess, values,
    Jacobian and drift bit-identical to the checker (the compiled reference where present -- the CPU suite shows oracle == reference bit
    for bit on the same cases), solves within the polynomial-model bars for T = 1 and the widest thread mapping; switching back to a
    static reference works; the shooting grid refuses (see include/b200sqp.h)."""
    from oracle import bindings

    ocp = make()
    N, nx = ocp.n_grid, ocp.nx
    B = 40
    rng = np.random.default_rng(5)
    x0, _ = problems.instance_data(ocp, B, seed=3)
    t = np.linspace(0.0, 1.0, N)
    xref = np.stack([np.stack([0.5 * np.sin(2.0 * t + 0.3 * i) + 0.1 * j for j in range(nx)], axis=1) for i in range(B)])  # [B, N, nx]
    chk = bindings.Reference() if bindings.Reference.available() else oracle
    chk.set_xref_points(N)
    try:
        lm = solver.BatchedLevenbergMarquardt(ocp, B)
        lm.set_problem_data(x0, None)
        lm.set_reference_trajectory(xref)
        lm.initialize_trajectories()
        p_init = lm.get_params()
        for i in range(4):
            assert np.array_equal(p_init[i], chk.initial_params(ocp, x0[i], xref[i]))
        p = p_init + rng.uniform(-0.2, 0.2, p_init.shape)
        lm.set_params(p)
        w = (2.0, 3.0, 4.0)
        values, jac = lm.evaluate(w)
        after = lm.get_params()
        for i in range(4):
            v_c, J_c, _, a_c = chk.evaluate(ocp, x0[i], xref[i], p[i], w)
            assert np.array_equal(values[i], v_c), np.abs(values[i] - v_c).max()
            assert np.array_equal(_csc_to_dense(ocp, jac[i]), J_c)
            assert np.array_equal(after[i], a_c)
        opts = abi.LmOptions.defaults(iterations=8)
        p_c, c_c, s_c, _ = chk.solve_batch(ocp, opts, x0, xref, threads=4)
        for T in (1, 8):
            lm.setIterations(8)
            lm.set_threads_per_instance(T)
            lm.initialize_trajectories()
            status, chi2 = lm.solve(new_run=True)
            err = _traj_err(lm.get_params(), p_c)
            assert (err <= 1e-6).mean() >= 0.95 and err.max() <= 1e-4, (T, err.max())
            np.testing.assert_allclose(chi2, c_c, rtol=1e-6)
        # back to a static reference (the last row): same as never having set a trajectory
        chk.set_xref_points(0)
        last = np.ascontiguousarray(xref[:, -1, :])
        lm.set_problem_data(x0, last)
        lm.initialize_trajectories()
        _, chi2_s = lm.solve(new_run=True)
        p_s, c_s, _, _ = chk.solve_batch(ocp, opts, x0, last, threads=4)
        assert _traj_err(lm.get_params(), p_s).max() <= 1e-4 and np.abs(lm.get_params() - p_c).max() > 1e-3
        np.testing.assert_allclose(chi2_s, c_s, rtol=1e-6)
        lm.clear()
    finally:
        chk.set_xref_points(0)
    ms = solver.BatchedLevenbergMarquardt(problems.van_der_pol_shooting(10), 4)
    with pytest.raises(solver.B200SqpError) as info:
        ms.set_reference_trajectory(np.zeros((4, 10, 2)))
    assert info.value.code == abi.ERR_UNSUPPORTED
    ms.clear()


def test_full_weight_matrices_match_the_checkers(oracle):
    """Non-diagonal Q / R / Qf (upper Cholesky square roots, dense Jacobian blocks of the cost edges): values, Jacobian and drift
    bit-identical to the checker for up to three states, solves within the bars of the polynomial models; a combination without a
    compiled dense-cost kernel is refused."""
    import sys as _sys
    _sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_oracle_vs_reference import DENSE_CASES
    from oracle import bindings

    chk = bindings.Reference() if bindings.Reference.available() else oracle
    for name, (make, xr, exact) in DENSE_CASES.items():
        ocp = make()
        B = 24
        x0, _ = problems.instance_data(ocp, B, seed=2)
        xref = np.tile(np.array(xr, dtype=float), (B, 1))
        rng = np.random.default_rng(1)
        lm = solver.BatchedLevenbergMarquardt(ocp, B)
        lm.set_problem_data(x0, xref)
        lm.initialize_trajectories()
        p = lm.get_params() + rng.uniform(-0.2, 0.2, (B, lm.dims.n_params))
        lm.set_params(p)
        w = (2.0, 3.0, 4.0)
        values, jac = lm.evaluate(w)
        after = lm.get_params()
        for i in range(3):
            v_c, J_c, _, a_c = chk.evaluate(ocp, x0[i], xref[i], p[i], w)
            J = _csc_to_dense(ocp, jac[i])
            if exact:
                assert np.array_equal(values[i], v_c), (name, np.abs(values[i] - v_c).max())
                assert np.array_equal(J, J_c), (name, np.abs(J - J_c).max())
                assert np.array_equal(after[i], a_c), name
            else:
                np.testing.assert_allclose(values[i], v_c, rtol=1e-13, atol=1e-13)
                np.testing.assert_allclose(J, J_c, rtol=0, atol=2e-6 * max(1.0, np.abs(J_c).max()))
        opts = abi.LmOptions.defaults(iterations=8)
        p_c, c_c, _, _ = chk.solve_batch(ocp, opts, x0, xref, threads=4)
        for T in (1, 4):
            lm.setIterations(8)
            lm.set_threads_per_instance(T)
            lm.initialize_trajectories()
            _, chi2 = lm.solve(new_run=True)
            err = _traj_err(lm.get_params(), p_c)
            assert err.max() <= (1e-5 if exact else 1e-3), (name, T, err.max())
            np.testing.assert_allclose(chi2, c_c, rtol=1e-6)
        lm.clear()
    Q4 = np.eye(4) + 0.1 * np.ones((4, 4))
    with pytest.raises(solver.B200SqpError) as info:
        solver.BatchedLevenbergMarquardt(problems.cart_pole_shooting(10, q_full=Q4), 4)
    assert info.value.code == abi.ERR_UNSUPPORTED


def test_dt_equality_edges_match_the_checkers(oracle):
    """NonUniformFiniteDifferencesVariableGrid::setDtEqConstraint(true) (TwoScalarEqualEdge, edges/misc_edges.h:40-67): consecutive dt
    vertices are tied by equality edges, which couples the dt slots of neighbouring stage blocks on the device (sub-diagonal blocks one
    column wider).  Values, Jacobian and drift bit-identical to the checker for the polynomial models (incl. the bound rows of the dt
    vertices, finished one interval late), solves for T = 1 and the widest thread mapping (chunk boundaries cut through the dt
    coupling), N = 3 (a single equality edge) included."""
    import sys as _sys
    _sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_oracle_vs_reference import DT_EQ_CASES
    from oracle import bindings

    chk = bindings.Reference() if bindings.Reference.available() else oracle
    for name, make in DT_EQ_CASES.items():
        ocp = make()
        B = 40
        x0, xref = problems.instance_data(ocp, B, seed=4)
        rng = np.random.default_rng(2)
        lm = solver.BatchedLevenbergMarquardt(ocp, B)
        lm.set_problem_data(x0, xref)
        lm.initialize_trajectories()
        p = lm.get_params() + rng.uniform(-0.2, 0.2, (B, lm.dims.n_params))
        dt_idx = solver.vertex_indices(ocp)[2]
        p[:, dt_idx] = np.abs(p[:, dt_idx]) + 0.05
        p[::3, dt_idx[0]] = 1.3  # some dt vertices beyond their upper bound: active bound rows
        lm.set_params(p)
        w = (2.0, 3.0, 4.0)
        values, jac = lm.evaluate(w)
        after = lm.get_params()
        exact = ocp.dynamics != abi.DYN_UNICYCLE
        for i in range(4):
            v_c, J_c, _, a_c = chk.evaluate(ocp, x0[i], xref[i], p[i], w)
            J = _csc_to_dense(ocp, jac[i])
            if exact:
                assert np.array_equal(values[i], v_c), (name, np.abs(values[i] - v_c).max())
                assert np.array_equal(J, J_c), (name, np.abs(J - J_c).max())
                assert np.array_equal(after[i], a_c), (name, np.abs(after[i] - a_c).max())
            else:
                np.testing.assert_allclose(values[i], v_c, rtol=1e-13, atol=1e-13)
                np.testing.assert_allclose(J, J_c, rtol=0, atol=2e-6 * max(1.0, np.abs(J_c).max()))
        opts = abi.LmOptions.defaults(iterations=6, weights=w)
        p_c, c_c, _, _ = chk.solve_batch(ocp, opts, x0, xref, threads=4)
        for T in (1, 8):
            lm.setIterations(6)
            lm.setPenaltyWeights(*w)
            lm.set_threads_per_instance(T)
            lm.initialize_trajectories()
            _, chi2 = lm.solve(new_run=True)
            err = _traj_err(lm.get_params(), p_c)
            print(f"{name} T={T}: trajectory error max {err.max():.2e}, chi2 rel {np.abs(chi2 / c_c - 1).max():.2e}")
            assert err.max() <= 1e-3 and np.percentile(err, 90) <= 1e-4, (name, T, err.max())
            np.testing.assert_allclose(chi2, c_c, rtol=1e-4)
        lm.clear()


@pytest.mark.parametrize("name,make", [
    ("unicycle_timeopt", lambda: problems.unicycle_time_optimal(24)),
    ("unicycle_timeopt_dteq", lambda: problems.unicycle_time_optimal(16, dt_eq_constraint=True)),
    ("dint_timeopt", lambda: problems.make_ocp(grid=abi.GRID_FD_NONUNIFORM_VARDT, dynamics=abi.DYN_DOUBLE_INTEGRATOR, n_grid=14, dt=0.1,
                                               stage_cost=abi.COST_MINIMUM_TIME_LSQ, u_lb=(-1.0,), u_ub=(1.0,), xf_fixed=(1, 0), dt_lb=0.0, dt_ub=1.0,
                                               dyn_params=(1.0,))),
], ids=["unicycle", "unicycle_dteq", "dint_partial_goal"])
@pytest.mark.parametrize("threads", [1, 8], ids=["T1", "T8"])
def test_time_optimal_feature_set_is_bit_identical_to_the_general_one(name, make, threads):
    """Time-optimal structures without state bounds / final-stage constraint run a compile-time feature set of their own (FeatTimeOpt,
    lm_device.cuh): it only removes code of absent features, so parameters, chi2 and statuses equal the general set's bit for bit"""
    ocp = make()
    B = 96
    x0, xref = problems.instance_data(ocp, B, seed=21)
    out = []
    for general in (True, False):
        lm = solver.BatchedLevenbergMarquardt(ocp, B)
        lm.setIterations(8)
        lm.set_threads_per_instance(threads)
        lm.set_feature_set(general)
        lm.set_problem_data(x0, xref)
        lm.initialize_trajectories()
        status, chi2 = lm.solve(new_run=True)
        out.append((lm.get_params(), chi2, status))
        lm.clear()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][2], out[1][2])

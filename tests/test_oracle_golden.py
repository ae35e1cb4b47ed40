"""Pins the CPU oracle (oracle/sqp_oracle.cpp) against the golden vectors generated from the compiled, unmodified reference
(tests/golden/make_golden.py) and against the reference's own known-answer solver tests
(optimization/test/test_levenberg_marquardt_sparse.cpp:72-296)."""
import os
import sys

import numpy as np
import pytest

from control_box_rst_b200 import _abi as abi
from oracle import bindings

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import cases  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# models whose edge functions only use + - * / : the oracle reproduces the reference bit for bit
POLYNOMIAL = {"vdp20_cn", "vdp50_cn", "vdp50_cn_nofinal", "vdp30_forward", "vdp30_backward", "vdp30_midpoint", "vdp2_minimal",
              "vdp30_terminal_eq", "vdp30_terminal_ball", "vdp20_terminal_ball_xf_partly_fixed",
              "duffing20_cn", "dint20_cn", "dint20_forward", "vdp20_timeopt", "vdp20_ms_euler", "vdp20_ms_rk4",
              "rocket20_cn", "rocket20_timeopt", "toy20_cn", "artstein20_cn",
              "dint20_timeopt", "duffing20_ms_rk4", "dint20_ms_rk4", "dint20_ms_euler",
              "linear20_cn", "linear20_timeopt", "linear20_ms_rk4", "linear3_20_cn", "linear3_20_timeopt", "linear3_20_ms_rk4",
              "tint20_cn", "tint20_timeopt", "qint20_cn"}
# FD-noise floor of the reference algorithm per case: trajectory tolerance for 10 LM iterations
TRAJ_TOL = {"unicycle30_timeopt": 1e-3, "cartpole40_rk4": 5e-3, "quadrotor12_cn": 1e-3, "cartpole20_terminal_ball": 5e-3,
            "quadrotor8_terminal_ball": 1e-3, "pendulum20_cn": 1e-3, "cartpole20_cn_fd_grid": 5e-3, "unicycle20_cn_fixed_dt": 1e-3,
            "vdp20_timeopt": 2e-5,
            # rational / bilinear right-hand sides: values and Jacobians bit-exact, the solve meets the FD-noise floor within 10 iterations
            "rocket20_cn": 2e-5, "rocket20_timeopt": 2e-5, "toy20_cn": 2e-5,
            "pendulum20_timeopt": 1e-3, "cartpole20_timeopt": 5e-3, "cartpole20_forward": 5e-3, "cartpole20_ms_euler": 5e-3,
            "unicycle20_ms_rk4": 1e-3, "duffing20_ms_rk4": 1e-5, "linear20_timeopt": 2e-5, "tint20_timeopt": 2e-5}  # time-optimal: polynomial values/Jacobians (bit-exact), but the free dt makes the solve noise-sensitive


@pytest.mark.parametrize("name", list(cases.CASES))
def test_indices_and_dims(oracle, name):
    ocp = cases.CASES[name][0]()
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = oracle.dims(ocp)
    got = np.array([d.n_params, d.m_lsq, d.m_eq, d.m_ineq, d.m_bounds, d.nnz_jacobian, d.nnz_hessian_upper, d.algorithmic_bytes_per_iteration])
    assert np.array_equal(got, gold["dims"])
    for a, key in zip(oracle.vertex_indices(ocp), ("x_idx", "u_idx", "dt_idx")):
        assert np.array_equal(a, gold[key])
    assert np.array_equal(oracle.edge_table(ocp, 0), gold["edges_lsq"])
    assert np.array_equal(oracle.edge_table(ocp, 1), gold["edges_eq"])
    if "edges_ineq" in gold:
        assert np.array_equal(oracle.edge_table(ocp, 2), gold["edges_ineq"])


@pytest.mark.parametrize("name", list(cases.CASES))
def test_initial_guess_values_and_jacobian(oracle, name):
    ocp = cases.CASES[name][0]()
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    x0, xref = gold["x0"], gold["xref"]
    for i in range(len(x0)):
        np.testing.assert_allclose(oracle.initial_params(ocp, x0[i], xref[i]), gold["p_init"][i], rtol=0, atol=1e-15)
    values, J, pattern, after = oracle.evaluate(ocp, x0[0], xref[0], gold["p_eval"], cases.EVAL_WEIGHTS)
    jac_values = cases.dense_to_csc_values(J, gold["col_ptr"], gold["row_idx"])
    assert int(pattern.sum()) == len(gold["row_idx"])
    if name in POLYNOMIAL:
        assert np.array_equal(values, gold["values"])
        assert np.array_equal(jac_values, gold["jac_values"])
        assert np.array_equal(after, gold["p_after"])
    else:
        np.testing.assert_allclose(values, gold["values"], rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(jac_values, gold["jac_values"], rtol=0, atol=2e-6 * max(1.0, np.abs(gold["jac_values"]).max()))
        np.testing.assert_allclose(after, gold["p_after"], rtol=0, atol=1e-14)


@pytest.mark.parametrize("name", list(cases.CASES))
def test_lm_solve(oracle, name):
    make, weights, B = cases.CASES[name]
    ocp = make()
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    opts = abi.LmOptions.defaults(iterations=10, weights=weights)
    p, chi2, status, _ = oracle.solve_batch(ocp, opts, gold["x0"], gold["xref"], threads=2)
    err = np.abs(p - gold["p_final"]).max(axis=1) / np.maximum(1.0, np.abs(gold["p_final"]).max(axis=1))
    tol = TRAJ_TOL.get(name, 1e-6)
    assert err.max() <= tol, err
    np.testing.assert_allclose(chi2, gold["chi2"], rtol=1e-4 if name in TRAJ_TOL else 1e-8)
    assert np.array_equal(status, gold["status"])
    # event-level trace of instance 0: same sequence of Jacobian evaluations / increments / accepts / rejects
    tr = oracle.trace(ocp, opts, gold["x0"][0], gold["xref"][0])
    types = np.array([e[0] for e in tr["events"]], np.int32)
    assert np.array_equal(types, gold["trace_types"])
    t_chi2 = np.array([e[1] for e in tr["events"]])
    jac_events = types == bindings.EV_JACOBIAN
    np.testing.assert_allclose(t_chi2[jac_events], gold["trace_chi2"][jac_events], rtol=1e-4 if name in TRAJ_TOL else 1e-7)


def test_reference_known_answers(oracle):
    table = np.load(os.path.join(GOLDEN, "lm_known_answers.npz"))["table"]
    for row in table:
        cid, stage, n, tol = int(row[0]), int(row[1]), int(row[2]), row[3]
        x_ref, expected = row[4:4 + n], row[7:7 + n]
        x, exp, t = oracle.known_answer(cid, stage)
        assert t == tol and np.array_equal(exp, expected)
        assert np.abs(x - expected).max() <= tol          # the reference test's own EXPECT_NEAR
        assert np.abs(x_ref - expected).max() <= tol      # ... which the compiled reference meets too
        assert np.abs(x - x_ref).max() <= max(2e-3, tol)  # and both land on the same penalty-method optimum

"""Bit-exact hypergraph indexing: dimensions, vertex indices, edge indices and the CSC pattern of the combined Jacobian from the
C ABI (host-only part, no GPU) against the golden fixtures generated from the compiled reference, and against the oracle."""
import os
import sys

import numpy as np
import pytest

from control_box_rst_b200 import _abi as abi
from control_box_rst_b200 import problems, solver

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import cases  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", list(cases.CASES))
def test_dims_and_indices_match_reference_golden(name):
    ocp = cases.CASES[name][0]()
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = solver.dims_of(ocp)
    got = np.array([d.n_params, d.m_lsq, d.m_eq, d.m_ineq, d.m_bounds, d.nnz_jacobian, d.nnz_hessian_upper, d.algorithmic_bytes_per_iteration])
    assert np.array_equal(got, gold["dims"]), (got, gold["dims"])
    x_idx, u_idx, dt_idx = solver.vertex_indices(ocp)
    assert np.array_equal(x_idx, gold["x_idx"]) and np.array_equal(u_idx, gold["u_idx"]) and np.array_equal(dt_idx, gold["dt_idx"])
    col_ptr, row_idx = solver.jacobian_pattern(ocp)
    assert np.array_equal(col_ptr, gold["col_ptr"]) and np.array_equal(row_idx, gold["row_idx"])
    assert col_ptr.dtype == np.int32 and x_idx.dtype == np.int32


@pytest.mark.parametrize("name", list(cases.CASES))
def test_edge_indices_match_reference_golden(name):
    ocp = cases.CASES[name][0]()
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    e = solver.edge_indices(ocp)
    # reference edge tables: rows [dim, edge_idx, n_vertices, vertex_idx0..3] in creation order
    lsq, eq = gold["edges_lsq"], gold["edges_eq"]
    # final-stage constraint edge: last equality edge (TerminalEqualityConstraint) or the only inequality edge (TerminalBall)
    feq, fineq = solver.final_constraint_indices(ocp)
    K = ocp.n_grid - 1
    if ocp.final_constraint == abi.FINAL_CONSTRAINT_EQUALITY:
        assert len(eq) == K + 1 and eq[-1, 1] == feq and eq[-1, 0] == ocp.nx and eq[-1, 2] == 1
        eq = eq[:-1]
    else:
        assert feq == -1
    if "edges_ineq" in gold and len(gold["edges_ineq"]):
        assert ocp.final_constraint == abi.FINAL_CONSTRAINT_BALL and gold["edges_ineq"].shape[0] == 1
        assert gold["edges_ineq"][0, 1] == fineq and gold["edges_ineq"][0, 0] == 1
    else:
        assert fineq == -1
    assert np.array_equal(eq[:, 1], e["dynamics"])
    assert np.all(eq[:, 0] == ocp.nx)
    mine = []
    for k in range(K):
        if e["state_cost"][k] >= 0:
            mine.append(e["state_cost"][k])
        if e["control_cost"][k] >= 0:
            mine.append(e["control_cost"][k])
        for r in range(2):
            if e["dt_cost"][k, r] >= 0:
                mine.append(e["dt_cost"][k, r])
    if e["final_cost"] >= 0:
        mine.append(e["final_cost"])
    assert np.array_equal(np.array(mine), lsq[:, 1])


def test_structure_matches_oracle_on_sweep(oracle):
    """every grid kind x a sweep of horizon lengths, fixed-goal masks and bound patterns"""
    inf = abi.CORBO_INF_DBL
    variants = []
    for n in (2, 3, 7, 20):
        variants.append(problems.van_der_pol(n))
        o = problems.van_der_pol(n)
        o.xf_fixed[0] = 1
        variants.append(o)
        o = problems.van_der_pol(n)
        o.xf_fixed[0] = o.xf_fixed[1] = 1
        variants.append(o)
        o = problems.van_der_pol(n)
        o.x_lb[1], o.x_ub[0] = -3.0, 4.0
        o.u_lb[0], o.u_ub[0] = -inf, inf
        variants.append(o)
        variants.append(problems.unicycle_time_optimal(max(n, 3)))
        variants.append(problems.cart_pole_shooting(max(n, 3)))
    o = problems.unicycle_time_optimal(9)
    o.xf_fixed[1] = 0
    variants.append(o)
    # final-stage constraints, incl. on a partially / fully fixed goal (fully fixed: the reference creates no final-stage edge)
    for n in (2, 5, 20):
        variants.append(problems.van_der_pol(n, terminal_equality=(0.1, -0.2)))
        variants.append(problems.van_der_pol(n, terminal_ball=((1.0, 2.0), 0.1)))
        variants.append(problems.van_der_pol(n, terminal_ball=((1.0, 2.0), 0.1), xf_fixed=(0, 1)))
        variants.append(problems.van_der_pol(n, terminal_equality=(0.1, -0.2), xf_fixed=(1, 1)))
        variants.append(problems.cart_pole_shooting(max(n, 3), terminal_ball=((1.0,) * 4, 0.5)))
    for ocp in variants:
        d, od = solver.dims_of(ocp), oracle.dims(ocp)
        for f in ("n_params", "m_lsq", "m_eq", "m_ineq", "m_bounds", "nnz_jacobian", "nnz_hessian_upper", "algorithmic_bytes_per_iteration"):
            assert getattr(d, f) == getattr(od, f), f
        for a, b in zip(solver.vertex_indices(ocp), oracle.vertex_indices(ocp)):
            assert np.array_equal(a, b)
        x0 = np.full(ocp.nx, 0.25)
        _, _, pattern, _ = oracle.evaluate(ocp, x0)
        col_ptr, row_idx = solver.jacobian_pattern(ocp)
        mine = np.zeros_like(pattern)
        for c in range(d.n_params):
            mine[row_idx[col_ptr[c]:col_ptr[c + 1]], c] = True
        assert np.array_equal(mine, pattern)


def test_survey_numbers():
    """VdP N=20 -> n=57, eq=38, bounds=19, nnzJ=300; N=50 -> n=147, nnzJ=780, nnzH(upper)=582, 49 984 B."""
    d20, d50 = solver.dims_of(problems.van_der_pol(20)), solver.dims_of(problems.van_der_pol(50))
    assert (d20.n_params, d20.m_eq, d20.m_bounds, d20.nnz_jacobian) == (57, 38, 19, 300)
    assert (d50.n_params, d50.nnz_jacobian, d50.nnz_hessian_upper, d50.algorithmic_bytes_per_iteration) == (147, 780, 582, 49984)
    dc = solver.dims_of(problems.cart_pole_shooting(100))
    assert (dc.n_params, dc.m_lsq, dc.m_eq, dc.m_bounds) == (495, 499, 396, 99)
    du = solver.dims_of(problems.unicycle_time_optimal(30))
    assert du.m_lsq == 58  # the dt-cost edge is created twice per interval (nlp_functions.cpp:91-107)


def test_full_weight_matrices_structure_and_refusals(oracle):
    """Full (non-diagonal) Q / R / Qf in the descriptor: same dimensions and Jacobian pattern as with diagonal weights (the cost edges'
    blocks are dense in the pattern either way); the reference's broken branch (non-diagonal Q with a zero state reference,
    quadratic_cost.cpp:112) and matrices that are not positive definite are refused."""
    import numpy as np
    from control_box_rst_b200 import _abi as abi, problems, solver

    Q = np.array([[2.0, 0.3], [0.3, 1.0]])
    dense, diag = problems.van_der_pol(12, q_full=Q, qf_full=2.0 * Q), problems.van_der_pol(12)
    d1, d2 = solver.dims_of(dense), solver.dims_of(diag)
    for f in ("n_params", "m_lsq", "m_eq", "m_ineq", "m_bounds", "nnz_jacobian", "nnz_hessian_upper"):
        assert getattr(d1, f) == getattr(d2, f) == getattr(oracle.dims(dense), f), f
    for a, b in zip(solver.jacobian_pattern(dense), solver.jacobian_pattern(diag)):
        assert np.array_equal(a, b)
    zero_ref = problems.van_der_pol(12, q_full=Q)
    zero_ref.zero_x_ref = 1
    with pytest.raises(solver.B200SqpError) as info:
        solver.dims_of(zero_ref)
    assert info.value.code == abi.ERR_UNSUPPORTED
    with pytest.raises(solver.B200SqpError) as info:
        solver.dims_of(problems.van_der_pol(12, q_full=np.array([[1.0, 2.0], [2.0, 1.0]])))
    assert info.value.code == abi.ERR_INVALID
    solver.dims_of(problems.van_der_pol(12, q_full=np.diag([2.0, 0.5])))  # full but diagonal: the diagonal branch, zero reference allowed


def test_dt_equality_edges_structure(oracle):
    """setDtEqConstraint(true) on the non-uniform grid: N-2 more equality rows, each right after the dynamics rows of its interval, two
    more Jacobian entries per row; dimensions, indices and the CSC pattern equal the oracle's (which the CPU suite pins to the compiled
    reference)."""
    import sys as _sys
    _sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_oracle_vs_reference import DT_EQ_CASES

    for name, make in DT_EQ_CASES.items():
        ocp = make()
        d, d_o = solver.dims_of(ocp), oracle.dims(ocp)
        for f in ("n_params", "m_lsq", "m_eq", "m_ineq", "m_bounds", "nnz_jacobian", "nnz_hessian_upper", "algorithmic_bytes_per_iteration"):
            assert getattr(d, f) == getattr(d_o, f), (name, f)
        K, nx = ocp.n_grid - 1, ocp.nx
        idx = solver.dt_equality_indices(ocp)
        dyn = solver.edge_indices(ocp)["dynamics"]
        assert idx[0] == -1
        for k in range(1, K):
            assert idx[k] == dyn[k] + nx and (k + 1 == K or dyn[k + 1] == idx[k] + 1)
        # pattern: the oracle's stored-entry mask at any point
        x0, xref = problems.instance_data(ocp, 1, seed=1)
        _, _, P_o, _ = oracle.evaluate(ocp, x0[0], xref[0], None, (2.0, 2.0, 2.0))
        col_ptr, row_idx = solver.jacobian_pattern(ocp)
        pat = np.zeros_like(P_o)
        for c in range(d.n_params):
            pat[row_idx[col_ptr[c]:col_ptr[c + 1]], c] = True
        assert np.array_equal(pat, P_o), name
    plain = problems.unicycle_time_optimal(12)
    assert np.all(solver.dt_equality_indices(plain) == -1)
    fixed = problems.van_der_pol(10)
    fixed.dt_eq_constraint = 1
    with pytest.raises(solver.B200SqpError):
        solver.dims_of(fixed)


def test_cart_pole_parameters_are_the_reference_constants():
    """CartPoleSystem has no parameter setters (nonlinear_benchmark_systems.h:337-365): the descriptor may carry its constants or zeros,
    anything else is refused instead of silently running the default plant"""
    ocp = problems.cart_pole_shooting(20)
    assert solver.dims_of(ocp).n_params > 0
    for i in range(4):
        ocp.dyn_params[i] = 0.0
    assert solver.dims_of(ocp).n_params > 0
    ocp.dyn_params[1] = 0.4
    with pytest.raises(solver.B200SqpError) as e:
        solver.dims_of(ocp)
    assert e.value.code == abi.ERR_UNSUPPORTED

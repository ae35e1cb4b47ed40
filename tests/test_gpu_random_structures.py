"""Device-side counterpart of tests/test_oracle_vs_reference.py::test_random_structures_*: seeded random OCPs over the polynomial /
rational models (random model, grid, collocation or integrator, horizon, step, cost weights, bounds, partially fixed goal, penalty
weights), restricted to the (dynamics, defect, grid) combinations compiled into libb200sqp.so, through the C ABI on the GPU:

  * dimensions, vertex and edge indices, the initial guess, the value vector, the combined Jacobian in the reference's CSC
    pattern and the parameter drift of the in-place differences are BIT-IDENTICAL to the checker's;
  * three LM iterations from six random start states, for every compiled variant of the LM kernel the structure can run --
    cooperating threads per instance T = 1 and the widest compiled T, general feature set and (where eligible) lean feature
    set -- give the checker's event counts per instance (Jacobian evaluations, factorisations that produced a trial point,
    rejected steps) and its status; its chi2 at the first Jacobian evaluation to 1e-13 (same point, other summation order), at the
    second to 1e-9, and from then on -- like the final trajectories and chi2 -- within the bars of the CPU test (1e-5 / 1e-6) or twice
    the checker's OWN sensitivity to a change of the start state by 1..4 ulp measured in the same test, whichever is larger: the
    central differences (delta = 1e-9) amplify the last-bit differences between two valid elimination orders of the linear solver by
    5e8 from the second re-linearisation on (worst structure: the checker moves by 2.9e-6 in chi2 under a 1-ulp change).

The checker is the unmodified reference compiled from /root/reference (oracle/_ref/libcorbo_ref.so, which travels to the GPU box)
when present, else the oracle port, which the CPU suite proves bit-identical to it on the same 16 structures.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from control_box_rst_b200 import _abi as abi  # noqa: E402
from control_box_rst_b200 import problems, solver  # noqa: E402
from oracle import bindings  # noqa: E402
from test_oracle_vs_reference import _random_ocp  # noqa: E402

pytestmark = pytest.mark.gpu

N_STRUCTURES = 16


def compiled(ocp):
    """is the (dynamics, defect, grid) combination in the closed kernel registry?  (b200sqp_create answers UNSUPPORTED before it
    looks for a device, so this also works on a host without a GPU)"""
    try:
        lm = solver.BatchedLevenbergMarquardt(ocp, 1)
    except solver.B200SqpError as e:
        if e.code == abi.ERR_UNSUPPORTED:
            return False
        if e.code == abi.ERR_NO_DEVICE:
            return True
        raise
    lm.clear()
    return True


def random_compiled_ocp(seed):
    """the first compiled member of the seeded random sequence of structures"""
    rng = np.random.default_rng(1000 + seed)
    for _ in range(200):
        ocp = _random_ocp(rng)
        if compiled(ocp):
            return ocp, rng
    raise AssertionError("no compiled combination in 200 draws")


@pytest.fixture(scope="module")
def checker(oracle):
    return bindings.Reference() if bindings.Reference.available() else oracle


def lean_eligible(ocp):
    inf = abi.CORBO_INF_DBL
    no_x_bounds = all(ocp.x_lb[j] <= -inf and ocp.x_ub[j] >= inf for j in range(ocp.nx))
    return (ocp.stage_cost == abi.COST_QUADRATIC_LSQ and ocp.final_constraint == 0 and no_x_bounds
            and not any(ocp.xf_fixed[j] for j in range(ocp.nx)))


@pytest.mark.parametrize("seed", range(N_STRUCTURES))
def test_random_structure_values_jacobian_drift_bit_identical(checker, seed):
    ocp, rng = random_compiled_ocp(seed)
    d_c, d_d = checker.dims(ocp), solver.dims_of(ocp)
    for f in ("n_params", "m_lsq", "m_eq", "m_ineq", "m_bounds", "nnz_jacobian", "nnz_hessian_upper", "algorithmic_bytes_per_iteration"):
        assert getattr(d_c, f) == getattr(d_d, f), f
    for a, b in zip(checker.vertex_indices(ocp), solver.vertex_indices(ocp)):
        assert np.array_equal(a, b)
    B = 4
    x0, xref = problems.instance_data(ocp, B, seed=seed)
    lm = solver.BatchedLevenbergMarquardt(ocp, B)
    lm.set_problem_data(x0, xref)
    lm.initialize_trajectories()
    p_init = lm.get_params()
    p = np.zeros_like(p_init)
    for i in range(B):
        np.testing.assert_allclose(p_init[i], checker.initial_params(ocp, x0[i], xref[i]), rtol=0, atol=1e-15)
        p[i] = p_init[i] + rng.uniform(-0.3, 0.3, p_init[i].shape)
    if ocp.grid == abi.GRID_FD_NONUNIFORM_VARDT:
        dt_idx = solver.vertex_indices(ocp)[2]
        p[:, dt_idx] = np.abs(p[:, dt_idx]) + 0.05
    weights = tuple(rng.uniform(1.0, 10.0, 3))
    lm.set_params(p)
    values, jac = lm.evaluate(weights)
    after = lm.get_params()
    col_ptr, row_idx = solver.jacobian_pattern(ocp)
    for i in range(B):
        v_c, J_c, P_c, a_c = checker.evaluate(ocp, x0[i], xref[i], p[i], weights)
        pat = np.zeros_like(P_c)
        J = np.zeros_like(J_c)
        for c in range(d_d.n_params):
            sl = slice(col_ptr[c], col_ptr[c + 1])
            pat[row_idx[sl], c] = True
            J[row_idx[sl], c] = jac[i, sl]
        assert np.array_equal(pat, P_c), "stored-entry pattern of the combined Jacobian"
        assert np.array_equal(values[i], v_c), np.abs(values[i] - v_c).max()
        assert np.array_equal(J, J_c), np.abs(J - J_c).max()
        assert np.array_equal(after[i], a_c), np.abs(after[i] - a_c).max()
    lm.clear()


def _event_counts(tr):
    types = np.array([e[0] for e in tr["events"]])
    chi2 = np.array([e[1] for e in tr["events"]])
    return dict(jacobians=int((types == bindings.EV_JACOBIAN).sum()), increments=int((types == bindings.EV_INCREMENT).sum()),
                restores=int((types == bindings.EV_RESTORE).sum()), chi2_at_jacobian=chi2[types == bindings.EV_JACOBIAN])


@pytest.mark.parametrize("seed", range(N_STRUCTURES))
def test_random_structure_first_lm_iterations_every_kernel_variant(checker, seed):
    ocp, rng = random_compiled_ocp(seed)
    weights = tuple(rng.uniform(1.0, 10.0, 3))
    B, iters = 6, 3
    x0, xref = problems.instance_data(ocp, B, seed=seed)
    opts = abi.LmOptions.defaults(iterations=iters, weights=weights)
    p_c, c_c, s_c, _ = checker.solve_batch(ocp, opts, x0, xref, threads=2)
    # the checker's own sensitivity: start states moved by 1..4 ulp
    floor_traj, floor_chi2, xk = 0.0, 0.0, x0
    for _ in range(4):
        xk = np.nextafter(xk, np.inf)
        p_u, c_u, _, _ = checker.solve_batch(ocp, opts, xk, xref, threads=2)
        floor_traj = max(floor_traj, (np.abs(p_u - p_c).max(axis=1) / np.maximum(1.0, np.abs(p_c).max(axis=1))).max())
        floor_chi2 = max(floor_chi2, np.abs(c_u / c_c - 1.0).max())
    tol_traj, tol_chi2 = max(1e-5, 2.0 * floor_traj), max(1e-6, 2.0 * floor_chi2)
    events = [_event_counts(checker.trace(ocp, opts, x0[i], xref[i])) for i in range(B)]
    variants = [(T, general) for T in (1, 8) for general in ((True, False) if lean_eligible(ocp) else (True,))]
    for T, general in variants:
        lm = solver.BatchedLevenbergMarquardt(ocp, B)
        lm.setIterations(iters)
        lm.setPenaltyWeights(*weights)
        lm.set_threads_per_instance(T)  # clamped to the widest variant compiled for the combination
        lm.set_feature_set(general)
        lm.set_problem_data(x0, xref)
        lm.initialize_trajectories()
        status, chi2 = lm.solve(new_run=True)
        p = lm.get_params()
        st = lm.statistics()
        trace = lm.chi2_trace()
        lm.clear()
        tag = f"seed {seed} T={T} {'general' if general else 'lean'}"
        err = np.abs(p - p_c).max(axis=1) / np.maximum(1.0, np.abs(p_c).max(axis=1))
        assert err.max() <= tol_traj, (tag, err, floor_traj)
        np.testing.assert_allclose(chi2, c_c, rtol=tol_chi2, err_msg=tag)
        assert np.array_equal(status, s_c), tag
        for i in range(B):
            ev = events[i]
            assert st["relinearizations"][i] == ev["jacobians"], (tag, i)
            assert st["rejects"][i] == ev["restores"], (tag, i)
            # a factorisation whose step is below eps2 produces no trial point (levenberg_marquardt_sparse.cpp:151-154)
            assert st["inner_passes"][i] >= ev["increments"], (tag, i)
            # chi2 at every Jacobian evaluation = chi2 of the accepted iterates, in order: the device trace holds chi2 after every
            # outer iteration, which changes exactly when a step was accepted
            accepted = [trace[i, 0]] + [trace[i, k] for k in range(1, iters + 1) if trace[i, k] != trace[i, k - 1]]
            ref = ev["chi2_at_jacobian"]
            assert len(accepted) >= len(ref), (tag, i)
            for j in range(len(ref)):
                np.testing.assert_allclose(accepted[j], ref[j], rtol=(1e-13, 1e-9)[j] if j < 2 else tol_chi2, err_msg=f"{tag} instance {i} evaluation {j}")

"""Grid adaptation of the time-optimal grid, CPU side: the numpy restatement (oracle/grid_adaptation.py) against the
compiled reference's NonUniformFiniteDifferencesVariableGrid::adaptGridTimeBasedSingleStep run in isolation, the reference's adaptive MPC
loop as the GPU tests use it, and the argument checks of the C ABI that need no device."""
import numpy as np
import pytest

from control_box_rst_b200 import _abi as abi
from control_box_rst_b200 import problems, solver
from oracle import bindings, grid_adaptation as ga


@pytest.fixture(scope="module")
def reference():
    if not bindings.Reference.available():
        pytest.skip("compiled reference not present (oracle/_ref)")
    return bindings.Reference()


def _trajectory(rng, N, nx, nu, dt_profile):
    return rng.uniform(-1, 1, (N, nx)), rng.uniform(-1, 1, (N - 1, nu)), np.array(dt_profile, float)


@pytest.mark.parametrize("case", ["split_first", "split_middle", "merge_first", "merge_middle", "none", "split_blocked_by_n_max",
                                  "merge_blocked_by_n_min", "large_then_small", "small_then_large", "at_thresholds"])
def test_restatement_matches_the_compiled_reference_bit_for_bit(reference, case):
    N, dt_ref, hyst = 7, 0.1, 0.1
    n_min, n_max = 3, 12
    base = [0.1] * (N - 1)
    prof = {
        "split_first": [0.2] + base[1:],
        "split_middle": base[:2] + [0.1101] + base[3:],
        "merge_first": [0.05] + base[1:],
        "merge_middle": base[:3] + [0.0899] + base[4:],
        "none": [0.105, 0.095, 0.1, 0.109, 0.091, 0.1],
        "split_blocked_by_n_max": base[:2] + [0.3] + base[3:],
        "merge_blocked_by_n_min": base[:2] + [0.01] + base[3:],
        "large_then_small": [0.1, 0.15, 0.02, 0.1, 0.1, 0.1],
        "small_then_large": [0.1, 0.02, 0.15, 0.1, 0.1, 0.1],
        # dt_ref (1 +- hyst) themselves change nothing (strict comparisons), one ulp beyond does
        "at_thresholds": [0.1 * (1.0 + 0.1), 0.1 * (1.0 - 0.1), np.nextafter(0.1 * (1.0 - 0.1), 0.0), 0.1, 0.1, 0.1],
    }[case]
    if case == "split_blocked_by_n_max":
        n_max = N
    if case == "merge_blocked_by_n_min":
        n_min = N
    ocp = problems.unicycle_time_optimal(N, dt_ref)
    rng = np.random.default_rng(3)
    x, u, dt = _trajectory(rng, N, ocp.nx, ocp.nu, prof)
    xr, ur, dtr = reference.adapt_once(ocp, x, u, dt, n_min, n_max, hyst)
    xo, uo, dto, kind, i = ga.adapt_time_based_single_step(x, u, dt, n_min, n_max, dt_ref, hyst)
    assert xr.shape == xo.shape and np.array_equal(xr, xo)
    assert np.array_equal(ur, uo) and np.array_equal(dtr, dto)
    expect = {"split_first": (ga.SPLIT, 0), "split_middle": (ga.SPLIT, 2), "merge_first": (ga.MERGE, 0), "merge_middle": (ga.MERGE, 3),
              "none": (ga.NONE, -1), "split_blocked_by_n_max": (ga.NONE, -1), "merge_blocked_by_n_min": (ga.NONE, -1),
              "large_then_small": (ga.SPLIT, 1), "small_then_large": (ga.MERGE, 1), "at_thresholds": (ga.MERGE, 2)}[case]
    assert (kind, i) == expect


def test_split_keeps_the_interval_dt_and_halves_only_the_new_one(reference):
    """the quirk the device reproduces: after a split the total time GROWS by dt_i / 2"""
    ocp = problems.unicycle_time_optimal(5, 0.1)
    rng = np.random.default_rng(0)
    x, u, dt = _trajectory(rng, 5, 3, 2, [0.1, 0.4, 0.1, 0.1])
    xr, ur, dtr = reference.adapt_once(ocp, x, u, dt, 3, 10, 0.1)
    assert np.array_equal(dtr, [0.1, 0.4, 0.2, 0.1, 0.1])
    assert np.array_equal(xr[2], 0.5 * (x[1] + x[2])) and np.array_equal(ur[2], u[1])


def test_reference_adaptive_loop_obeys_its_own_rules(reference):
    ocp = problems.unicycle_time_optimal(16, 0.1)
    opts = abi.LmOptions.defaults(iterations=6, weights=(2.0, 2.0, 2.0))
    steps, m = 4, 3
    for goal, direction in (([0.3, 0.1, 0.2], -1), ([4.0, 1.0, 0.3], +1)):
        x0_seq = np.zeros((steps, 3))
        for warm in (True, False):
            n_trace, u0, x, u, dt = reference.adaptive_steps(ocp, opts, x0_seq, np.array(goal), 3, 26, 0.1, warm, m)
            flat = np.concatenate([[16], n_trace.ravel()])
            assert (np.abs(np.diff(flat)) <= 1).all()                      # one grid point per OCP iteration
            assert (n_trace[1:, 0] == n_trace[:-1, -1]).all() and n_trace[0, 0] == 16   # a new run does not adapt
            assert np.sign(n_trace[-1, -1] - 16) == direction
            assert x.shape == (n_trace[-1, -1], 3) and dt.shape == (n_trace[-1, -1] - 1,)
            assert np.array_equal(x[-1], goal)
            # a removed grid point 0 (continued run) promotes the old x_1 to start state until the next measurement: only a growing
            # warm-started grid is sure to still start at the measured state
            if direction > 0 or not warm:
                assert np.array_equal(x[0], x0_seq[-1])


def test_adaptive_front_end_checks_arguments_and_needs_a_device():
    ocp = problems.unicycle_time_optimal(16, 0.1)
    with pytest.raises(solver.B200SqpError) as e:
        solver.AdaptiveGridBatch(problems.van_der_pol(20), 4, 3, 30)
    assert e.value.code == abi.ERR_UNSUPPORTED
    for bad in (dict(n_min=2, n_max=30), dict(n_min=10, n_max=9), dict(n_min=3, n_max=30, dt_hyst_ratio=1.0)):
        with pytest.raises(solver.B200SqpError) as e:
            solver.AdaptiveGridBatch(ocp, 4, **bad)
        assert e.value.code == abi.ERR_INVALID
    if not solver.device_available():
        with pytest.raises(solver.B200SqpError) as e:
            solver.AdaptiveGridBatch(ocp, 4, 3, 30)
        assert e.value.code == abi.ERR_NO_DEVICE  # no CPU fallback


@pytest.mark.parametrize("seed", range(48))
def test_redundant_controls_restatement_matches_the_compiled_reference_bit_for_bit(reference, seed):
    """adaptGridRedundantControls on random trajectories with plateaus in the controls and a few vanishing dts: every combination of
    surplus (removals from the back), deficit (halving the largest interval, repeatedly) and the n_min / n_max stops"""
    rng = np.random.default_rng(100 + seed)
    N = int(rng.integers(4, 14))
    ocp = problems.unicycle_time_optimal(N, 0.1)
    x = rng.uniform(-1, 1, (N, 3))
    u = rng.uniform(-1, 1, (N - 1, 2))
    for k in range(1, N - 1):  # plateaus: a control repeats its predecessor (exactly, or within / just outside the threshold)
        r = rng.uniform()
        if r < 0.35:
            u[k] = u[k - 1]
        elif r < 0.5:
            u[k] = u[k - 1] + rng.choice([0.5e-3, 0.99e-3, 1.01e-3, 2e-3]) * rng.choice([-1, 1], 2)
    dt = rng.uniform(0.02, 0.3, N - 1)
    if seed % 3 == 0:
        dt[rng.integers(0, N - 1)] = 1e-7
    backup = int(rng.integers(0, 5))
    n_min = int(rng.integers(3, max(4, N - 1)))
    n_max = int(rng.integers(N, N + 4))
    xr, ur, dtr = reference.adapt_once(ocp, x, u, dt, n_min, n_max, redundant_controls=(backup, 1e-3))
    xo, uo, dto, ops = ga.adapt_redundant_controls(x, u, dt, n_min, n_max, 1e-3, backup)
    assert xr.shape == xo.shape, (xr.shape, xo.shape, ops)
    assert np.array_equal(xr, xo) and np.array_equal(ur, uo) and np.array_equal(dtr, dto)


def test_golden_fixture_of_the_adaptive_loop_is_what_the_compiled_reference_answers(reference):
    """tests/golden/grid_adaptation.npz (made by tests/golden/make_grid_adaptation.py) is what the GPU tests fall back to where the compiled
    reference did not travel: spot-check it against the live reference"""
    import os
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import test_gpu_grid_adaptation as T

    g = np.load(T.GOLDEN)
    for name, warm, redundant in (("vdp10", True, None), ("unicycle16", False, T.REDUNDANT)):
        ocp, opts, x0_seq, xf, n_min, n_max, hyst = T._inputs(name)
        stored = T._unpack(T._key(name, warm, redundant), g)
        for i in (0, 7, 23):
            live = bindings.isolated(lambda: reference.adaptive_steps(ocp, opts, x0_seq[:, i], xf[i], n_min, n_max, hyst, warm, T.M, redundant_controls=redundant))
            assert (live is None) == (stored[i] is None)
            if live is not None:
                assert np.array_equal(live[0][:, -1], stored[i][0][:, -1]) and np.array_equal(live[1], stored[i][1])
                assert np.array_equal(live[2], stored[i][2]) and np.array_equal(live[4], stored[i][4])


def _pack_params(ocp, x, u, dt, idx):
    """trajectory -> parameter vector of `ocp` (n_grid = len(x)); start and goal state fully fixed (not parameters)"""
    x_idx, u_idx, dt_idx = idx
    p = np.zeros(solver.dims_of(ocp).n_params)
    for k in range(len(x)):
        if x_idx[k] >= 0:
            p[x_idx[k]:x_idx[k] + ocp.nx] = x[k]
    for k in range(len(u)):
        p[u_idx[k]:u_idx[k] + ocp.nu] = u[k]
        p[dt_idx[k]] = dt[k]
    return p


def _unpack_params(ocp, p, x_start, x_goal, idx):
    x_idx, u_idx, dt_idx = idx
    N = ocp.n_grid
    x = np.array([p[x_idx[k]:x_idx[k] + ocp.nx] if x_idx[k] >= 0 else (x_start if k == 0 else x_goal) for k in range(N)])
    u = np.array([p[u_idx[k]:u_idx[k] + ocp.nu] for k in range(N - 1)])
    dt = np.array([p[dt_idx[k]] for k in range(N - 1)])
    return x, u, dt


@pytest.mark.parametrize("name", ["vdp10", "unicycle16"])
def test_cpu_restatement_of_the_whole_adaptive_loop_follows_the_reference(oracle, name):
    """The adaptive controller step restated end to end on the CPU -- LM solves by the oracle (oracle/sqp_oracle.cpp) on the grid of the
    moment, grid adaptation by oracle/grid_adaptation.py, the weights reset / adapted like LevenbergMarquardtSparse does -- reproduces the
    compiled reference's grid sizes, first controls and final trajectories (golden fixture) for the warm-started time-based strategy."""
    import os
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import test_gpu_grid_adaptation as T

    ocp0, opts, x0_seq, xf, n_min, n_max, hyst = T._inputs(name)
    stored = T._unpack(T._key(name, True, None), np.load(T.GOLDEN))
    checked = 0
    for i in range(0, T.B, 3):
        if stored[i] is None:
            continue
        x = u = dt = None
        n_after, u0s, undefined = [], [], False
        w = None
        for s in range(T.STEPS):
            for it in range(T.M):
                new_run = it == 0
                if x is not None and not new_run:
                    kind, at = ga.decide(dt, n_min, n_max, ocp0.dt_ref, hyst)
                    undefined = undefined or (kind != ga.NONE and at == len(dt) - 1)  # the reference indexes past its vectors there
                    x, u, dt, _, _ = ga.adapt_time_based_single_step(x, u, dt, n_min, n_max, ocp0.dt_ref, hyst)
                ocp = type(ocp0).from_buffer_copy(ocp0)
                if x is not None:
                    ocp.n_grid = len(x)
                idx = solver.vertex_indices(ocp)
                if x is None:  # first run: initializeSequences
                    x, u, dt = _unpack_params(ocp, oracle.initial_params(ocp, x0_seq[s, i], xf[i]), x0_seq[s, i], xf[i], idx)
                if new_run:
                    x[0] = x0_seq[s, i]
                    w = [opts.weight_eq, opts.weight_ineq, opts.weight_bounds]
                else:
                    w = [min(w[0] * opts.adapt_factor_eq, opts.adapt_max_eq), min(w[1] * opts.adapt_factor_ineq, opts.adapt_max_ineq),
                         min(w[2] * opts.adapt_factor_bounds, opts.adapt_max_bounds)]
                o = abi.LmOptions.defaults(iterations=opts.iterations, weights=tuple(w))
                p, _, _, _ = oracle.solve_batch(ocp, o, x[0][None, :], xf[i][None, :], _pack_params(ocp, x, u, dt, idx)[None, :])
                x, u, dt = _unpack_params(ocp, p[0], x[0], xf[i], idx)
            n_after.append(len(x))
            u0s.append(u[0].copy())
        if undefined:
            continue
        n_ref, u0_ref, x_ref, u_ref, dt_ref = stored[i]
        assert n_after == [int(v) for v in n_ref[:, -1]], (name, i, n_after, n_ref[:, -1])
        tol = 2e-4 if name.startswith("unicycle") else 1e-6
        assert np.abs(np.array(u0s) - u0_ref).max() <= tol and np.abs(x - x_ref).max() <= tol * max(1.0, np.abs(x_ref).max())
        assert np.abs(dt - dt_ref).max() <= tol
        checked += 1
    assert checked >= 4


@pytest.mark.parametrize("seed", range(40))
def test_time_based_restatement_on_random_trajectories(reference, seed):
    """adaptGridTimeBasedSingleStep on random dt profiles around the thresholds (values at, one ulp beside and far from
    dt_ref (1 +- hyst)), random grid sizes and n_min / n_max stops -- bit for bit against the compiled reference.  Profiles whose decision
    falls on the last interval are skipped: the reference indexes past the end of its vectors there."""
    rng = np.random.default_rng(500 + seed)
    N = int(rng.integers(3, 16))
    dt_ref, hyst = 0.1, float(rng.choice([0.05, 0.1, 0.25]))
    hi, lo = dt_ref * (1.0 + hyst), dt_ref * (1.0 - hyst)
    pool = [dt_ref, hi, lo, np.nextafter(hi, 1.0), np.nextafter(lo, 0.0), np.nextafter(hi, 0.0), np.nextafter(lo, 1.0), 0.5 * dt_ref, 2.0 * dt_ref]
    dt = np.array([rng.choice(pool) if rng.uniform() < 0.5 else dt_ref * rng.uniform(0.8, 1.2) for _ in range(N - 1)])
    n_min = int(rng.integers(2, N + 1))
    n_max = int(rng.integers(N, N + 3))
    kind, at = ga.decide(dt, n_min, n_max, dt_ref, hyst)
    if kind != ga.NONE and at == N - 2:
        pytest.skip("decision on the last interval: undefined in the reference")
    ocp = problems.unicycle_time_optimal(N, dt_ref)
    x, u = rng.uniform(-1, 1, (N, 3)), rng.uniform(-1, 1, (N - 1, 2))
    xr, ur, dtr = reference.adapt_once(ocp, x, u, dt, n_min, n_max, hyst)
    xo, uo, dto, _, _ = ga.adapt_time_based_single_step(x, u, dt, n_min, n_max, dt_ref, hyst)
    assert xr.shape == xo.shape and np.array_equal(xr, xo) and np.array_equal(ur, uo) and np.array_equal(dtr, dto)

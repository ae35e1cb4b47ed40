"""TEST INFRASTRUCTURE ONLY (see oracle/README in DESIGN.md section 5): numpy restatement of the reference's grid adaptation for the
non-uniform time-optimal grid, for one instance.  Only tests/ may import this; the product path is adapt_kernels.cu.

Follows NonUniformFiniteDifferencesVariableGrid::adaptGridTimeBasedSingleStep
(src/optimal_control/src/structured_ocp/discretization_grids/non_uniform_finite_differences_variable_grid.cpp:206-257):
walk the dt sequence once; the first dt above dt_ref (1 + hyst) while N < n_max gets a new grid point behind it (mid state, the interval's
control, half its dt -- the interval itself keeps its dt), the first dt below dt_ref (1 - hyst) while N > n_min loses its left grid point
(its dt is added to the successor); one change per call.  Pinned against the compiled reference in tests/test_grid_adaptation.py.

Where the reference indexes past the end of its vectors (the LAST interval is the one to change: `_x_seq[i + 1]` :225, `_dt_seq[i + 1]` :237)
it has no defined answer; this restatement follows the device's definition: x_f is the right neighbour of the last interval, the dt of a
removed last interval is dropped.
"""
import numpy as np

NONE, SPLIT, MERGE = 0, 1, 2


def decide(dt, n_min, n_max, dt_ref, hyst):
    """-> (kind, interval)"""
    n = len(dt) + 1
    hi, lo = dt_ref * (1.0 + hyst), dt_ref * (1.0 - hyst)
    for i, v in enumerate(dt):
        if v > hi and n < n_max:
            return SPLIT, i
        if v < lo and n > n_min:
            return MERGE, i
    return NONE, -1


def adapt_time_based_single_step(x, u, dt, n_min, n_max, dt_ref, hyst):
    """x [N][nx] (x[0] the start state, x[-1] = x_f), u [N-1][nu], dt [N-1] -> (x, u, dt, kind, interval) after one adaptation call"""
    x, u, dt = np.array(x, float), np.array(u, float), np.array(dt, float)
    kind, i = decide(dt, n_min, n_max, dt_ref, hyst)
    if kind == SPLIT:
        mid = 0.5 * (x[i] + x[i + 1])
        x = np.insert(x, i + 1, mid, axis=0)
        u = np.insert(u, i + 1, u[i], axis=0)
        dt = np.insert(dt, i + 1, 0.5 * dt[i])
    elif kind == MERGE:
        if i + 1 < len(dt):
            dt[i + 1] += dt[i]
        x = np.delete(x, i, axis=0)  # i == 0: the old x_1 becomes the (fixed) start state until the next measurement replaces it
        u = np.delete(u, i, axis=0)
        dt = np.delete(dt, i)
    return x, u, dt, kind, i


def adapt_redundant_controls(x, u, dt, n_min, n_max, epsilon, num_backup_nodes):
    """NonUniformFiniteDifferencesVariableGrid::adaptGridRedundantControls (non_uniform_finite_differences_variable_grid.cpp:259-352):
    an interval whose control repeats in its successor (all components within epsilon) or whose dt is below 1e-6 is redundant (the last
    interval never counts).  More redundant intervals than `num_backup_nodes`: the surplus is removed from the back (a removed interval's dt
    goes to its predecessor), down to n_min grid points.  Fewer: the missing ones are created by halving the interval with the largest dt
    (first maximum, last interval excluded; both halves get half the dt), up to n_max grid points.
    -> (x, u, dt, ops) with ops = [(+1, interval) for an insertion behind `interval` | (-1, k) for the removal of grid point k + 1]"""
    x, u, dt = [np.array(r, float) for r in x], [np.array(r, float) for r in u], [float(v) for v in dt]
    ops = []
    n = len(x)
    if n < 3:
        return np.array(x), np.array(u), np.array(dt), ops
    num_interv = len(u)
    non_unique = []
    for idx in range(num_interv - 1):
        if dt[idx] < 1e-6:
            non_unique.append(idx)
            continue
        if np.all(np.abs(u[idx + 1] - u[idx]) <= epsilon):
            non_unique.append(idx)
    diff = len(non_unique) - num_backup_nodes
    if diff < 0:
        for _ in range(-diff):
            if len(x) >= n_max:
                break
            i = 0
            if len(x) > 2:
                i = int(np.argmax(np.array(dt[:-1])))  # first maximum
            new_dt = 0.5 * dt[i]
            dt[i] = new_dt
            x.insert(i + 1, 0.5 * (x[i] + x[i + 1]))
            u.insert(i + 1, u[i].copy())
            dt.insert(i + 1, new_dt)
            ops.append((+1, i))
    elif diff > 0:
        it = len(non_unique) - 1
        for _ in range(diff):
            if len(x) <= n_min:
                break
            k = non_unique[it]
            if k >= len(x) - 2:
                k -= 1
            dt[k] += dt[k + 1]
            del x[k + 1], u[k + 1], dt[k + 1]
            ops.append((-1, k))
            it -= 1
    return np.array(x), np.array(u), np.array(dt), ops

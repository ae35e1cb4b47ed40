"""Python host side over the C ABI of libb200sqp.so (include/b200sqp.h).

`BatchedLevenbergMarquardt` mirrors the reference's solver interface for this path -- corbo::LevenbergMarquardtSparse
(src/optimization/include/corbo-optimization/solver/levenberg_marquardt_sparse.h:68-163) behind corbo::NlpSolverInterface
(solver/nlp_solver_interface.h:67-118) -- with the same method names, argument meaning and error behaviour, but for a whole batch
of independent instances of one OCP structure.  All computation happens in the CUDA library; there is no Python/numpy/CPU fallback:
if the library or a B200-class device is missing, construction raises.
"""
import ctypes as C
import os

import numpy as np

from . import _abi as abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200SQP_LIB") or os.path.join(_HERE, "libb200sqp.so")  # B200SQP_LIB: development builds only

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)

# every symbol include/b200sqp.h declares (tests check the library exports all of them)
ABI_SYMBOLS = [
    "b200sqp_create", "b200sqp_destroy", "b200sqp_last_error", "b200sqp_device_available", "b200sqp_dims_of", "b200sqp_get_dims",
    "b200sqp_vertex_indices", "b200sqp_edge_indices", "b200sqp_jacobian_pattern", "b200sqp_set_problem_data",
    "b200sqp_initialize_trajectories", "b200sqp_set_params", "b200sqp_get_params", "b200sqp_get_first_controls", "b200sqp_solve",
    "b200sqp_solve_async", "b200sqp_synchronize", "b200sqp_step", "b200sqp_evaluate", "b200sqp_get_statistics", "b200sqp_get_chi2_trace",
    "b200sqp_last_solve_ms", "b200sqp_launch_count", "b200sqp_device_pointers", "b200sqp_set_stream", "b200sqp_set_threads_per_instance",
    "b200sqp_set_phase_profile", "b200sqp_get_phase_cycles", "b200sqp_final_constraint_indices",
    "b200sqp_peer_export", "b200sqp_peer_attach", "b200sqp_peer_wait", "b200sqp_peer_gathered", "b200sqp_peer_detach", "b200sqp_peer_status", "b200sqp_linearize_dynamics", "b200sqp_warm_start_shift", "b200sqp_mpc_step",
    "b200sqp_plant_step", "b200sqp_closed_loop", "b200sqp_dynamics_hessian", "b200sqp_set_feature_set", "b200sqp_measure_fp64_peak", "b200sqp_set_precision", "b200sqp_set_reference_trajectory", "b200sqp_dt_equality_indices",
    "b200sqp_adaptive_create", "b200sqp_adaptive_destroy", "b200sqp_adaptive_step", "b200sqp_adaptive_get_trajectories",
    "b200sqp_adaptive_statistics", "b200sqp_adaptive_reserve", "b200sqp_adaptive_last_interval_changes", "b200sqp_adaptive_set_redundant_controls",
]


class B200SqpError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"b200sqp error {code}: {message}")
        self.code = code


_lib = None


def load_library():
    """dlopen libb200sqp.so (built in-tree by __graft_entry__.build / csrc/Makefile).  Fails loudly if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc); "
                                    "there is no CPU fallback for the LM hot path")
        lib = C.CDLL(LIB_PATH)
        lib.b200sqp_last_error.restype = C.c_char_p
        for name in ABI_SYMBOLS:
            if name != "b200sqp_last_error":
                getattr(lib, name).restype = C.c_int
        _lib = lib
    return _lib


def _check(rc):
    if rc != 0:
        raise B200SqpError(rc, load_library().b200sqp_last_error().decode())


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def device_available():
    return bool(load_library().b200sqp_device_available())


# ---- handle-less structure queries (run without a GPU) -----------------------------------------------------------------------
def dims_of(ocp):
    out = abi.Dims()
    _check(load_library().b200sqp_dims_of(C.byref(ocp), C.byref(out)))
    return out


def vertex_indices(ocp):
    N = ocp.n_grid
    x_idx, u_idx, dt_idx = np.full(N, -2, np.int32), np.full(N - 1, -2, np.int32), np.full(N - 1, -2, np.int32)
    _check(load_library().b200sqp_vertex_indices(C.byref(ocp), _i(x_idx), _i(u_idx), _i(dt_idx)))
    return x_idx, u_idx, dt_idx


def edge_indices(ocp):
    K = ocp.n_grid - 1
    sc, cc, tc, dy = np.full(K, -2, np.int32), np.full(K, -2, np.int32), np.full(2 * K, -2, np.int32), np.full(K, -2, np.int32)
    fc = C.c_int32(-2)
    _check(load_library().b200sqp_edge_indices(C.byref(ocp), _i(sc), _i(cc), _i(tc), _i(dy), C.byref(fc)))
    return dict(state_cost=sc, control_cost=cc, dt_cost=tc.reshape(K, 2), dynamics=dy, final_cost=fc.value)


def dt_equality_indices(ocp):
    """row offset (equality category) of the TwoScalarEqualEdge between dt_{k-1} and dt_k, [N-1], -1 where there is none"""
    out = np.full(ocp.n_grid - 1, -2, np.int32)
    _check(load_library().b200sqp_dt_equality_indices(C.byref(ocp), _i(out)))
    return out


def final_constraint_indices(ocp):
    eq, ineq = C.c_int32(-2), C.c_int32(-2)
    _check(load_library().b200sqp_final_constraint_indices(C.byref(ocp), C.byref(eq), C.byref(ineq)))
    return eq.value, ineq.value


def jacobian_pattern(ocp):
    d = dims_of(ocp)
    col_ptr, row_idx = np.zeros(d.n_params + 1, np.int32), np.zeros(d.nnz_jacobian, np.int32)
    _check(load_library().b200sqp_jacobian_pattern(C.byref(ocp), _i(col_ptr), _i(row_idx)))
    return col_ptr, row_idx


def linearize_dynamics(dynamics, dyn_params, x, u, method="forward", device=0):
    """SystemDynamicsInterface::getLinearA / getLinearB for a batch of points on the device.
    x [B, nx], u [B, nu] -> A [B, nx, nx], B [B, nx, nu]; method 'forward' (the reference's default) or 'central'."""
    nx, nu = abi.DYN_DIMS[dynamics]
    x = np.ascontiguousarray(x, np.float64).reshape(-1, nx)
    u = np.ascontiguousarray(u, np.float64).reshape(-1, nu)
    B = x.shape[0]
    assert u.shape[0] == B
    p = np.zeros(abi.MAX_DYN_PARAMS)
    p[:len(dyn_params)] = dyn_params
    A = np.zeros((B, nx, nx))
    Bm = np.zeros((B, nu, nx))
    _check(load_library().b200sqp_linearize_dynamics(C.c_int32(dynamics), _d(p), C.c_int32({"forward": 0, "central": 1}[method]), C.c_int32(B),
                                                      _d(x), _d(u), _d(A), _d(Bm), C.c_int32(device)))
    # the library writes column-major blocks per point
    return A.transpose(0, 2, 1).copy(), Bm.transpose(0, 2, 1).copy()


def dynamics_hessian(dynamics, dyn_params, x, u, multipliers=None, method="forward", device=0):
    """ForwardDifferences / CentralDifferences::hessian of the dynamics w.r.t. [x; u] for a batch of points on the device.
    x [B, nx], u [B, nu], multipliers [B, nx] or None -> H [B, nx+nu, nx+nu] (H[b, i, j] = sum_v mult_v d2 f_v / dz_i dz_j)"""
    nx, nu = abi.DYN_DIMS[dynamics]
    x = np.ascontiguousarray(x, np.float64).reshape(-1, nx)
    u = np.ascontiguousarray(u, np.float64).reshape(-1, nu)
    B = x.shape[0]
    m = None if multipliers is None else np.ascontiguousarray(multipliers, np.float64).reshape(B, nx)
    p = np.zeros(abi.MAX_DYN_PARAMS)
    p[:len(dyn_params)] = dyn_params
    H = np.zeros((B, nx + nu, nx + nu))
    _check(load_library().b200sqp_dynamics_hessian(C.c_int32(dynamics), _d(p), C.c_int32({"forward": 0, "central": 1}[method]), C.c_int32(B),
                                                    _d(x), _d(u), _d(m), _d(H), C.c_int32(device)))
    return H.transpose(0, 2, 1).copy()  # the library writes column-major blocks per point


def measure_fp64_peak(device=0):
    """fp64 FMA throughput of the device in TFLOP/s, measured live (the issue bound of the fused LM kernel)"""
    v = C.c_double(0)
    _check(load_library().b200sqp_measure_fp64_peak(C.c_int32(device), C.byref(v)))
    return v.value


INTEGRATORS = {"euler": 0, "rk4": 1}


def plant_step(dynamics, dyn_params, x, u, dt, integrator="euler", device=0):
    """SimulatedPlant::control for a batch of plants on the device: x [B, nx], u [B, nu] -> x_next [B, nx] after one solveIVP over dt
    ('euler' = IntegratorExplicitEuler, the plant's default; 'rk4' = IntegratorExplicitRungeKutta4)."""
    nx, nu = abi.DYN_DIMS[dynamics]
    x = np.ascontiguousarray(x, np.float64).reshape(-1, nx)
    u = np.ascontiguousarray(u, np.float64).reshape(-1, nu)
    B = x.shape[0]
    assert u.shape[0] == B
    p = np.zeros(abi.MAX_DYN_PARAMS)
    p[:len(dyn_params)] = dyn_params
    xn = np.zeros((B, nx))
    _check(load_library().b200sqp_plant_step(C.c_int32(dynamics), _d(p), C.c_int32(INTEGRATORS[integrator]), C.c_double(dt), C.c_int32(B),
                                              _d(x), _d(u), _d(xn), C.c_int32(device)))
    return xn


class BatchedLevenbergMarquardt:
    """LevenbergMarquardtSparse for `batch` instances of one OCP structure, resident on one B200.

    Method names follow the reference class: setIterations / setPenaltyWeights / setWeightAdapation / solve / clear.
    """

    def __init__(self, ocp, batch, device=0):
        self._lib = load_library()
        self.ocp = ocp
        self.batch = int(batch)
        self.device = int(device)
        self._h = C.c_void_p()
        # LevenbergMarquardtSparse defaults (levenberg_marquardt_sparse.h:112-124)
        self._opts = abi.LmOptions.defaults()
        _check(self._lib.b200sqp_create(C.byref(ocp), C.c_int32(self.batch), C.c_int32(self.device), C.byref(self._h)))
        self.dims = abi.Dims()
        _check(self._lib.b200sqp_get_dims(self._h, C.byref(self.dims)))

    # -- NlpSolverInterface ---------------------------------------------------------------------------------------------------
    def isLsqSolver(self):
        return True

    def setIterations(self, iterations):
        self._opts.iterations = int(iterations)

    def setPenaltyWeights(self, weight_eq, weight_ineq, weight_bounds):
        self._opts.weight_eq, self._opts.weight_ineq, self._opts.weight_bounds = weight_eq, weight_ineq, weight_bounds

    def setWeightAdapation(self, factor_eq, factor_ineq, factor_bounds, max_eq, max_ineq, max_bounds):
        o = self._opts
        o.adapt_factor_eq, o.adapt_factor_ineq, o.adapt_factor_bounds = factor_eq, factor_ineq, factor_bounds
        o.adapt_max_eq, o.adapt_max_ineq, o.adapt_max_bounds = max_eq, max_ineq, max_bounds

    def solve(self, new_run=True, fetch=True):
        """LevenbergMarquardtSparse::solve for every instance.  -> (status [B] int32, chi2 [B]) or None when fetch=False
        (results stay in HBM; use synchronize()/status()/get_params())."""
        if not fetch:
            _check(self._lib.b200sqp_solve_async(self._h, C.byref(self._opts), C.c_int32(1 if new_run else 0)))
            return None
        status = np.zeros(self.batch, np.int32)
        chi2 = np.zeros(self.batch)
        _check(self._lib.b200sqp_solve(self._h, C.byref(self._opts), C.c_int32(1 if new_run else 0), _i(status), _d(chi2)))
        return status, chi2

    def clear(self):
        if self._h:
            self._lib.b200sqp_destroy(self._h)
            self._h = C.c_void_p()

    close = clear

    def __del__(self):
        try:
            self.clear()
        except Exception:
            pass

    # -- problem data -------------------------------------------------------------------------------------------------------
    def set_problem_data(self, x0, xref=None):
        x0 = np.ascontiguousarray(x0, np.float64)
        assert x0.shape == (self.batch, self.ocp.nx), x0.shape
        if xref is not None:
            xref = np.ascontiguousarray(xref, np.float64)
            assert xref.shape == x0.shape
        _check(self._lib.b200sqp_set_problem_data(self._h, _d(x0), _d(xref)))

    def set_reference_trajectory(self, xref_traj):
        """time-varying state reference [batch, n_grid, nx] (row k = getReferenceCached(k)); None = back to the static reference.  Call after
        set_problem_data."""
        if xref_traj is None:
            _check(self._lib.b200sqp_set_reference_trajectory(self._h, None))
            return
        xref_traj = np.ascontiguousarray(xref_traj, np.float64)
        assert xref_traj.shape == (self.batch, self.ocp.n_grid, self.ocp.nx), xref_traj.shape
        _check(self._lib.b200sqp_set_reference_trajectory(self._h, _d(xref_traj)))

    def initialize_trajectories(self):
        _check(self._lib.b200sqp_initialize_trajectories(self._h))

    def warm_start_shift(self, x0_new):
        """moving-horizon warm start on the device (FullDiscretizationGridBase::warmStartShifting per instance) -> num_shift [B]"""
        x0_new = np.ascontiguousarray(x0_new, np.float64)
        assert x0_new.shape == (self.batch, self.ocp.nx)
        shifts = np.zeros(self.batch, np.int32)
        _check(self._lib.b200sqp_warm_start_shift(self._h, _d(x0_new), _i(shifts)))
        return shifts

    def set_params(self, params):
        params = np.ascontiguousarray(params, np.float64)
        assert params.shape == (self.batch, self.dims.n_params), params.shape
        _check(self._lib.b200sqp_set_params(self._h, _d(params)))

    def get_params(self):
        out = np.zeros((self.batch, self.dims.n_params))
        _check(self._lib.b200sqp_get_params(self._h, _d(out)))
        return out

    def get_first_controls(self):
        out = np.zeros((self.batch, self.ocp.nu))
        _check(self._lib.b200sqp_get_first_controls(self._h, _d(out)))
        return out

    def step(self, x0, xref=None, cold_start=True, out=None):
        """One batched MPC step through host buffers: H2D x0 (+xref), [initialise], solve, D2H parameters, chi2, status."""
        x0 = np.ascontiguousarray(x0, np.float64)
        xref = None if xref is None else np.ascontiguousarray(xref, np.float64)
        if out is None:
            out = (np.zeros((self.batch, self.dims.n_params)), np.zeros(self.batch), np.zeros(self.batch, np.int32))
        params, chi2, status = out
        _check(self._lib.b200sqp_step(self._h, C.byref(self._opts), C.c_int32(1 if cold_start else 0), _d(x0), _d(xref), _d(params), _d(chi2),
                                      _i(status)))
        return params, chi2, status

    MPC_COLD, MPC_KEEP, MPC_SHIFT = 0, 1, 2

    def mpc_step(self, x0, xref=None, mode=1, out=None):
        """One closed-loop MPC step of the batch: measured states in, first controls out (trajectories stay in HBM).
        mode: MPC_COLD initialise, MPC_KEEP previous solution as guess, MPC_SHIFT moving-horizon warm start.  -> u0 [B, nu], chi2, status"""
        x0 = np.ascontiguousarray(x0, np.float64)
        xref = None if xref is None else np.ascontiguousarray(xref, np.float64)
        if out is None:
            out = (np.zeros((self.batch, self.ocp.nu)), np.zeros(self.batch), np.zeros(self.batch, np.int32))
        u0, chi2, status = out
        _check(self._lib.b200sqp_mpc_step(self._h, C.byref(self._opts), C.c_int32(mode), _d(x0), _d(xref), _d(u0), _d(chi2), _i(status)))
        return u0, chi2, status

    def closed_loop(self, x0, steps, xref=None, mode=1, integrator="euler", plant_dt=None):
        """The whole closed loop of the batch on the device: `steps` x (MPC step -> first control -> plant step over plant_dt), only
        x0 in and the log out.  -> u_applied [steps, B, nu], x_closed [steps+1, B, nx], chi2 [steps, B], status [steps, B]"""
        x0 = np.ascontiguousarray(x0, np.float64)
        xref = None if xref is None else np.ascontiguousarray(xref, np.float64)
        B, nx, nu = self.batch, self.ocp.nx, self.ocp.nu
        u = np.zeros((steps, B, nu))
        x = np.zeros((steps + 1, B, nx))
        chi2 = np.zeros((steps, B))
        status = np.zeros((steps, B), np.int32)
        dt = float(self.ocp.dt_ref if plant_dt is None else plant_dt)
        _check(self._lib.b200sqp_closed_loop(self._h, C.byref(self._opts), C.c_int32(mode), C.c_int32(INTEGRATORS[integrator]), C.c_double(dt),
                                             C.c_int32(steps), _d(x0), _d(xref), _d(u), _d(x), _d(chi2), _i(status)))
        return u, x, chi2, status

    def closed_loop_raw(self, mode, integrator, plant_dt, steps, x0_ptr, xref_ptr, u_ptr, x_ptr, chi2_ptr, status_ptr):
        """b200sqp_closed_loop on raw host addresses (pinned buffers; 0 = not wanted)"""
        _check(self._lib.b200sqp_closed_loop(self._h, C.byref(self._opts), C.c_int32(mode), C.c_int32(integrator), C.c_double(plant_dt),
                                             C.c_int32(steps), C.c_void_p(x0_ptr), C.c_void_p(xref_ptr), C.c_void_p(u_ptr or None),
                                             C.c_void_p(x_ptr or None), C.c_void_p(chi2_ptr or None), C.c_void_p(status_ptr or None)))

    def mpc_step_raw(self, mode, x0_ptr, xref_ptr, u0_ptr, chi2_ptr, status_ptr):
        """b200sqp_mpc_step on raw host addresses (pinned buffers)"""
        _check(self._lib.b200sqp_mpc_step(self._h, C.byref(self._opts), C.c_int32(mode), C.c_void_p(x0_ptr), C.c_void_p(xref_ptr),
                                          C.c_void_p(u0_ptr), C.c_void_p(chi2_ptr), C.c_void_p(status_ptr)))

    def step_raw(self, x0_ptr, xref_ptr, params_ptr, chi2_ptr, status_ptr, cold_start=True):
        """b200sqp_step on raw host addresses (e.g. pinned torch tensors' data_ptr())."""
        _check(self._lib.b200sqp_step(self._h, C.byref(self._opts), C.c_int32(1 if cold_start else 0), C.c_void_p(x0_ptr), C.c_void_p(xref_ptr),
                                      C.c_void_p(params_ptr), C.c_void_p(chi2_ptr), C.c_void_p(status_ptr)))

    # -- evaluation surface (computeValues + computeCombinedSparseJacobian) ------------------------------------------------------------
    def evaluate(self, weights=(2.0, 2.0, 2.0), jacobian=True):
        """-> values [B, m], J values [B, nnzJ] in the CSC order of jacobian_pattern(ocp).  Perturbs the parameters like the reference."""
        m, nnz = self.dims.m, self.dims.nnz_jacobian
        values = np.zeros((self.batch, m))
        jac = np.zeros((self.batch, nnz)) if jacobian else None
        _check(self._lib.b200sqp_evaluate(self._h, C.c_double(weights[0]), C.c_double(weights[1]), C.c_double(weights[2]), _d(values), _d(jac)))
        return values, jac

    # -- bookkeeping ----------------------------------------------------------------------------------------------------------
    def synchronize(self):
        _check(self._lib.b200sqp_synchronize(self._h))

    def statistics(self):
        B = self.batch
        f, r, l = np.zeros(B, np.int32), np.zeros(B, np.int32), np.zeros(B, np.int32)
        mu, rho = np.zeros(B), np.zeros(B)
        _check(self._lib.b200sqp_get_statistics(self._h, _i(f), _i(r), _i(l), _d(mu), _d(rho)))
        return dict(inner_passes=f, rejects=r, relinearizations=l, mu=mu, rho=rho)

    def chi2_trace(self):
        it = self._opts.iterations
        out = np.zeros((self.batch, it + 1))
        _check(self._lib.b200sqp_get_chi2_trace(self._h, _d(out), C.c_int32(it)))
        return out

    def last_solve_ms(self):
        ms = C.c_float(0)
        _check(self._lib.b200sqp_last_solve_ms(self._h, C.byref(ms)))
        return ms.value

    def launch_count(self):
        n = C.c_int64(0)
        _check(self._lib.b200sqp_launch_count(self._h, C.byref(n)))
        return n.value

    def device_pointers(self):
        chi2, status, x0 = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _check(self._lib.b200sqp_device_pointers(self._h, C.byref(chi2), C.byref(status), C.byref(x0)))
        return dict(chi2=chi2.value, status=status.value, x0=x0.value)

    def set_threads_per_instance(self, threads):
        """cooperating threads per instance (1, 2, 4, 8; 0 = pick from the batch size)"""
        _check(self._lib.b200sqp_set_threads_per_instance(self._h, C.c_int32(threads)))

    def set_precision(self, precision):
        """'f64' (default, the reference's arithmetic) or 'f32' (reduced-precision variant of BASELINE configs[4]; structures that run the
        warp-cooperative pipeline only, raises B200SqpError(UNSUPPORTED) otherwise)"""
        _check(self._lib.b200sqp_set_precision(self._h, C.c_int32({"f64": 0, "f32": 1}[precision])))

    def set_feature_set(self, general):
        """test knob: force the general compile-time feature set of the LM kernel on a lean-eligible structure (False = automatic)"""
        _check(self._lib.b200sqp_set_feature_set(self._h, C.c_int32(1 if general else 0)))

    def set_phase_profile(self, enable=True):
        """measurement aid: per-phase clock64() accounting inside the LM kernel (see phase_cycles)"""
        _check(self._lib.b200sqp_set_phase_profile(self._h, C.c_int32(1 if enable else 0)))

    def phase_cycles(self):
        """mean SM cycles per thread block of the last solve: dict(linearize, factor_solve, trial, control, exchange)"""
        out = np.zeros(5)
        _check(self._lib.b200sqp_get_phase_cycles(self._h, _d(out)))
        return dict(zip(("linearize", "factor_solve", "trial", "control", "exchange"), out.tolist()))

    # -- fused stop-test gather over NVLink peer memory (one process per GPU) ---------------------------------------------------
    def peer_export(self, world, rank):
        """-> 64-byte CUDA IPC handle of this rank's gather buffer (bytes)"""
        buf = C.create_string_buffer(64)
        _check(self._lib.b200sqp_peer_export(self._h, C.c_int32(world), C.c_int32(rank), buf))
        return buf.raw

    def peer_attach(self, handles):
        """handles: list of the 64-byte IPC handles of all ranks, in rank order"""
        blob = b"".join(handles)
        _check(self._lib.b200sqp_peer_attach(self._h, C.c_char_p(blob)))

    def peer_wait(self):
        _check(self._lib.b200sqp_peer_wait(self._h))

    def peer_gathered_ptr(self):
        p = C.c_void_p()
        _check(self._lib.b200sqp_peer_gathered(self._h, C.byref(p)))
        return p.value

    def peer_timed_out(self):
        t = C.c_int32(0)
        _check(self._lib.b200sqp_peer_status(self._h, C.byref(t)))
        return bool(t.value)

    def peer_detach(self):
        _check(self._lib.b200sqp_peer_detach(self._h))

    def set_stream(self, cuda_stream):
        _check(self._lib.b200sqp_set_stream(self._h, C.c_void_p(cuda_stream)))

    @property
    def options(self):
        return self._opts


class AdaptiveGridBatch:
    """`batch` time-optimal controllers whose grids adapt independently: the reference's NonUniformFiniteDifferencesVariableGrid with
    setGridAdaptTimeBasedSingleStep(n_max, dt_hyst_ratio), setNmin(n_min), setWarmStart(warm_start)
    (non_uniform_finite_differences_variable_grid.h:52-54) under PredictiveController::step's OCP iterations.  Instances are bucketed by
    grid size on the device (include/b200sqp.h, b200sqp_adaptive_*).  Setter names follow the reference's solver class."""

    def __init__(self, ocp, batch, n_min, n_max, dt_hyst_ratio=0.1, warm_start=True, device=0):
        self._lib = load_library()
        self.ocp = ocp
        self.batch = int(batch)
        self.n_min, self.n_max = int(n_min), int(n_max)
        self._opts = abi.LmOptions.defaults()
        self._h = C.c_void_p()
        _check(self._lib.b200sqp_adaptive_create(C.byref(ocp), C.c_int32(self.batch), C.c_int32(device), C.c_int32(n_min), C.c_int32(n_max),
                                                 C.c_double(dt_hyst_ratio), C.c_int32(1 if warm_start else 0), C.byref(self._h)))

    setIterations = BatchedLevenbergMarquardt.setIterations
    setPenaltyWeights = BatchedLevenbergMarquardt.setPenaltyWeights
    setWeightAdapation = BatchedLevenbergMarquardt.setWeightAdapation

    def step(self, x0, xref, num_ocp_iterations=1):
        """one controller step of every instance -> (u0 [B, nu], chi2 [B], status [B], n [B])"""
        x0 = np.ascontiguousarray(x0, np.float64)
        xref = np.ascontiguousarray(xref, np.float64)
        assert x0.shape == (self.batch, self.ocp.nx) and xref.shape == x0.shape
        u0 = np.zeros((self.batch, self.ocp.nu))
        chi2 = np.zeros(self.batch)
        status = np.zeros(self.batch, np.int32)
        n = np.zeros(self.batch, np.int32)
        _check(self._lib.b200sqp_adaptive_step(self._h, C.byref(self._opts), C.c_int32(num_ocp_iterations), _d(x0), _d(xref), _d(u0), _d(chi2),
                                               _i(status), _i(n)))
        return u0, chi2, status, n

    def setGridAdaptRedundantControls(self, num_backup_nodes=1, epsilon=1e-3):
        """the reference's second strategy (NonUniformFiniteDifferencesVariableGrid::setGridAdaptRedundantControls); before the first step"""
        _check(self._lib.b200sqp_adaptive_set_redundant_controls(self._h, C.c_int32(num_backup_nodes), C.c_double(epsilon)))

    def reserve(self, n_from=None, n_to=None):
        """create the buckets of these grid sizes now (default: the whole reachable range) instead of on first use"""
        _check(self._lib.b200sqp_adaptive_reserve(self._h, C.c_int32(self.n_min if n_from is None else n_from),
                                                  C.c_int32(self.n_max if n_to is None else n_to)))

    def last_interval_changes(self):
        """[B]: adaptations per instance that changed its last interval -- inputs the reference has no defined answer for"""
        out = np.zeros(self.batch, np.int32)
        _check(self._lib.b200sqp_adaptive_last_interval_changes(self._h, _i(out)))
        return out

    def trajectories(self):
        """-> (x [B, n_cap, nx], u [B, n_cap, nu], dt [B, n_cap], n [B]); rows beyond an instance's grid are zero"""
        cap = max(self.n_max, self.ocp.n_grid)
        x = np.zeros((self.batch, cap, self.ocp.nx))
        u = np.zeros((self.batch, cap, self.ocp.nu))
        dt = np.zeros((self.batch, cap))
        n = np.zeros(self.batch, np.int32)
        _check(self._lib.b200sqp_adaptive_get_trajectories(self._h, C.c_int32(cap), _d(x), _d(u), _d(dt), _i(n)))
        return x, u, dt, n

    def statistics(self):
        b = C.c_int32(0)
        sp, me, la = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        _check(self._lib.b200sqp_adaptive_statistics(self._h, C.byref(b), C.byref(sp), C.byref(me), C.byref(la)))
        return dict(occupied_buckets=b.value, splits=sp.value, merges=me.value, launches=la.value)

    def clear(self):
        if self._h:
            self._lib.b200sqp_adaptive_destroy(self._h)
            self._h = C.c_void_p()

    close = clear

    def __del__(self):
        try:
            self.clear()
        except Exception:
            pass

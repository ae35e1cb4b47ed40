"""TEST INFRASTRUCTURE ONLY -- ctypes bindings of the two CPU checkers.

  * `Oracle`     oracle/libsqp_oracle.so   plain-C++ restatement of the reference algorithm (oracle/sqp_oracle.cpp)
  * `Reference`  oracle/_ref/libcorbo_ref.so  the unmodified reference compiled from /root/reference (oracle/ref_driver.cpp)

Both expose the same calls with the same argument meaning, prefixed `sqp_oracle_` / `corbo_ref_`.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module; the product
(control_box_rst_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from control_box_rst_b200 import _abi as abi

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libsqp_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libcorbo_ref.so")

EV_JACOBIAN, EV_INCREMENT, EV_RESTORE, EV_DISCARD = 0, 1, 2, 3

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def build_oracle():
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)


def build_reference(reference_root="/root/reference"):
    """Compile the unmodified reference where it lies (only possible where /root/reference exists)."""
    if not os.path.isdir(reference_root):
        return False
    subprocess.run(["make", "-s", "-j8", "-C", HERE, "ref", "REF=" + reference_root], check=True)
    return True


def isolated(fn, timeout=120.0):
    """fn() in a forked child -> its result, or None if the child died.  Needed around Reference.adaptive_steps: the reference indexes one
    past the end of its vertex vectors when the interval it splits or merges is the last one
    (non_uniform_finite_differences_variable_grid.cpp:225,237: _x_seq[i + 1], _dt_seq[i + 1] with i = size - 1), which ends in heap
    corruption; such instances have no reference answer.  The child only runs CPU code and leaves through os._exit."""
    import pickle

    r, w = os.pipe()
    pid = os.fork()
    if pid == 0:
        code = 1
        try:
            import faulthandler

            faulthandler.disable()  # a dying child is an expected outcome here, not a crash report
            os.dup2(os.open(os.devnull, os.O_WRONLY), 2)
            os.close(r)
            with os.fdopen(w, "wb") as f:
                pickle.dump(fn(), f)
            code = 0
        finally:
            os._exit(code)
    os.close(w)
    import select
    import signal
    import time

    chunks, deadline = [], time.monotonic() + timeout
    while True:
        left = deadline - time.monotonic()
        ready = select.select([r], [], [], max(left, 0.0))[0] if left > 0 else []
        if not ready:  # a child that neither finishes nor dies (a fork of a multi-threaded parent may inherit a held lock)
            os.kill(pid, signal.SIGKILL)
            break
        chunk = os.read(r, 1 << 16)
        if not chunk:
            break
        chunks.append(chunk)
    os.close(r)
    _, status = os.waitpid(pid, 0)
    data = b"".join(chunks)
    if status != 0 or not data:
        return None
    return pickle.loads(data)


class _Checker:
    prefix = ""
    path = ""

    def __init__(self):
        if not os.path.exists(self.path):
            raise FileNotFoundError(self.path)
        self.lib = C.CDLL(self.path)
        p = self.prefix
        self._dims = getattr(self.lib, p + "dims")
        self._vertex_indices = getattr(self.lib, p + "vertex_indices")
        self._edge_table = getattr(self.lib, p + "edge_table")
        self._initial_params = getattr(self.lib, p + "initial_params")
        self._evaluate = getattr(self.lib, p + "evaluate")
        self._trace = getattr(self.lib, p + "trace")
        self._solve_batch = getattr(self.lib, p + "solve_batch")
        for f in (self._dims, self._vertex_indices, self._edge_table, self._initial_params, self._evaluate, self._trace, self._solve_batch):
            f.restype = C.c_int

    @classmethod
    def available(cls):
        return os.path.exists(cls.path)

    def set_xref_points(self, n_points):
        """n_points > 1: every xref argument is a time-varying reference [n_grid, nx] per instance (row k = what the reference's
        getReferenceCached(k) returns); 0: static reference [nx] (the default).  Process-global switch of the checker library."""
        f = getattr(self.lib, self.prefix + "set_xref_points")
        f.restype = C.c_int
        f(C.c_int(int(n_points)))

    def dims(self, ocp):
        out = abi.Dims()
        rc = self._dims(C.byref(ocp), C.byref(out))
        if rc != 0:
            raise RuntimeError(f"{self.prefix}dims failed: {rc}")
        return out

    def vertex_indices(self, ocp):
        N = ocp.n_grid
        x_idx = np.full(N, -2, np.int32)
        u_idx = np.full(N - 1, -2, np.int32)
        dt_idx = np.full(N - 1, -2, np.int32)
        rc = self._vertex_indices(C.byref(ocp), _i(x_idx), _i(u_idx), _i(dt_idx))
        if rc != 0:
            raise RuntimeError(f"{self.prefix}vertex_indices failed: {rc}")
        return x_idx, u_idx, dt_idx

    def edge_table(self, ocp, category, max_edges=4096):
        """rows: [dim, edge_idx, n_vertices, vertex_idx0..3] in creation order for category 0 lsq / 1 eq / 2 ineq"""
        table = np.zeros((max_edges, 7), np.int32)
        cnt = self._edge_table(C.byref(ocp), C.c_int(category), _i(table), C.c_int(max_edges))
        if cnt < 0:
            raise RuntimeError(f"{self.prefix}edge_table failed: {cnt}")
        return table[:cnt].copy()

    def linearize(self, ocp, x, u, method="forward"):
        """getLinearA / getLinearB of the descriptor's dynamics at one point -> A [nx, nx], B [nx, nu]"""
        nx, nu = ocp.nx, ocp.nu
        x = np.ascontiguousarray(x, np.float64)
        u = np.ascontiguousarray(u, np.float64)
        A, Bm = np.zeros((nx, nx)), np.zeros((nu, nx))
        fn = getattr(self.lib, self.prefix + "linearize")
        fn.restype = C.c_int
        rc = fn(C.byref(ocp), C.c_int({"forward": 0, "central": 1}[method]), _d(x), _d(u), _d(A), _d(Bm))
        if rc != 0:
            raise RuntimeError(f"{self.prefix}linearize failed: {rc}")
        return A.T.copy(), Bm.T.copy()  # the checkers write column-major

    def dynamics_hessian(self, ocp, x, u, multipliers=None, method="forward"):
        """ForwardDifferences / CentralDifferences::hessian of the descriptor's dynamics w.r.t. [x; u] at one point -> H [nz, nz]"""
        nz = ocp.nx + ocp.nu
        x = np.ascontiguousarray(x, np.float64)
        u = np.ascontiguousarray(u, np.float64)
        m = None if multipliers is None else np.ascontiguousarray(multipliers, np.float64)
        H = np.zeros((nz, nz))
        fn = getattr(self.lib, self.prefix + "dynamics_hessian")
        fn.restype = C.c_int
        rc = fn(C.byref(ocp), C.c_int({"forward": 0, "central": 1}[method]), _d(x), _d(u), _d(m), _d(H))
        if rc != 0:
            raise RuntimeError(f"{self.prefix}dynamics_hessian failed: {rc}")
        return H.T.copy()  # the checkers write column-major

    def plant_step(self, ocp, x, u, dt, integrator="euler"):
        """SimulatedPlant::control for a batch of points (dynamics of the descriptor): x [B, nx], u [B, nu] -> x_next [B, nx]"""
        x = np.ascontiguousarray(x, np.float64).reshape(-1, ocp.nx)
        u = np.ascontiguousarray(u, np.float64).reshape(-1, ocp.nu)
        out = np.zeros_like(x)
        fn = getattr(self.lib, self.prefix + "plant_step")
        fn.restype = C.c_int
        rc = fn(C.byref(ocp), C.c_int({"euler": 0, "rk4": 1}[integrator]), C.c_double(dt), C.c_int(x.shape[0]), _d(x), _d(u), _d(out))
        if rc != 0:
            raise RuntimeError(f"{self.prefix}plant_step failed: {rc}")
        return out

    def initial_params(self, ocp, x0, xref=None):
        n = self.dims(ocp).n_params
        x0 = np.ascontiguousarray(x0, np.float64)
        xref = None if xref is None else np.ascontiguousarray(xref, np.float64)
        p = np.zeros(n)
        rc = self._initial_params(C.byref(ocp), _d(x0), _d(xref), _d(p))
        if rc != 0:
            raise RuntimeError(f"{self.prefix}initial_params failed: {rc}")
        return p

    def evaluate(self, ocp, x0, xref=None, params=None, weights=(2.0, 2.0, 2.0), jacobian=True):
        """-> values [m], J dense [m,n], stored-entry pattern [m,n] (bool), parameters after the in-place FD sweep [n]"""
        dm = self.dims(ocp)
        n, m = dm.n_params, dm.m
        x0 = np.ascontiguousarray(x0, np.float64)
        xref = None if xref is None else np.ascontiguousarray(xref, np.float64)
        params = None if params is None else np.ascontiguousarray(params, np.float64)
        values = np.zeros(m)
        jac = np.zeros((m, n)) if jacobian else None
        pat = np.zeros((m, n), np.uint8) if jacobian else None
        after = np.zeros(n) if jacobian else None
        rc = self._evaluate(C.byref(ocp), _d(x0), _d(xref), _d(params), C.c_double(weights[0]), C.c_double(weights[1]), C.c_double(weights[2]),
                            _d(values), _d(jac), None if pat is None else pat.ctypes.data_as(C.POINTER(C.c_uint8)), _d(after))
        if rc != 0:
            raise RuntimeError(f"{self.prefix}evaluate failed: {rc}")
        return values, jac, None if pat is None else pat.astype(bool), after

    def trace(self, ocp, opts, x0, xref=None, params=None, max_events=4096):
        """One instance with the event log: -> dict(params, chi2, status, events=[(type, chi2, vec)])"""
        n = self.dims(ocp).n_params
        x0 = np.ascontiguousarray(x0, np.float64)
        xref = None if xref is None else np.ascontiguousarray(xref, np.float64)
        params = None if params is None else np.ascontiguousarray(params, np.float64)
        out = np.zeros(n)
        chi2 = C.c_double(0)
        status = C.c_int32(0)
        ev_type = np.zeros(max_events, np.int32)
        ev_chi2 = np.zeros(max_events)
        ev_vec = np.zeros((max_events, n))
        n_ev = C.c_int32(0)
        rc = self._trace(C.byref(ocp), C.byref(opts), _d(x0), _d(xref), _d(params), _d(out), C.byref(chi2), C.byref(status), C.c_int(max_events),
                         _i(ev_type), _d(ev_chi2), _d(ev_vec), C.byref(n_ev))
        if rc != 0:
            raise RuntimeError(f"{self.prefix}trace failed: {rc}")
        k = min(n_ev.value, max_events)
        return dict(params=out, chi2=chi2.value, status=status.value, n_events=n_ev.value,
                    events=[(int(ev_type[i]), float(ev_chi2[i]), ev_vec[i].copy()) for i in range(k)])

    def solve_batch(self, ocp, opts, x0, xref=None, params=None, threads=1):
        """-> params [B,n], chi2 [B], status [B], seconds (wall, sum solve, sum prepare)"""
        x0 = np.ascontiguousarray(x0, np.float64)
        B = x0.shape[0]
        n = self.dims(ocp).n_params
        xref = None if xref is None else np.ascontiguousarray(xref, np.float64)
        params = None if params is None else np.ascontiguousarray(params, np.float64)
        out = np.zeros((B, n))
        chi2 = np.zeros(B)
        status = np.zeros(B, np.int32)
        secs = np.zeros(3)
        rc = self._solve_batch(C.byref(ocp), C.byref(opts), C.c_int(B), _d(x0), _d(xref), _d(params), _d(out), _d(chi2), _i(status),
                               C.c_int(threads), _d(secs))
        if rc != 0:
            raise RuntimeError(f"{self.prefix}solve_batch failed: {rc}")
        return out, chi2, status, secs


def _known_answer(lib, fn_name, case_id, stage=0):
    f = getattr(lib, fn_name)
    f.restype = C.c_int
    x, exp = np.zeros(3), np.zeros(3)
    tol, n = C.c_double(0), C.c_int32(0)
    rc = f(C.c_int(case_id), C.c_int(stage), _d(x), _d(exp), C.byref(tol), C.byref(n))
    if rc != 0:
        raise RuntimeError(f"{fn_name}({case_id}) failed: {rc}")
    return x[:n.value], exp[:n.value], tol.value


KNOWN_ANSWER_CASES = [(0, 0), (1, 0), (2, 0), (3, 0), (4, 0), (5, 0), (6, 0), (7, 0), (7, 1)]


class Oracle(_Checker):
    prefix = "sqp_oracle_"
    path = ORACLE_SO

    def solve_sequence(self, ocp, opts, x0, xref=None, n_solves=2):
        """warm-started solves of one instance, new_run only on the first -> (params [n], chi2 [n_solves])"""
        n = self.dims(ocp).n_params
        x0 = np.ascontiguousarray(x0, np.float64)
        xref = None if xref is None else np.ascontiguousarray(xref, np.float64)
        p, chi2 = np.zeros(n), np.zeros(n_solves)
        f = self.lib.sqp_oracle_solve_sequence
        f.restype = C.c_int
        rc = f(C.byref(ocp), C.byref(opts), _d(x0), _d(xref), C.c_int(n_solves), _d(p), _d(chi2))
        if rc != 0:
            raise RuntimeError(f"sqp_oracle_solve_sequence failed: {rc}")
        return p, chi2

    def plant_interval(self, plant_dt, step):
        """interval the reference's plant integrates over at closed-loop step `step` (integer-nanosecond time, (t + dt) - t)"""
        f = self.lib.sqp_oracle_plant_interval
        f.restype = C.c_double
        return float(f(C.c_double(plant_dt), C.c_int(step)))

    def known_answer(self, case_id, stage=0):
        """reference's own LM known-answer tests restated: -> (x, expected, tol)"""
        return _known_answer(self.lib, "sqp_oracle_known_answer", case_id, stage)


class Reference(_Checker):
    prefix = "corbo_ref_"
    path = REF_SO

    def closed_loop(self, ocp, opts, x0, steps):
        x0 = np.ascontiguousarray(x0, np.float64)
        u = np.zeros((steps, ocp.nu))
        x = np.zeros((steps + 1, ocp.nx))
        f = self.lib.corbo_ref_closed_loop
        f.restype = C.c_int
        rc = f(C.byref(ocp), C.byref(opts), _d(x0), C.c_int(steps), _d(u), _d(x))
        if rc != 0:
            raise RuntimeError(f"corbo_ref_closed_loop failed: {rc}")
        return u, x

    def closed_loop_shift(self, ocp, opts, x0, steps):
        """closed loop with the grid's moving-horizon warm start (warmStartShifting) active"""
        x0 = np.ascontiguousarray(x0, np.float64)
        u = np.zeros((steps, ocp.nu))
        x = np.zeros((steps + 1, ocp.nx))
        f = self.lib.corbo_ref_closed_loop_shift
        f.restype = C.c_int
        rc = f(C.byref(ocp), C.byref(opts), _d(x0), C.c_int(steps), _d(u), _d(x))
        if rc != 0:
            raise RuntimeError(f"corbo_ref_closed_loop_shift failed: {rc}")
        return u, x

    def closed_loop_plant(self, ocp, opts, x0, steps, integrator="euler", plant_dt=None, warm_start=False):
        """ClosedLoopControlTask's loop for one instance with the reference's PredictiveController and SimulatedPlant"""
        x0 = np.ascontiguousarray(x0, np.float64)
        u = np.zeros((steps, ocp.nu))
        x = np.zeros((steps + 1, ocp.nx))
        f = self.lib.corbo_ref_closed_loop_plant
        f.restype = C.c_int
        rc = f(C.byref(ocp), C.byref(opts), _d(x0), C.c_int(steps), C.c_int({"euler": 0, "rk4": 1}[integrator]),
               C.c_double(ocp.dt_ref if plant_dt is None else plant_dt), C.c_int(1 if warm_start else 0), _d(u), _d(x))
        if rc != 0:
            raise RuntimeError(f"corbo_ref_closed_loop_plant failed: {rc}")
        return u, x

    def warm_start_shift(self, ocp, x0_old, x0_new, params, xref=None):
        """FullDiscretizationGridBase::warmStartShifting of the compiled reference on one trajectory -> shifted parameters"""
        params = np.ascontiguousarray(params, np.float64)
        out = np.zeros_like(params)
        f = self.lib.corbo_ref_warm_start_shift
        f.restype = C.c_int
        rc = f(C.byref(ocp), _d(np.ascontiguousarray(x0_old, np.float64)), _d(np.ascontiguousarray(x0_new, np.float64)),
               None if xref is None else _d(np.ascontiguousarray(xref, np.float64)), _d(params), _d(out))
        if rc != 0:
            raise RuntimeError(f"corbo_ref_warm_start_shift failed: {rc}")
        return out

    def adapt_once(self, ocp, x, u, dt, n_min, n_max, dt_hyst_ratio=0.1, redundant_controls=None):
        """one isolated call of NonUniformFiniteDifferencesVariableGrid::adaptGrid on a given trajectory (x [N][nx], u [N-1][nu], dt [N-1],
        N = ocp.n_grid) -> (x, u, dt) of the adapted grid.  Strategy TimeBasedSingleStep, or RedundantControls when
        redundant_controls = (num_backup_nodes, epsilon) is given."""
        N = ocp.n_grid
        x = np.ascontiguousarray(x, np.float64).reshape(N, ocp.nx)
        u = np.ascontiguousarray(u, np.float64).reshape(N - 1, ocp.nu)
        dt = np.ascontiguousarray(dt, np.float64).reshape(N - 1)
        cap = max(N, n_max) + 2
        xo, uo, dto = np.zeros((cap, ocp.nx)), np.zeros((cap, ocp.nu)), np.zeros(cap)
        n = C.c_int32(0)
        if redundant_controls is None:
            f = self.lib.corbo_ref_adapt_once
            f.restype = C.c_int
            rc = f(C.byref(ocp), C.c_int(n_min), C.c_int(n_max), C.c_double(dt_hyst_ratio), _d(x), _d(u), _d(dt), _d(xo), _d(uo), _d(dto), C.byref(n))
        else:
            f = self.lib.corbo_ref_adapt_once_redundant
            f.restype = C.c_int
            rc = f(C.byref(ocp), C.c_int(n_min), C.c_int(n_max), C.c_int(redundant_controls[0]), C.c_double(redundant_controls[1]), _d(x), _d(u), _d(dt),
                   _d(xo), _d(uo), _d(dto), C.byref(n))
        if rc != 0:
            raise RuntimeError(f"corbo_ref_adapt_once failed: {rc}")
        n = n.value
        return xo[:n].copy(), uo[:n - 1].copy(), dto[:n - 1].copy()

    def adaptive_steps(self, ocp, opts, x0_seq, xref, n_min, n_max, dt_hyst_ratio=0.1, warm_start=True, num_ocp_iterations=2, redundant_controls=None):
        """Time-optimal MPC with the reference's grid adaptation (NonUniformFiniteDifferencesVariableGrid::adaptGridTimeBasedSingleStep)
        for one instance: x0_seq [steps][nx] -> (n_trace [steps][num_ocp_iterations], u0 [steps][nu], x [N][nx], u [N-1][nu], dt [N-1])"""
        x0_seq = np.ascontiguousarray(x0_seq, np.float64).reshape(-1, ocp.nx)
        steps = x0_seq.shape[0]
        n_trace = np.zeros((steps, num_ocp_iterations), np.int32)
        u0 = np.zeros((steps, ocp.nu))
        cap = max(n_max, ocp.n_grid) + 2
        x = np.zeros((cap, ocp.nx))
        u = np.zeros((cap, ocp.nu))
        dt = np.zeros(cap)
        tail = (C.c_int(1 if warm_start else 0), C.c_int(num_ocp_iterations), C.c_int(steps), _d(x0_seq), _d(np.ascontiguousarray(xref, np.float64)),
                _i(n_trace), _d(u0), _d(x), _d(u), _d(dt))
        if redundant_controls is None:
            f = self.lib.corbo_ref_adaptive_steps
            f.restype = C.c_int
            rc = f(C.byref(ocp), C.byref(opts), C.c_int(n_min), C.c_int(n_max), C.c_double(dt_hyst_ratio), *tail)
        else:  # (num_backup_nodes, epsilon): setGridAdaptRedundantControls
            f = self.lib.corbo_ref_adaptive_steps_redundant
            f.restype = C.c_int
            rc = f(C.byref(ocp), C.byref(opts), C.c_int(n_min), C.c_int(n_max), C.c_int(redundant_controls[0]), C.c_double(redundant_controls[1]), *tail)
        if rc != 0:
            raise RuntimeError(f"corbo_ref_adaptive_steps failed: {rc}")
        n = int(n_trace[-1, -1])
        return n_trace, u0, x[:n].copy(), u[:n - 1].copy(), dt[:n - 1].copy()

    def known_answer(self, case_id, stage=0):
        return _known_answer(self.lib, "corbo_ref_known_answer", case_id, stage)

    def hardware_threads(self):
        return int(self.lib.corbo_ref_hardware_threads())

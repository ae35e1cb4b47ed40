"""compute-sanitizer target: small batched solves through every kernel variant family (lean / general features, partitioned and
twisted factorisation, variable-dt grid, shooting grid, final-stage constraints) -- used with --tool memcheck / racecheck / initcheck."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from control_box_rst_b200 import problems, solver  # noqa: E402

CASES = [
    ("vdp50 lean T=8 (partitioned)", problems.van_der_pol(50), 8, (2.0, 2.0, 2.0)),
    ("vdp12 lean T=8 (twisted, K < 2T)", problems.van_der_pol(12), 8, (2.0, 2.0, 2.0)),
    ("vdp30 ball general T=4", problems.van_der_pol(30, terminal_ball=((2.0, 0.5), 0.01)), 4, (2.0, 3.0, 2.0)),
    ("vdp30 eq xf partly fixed T=2", problems.van_der_pol(30, terminal_equality=(0.1, -0.05), xf_fixed=(1, 0)), 2, (2.0, 2.0, 2.0)),
    ("unicycle20 T=0", problems.unicycle_time_optimal(20), 0, (2.0, 2.0, 2.0)),
    ("cartpole20 T=0", problems.cart_pole_shooting(20), 0, (10.0, 10.0, 10.0)),
]
for name, ocp, T, w in CASES:
    B = 70  # ragged: not a multiple of 32
    x0, xref = problems.instance_data(ocp, B, seed=3)
    lm = solver.BatchedLevenbergMarquardt(ocp, B)
    lm.setIterations(4)
    lm.setPenaltyWeights(*w)
    lm.set_threads_per_instance(T)
    lm.set_problem_data(x0, xref)
    lm.initialize_trajectories()
    status, chi2 = lm.solve(new_run=True)
    lm.evaluate(w)
    if ocp.grid == 0:
        lm.warm_start_shift(x0 + 0.01)
    lm.get_first_controls()
    if ocp.grid != 1:  # the whole closed loop on the device (shift mode: search + out-of-place move kernels, plant kernel with the log)
        lm.closed_loop(x0, 3, xref=xref, mode=2, integrator="rk4")
    lm.closed_loop(x0, 2, xref=xref, mode=1, integrator="euler")
    print(name, "chi2[0] =", chi2[0], "finite:", bool(np.isfinite(chi2).all()), flush=True)
    lm.clear()
A, B_ = solver.linearize_dynamics(0, [1.0], np.ones((70, 2)), np.ones((70, 1)), "central")
print("linearize ok", A.shape)
H = solver.dynamics_hessian(0, [1.0], np.ones((70, 2)), np.ones((70, 1)), np.ones((70, 2)), "central")
xn = solver.plant_step(6, [1.0, 9.81, 0.01, 0.01, 0.02], np.zeros((70, 12)), np.ones((70, 4)), 0.05, "rk4")
print("hessian / plant ok", H.shape, xn.shape)
# quadrotor: warp-cooperative pipeline (TMA-prefetched factor kernel, DMMA tiles), small horizon
ocp = problems.quadrotor(8)
x0, xref = problems.instance_data(ocp, 37, seed=3)
lm = solver.BatchedLevenbergMarquardt(ocp, 37)
lm.setIterations(2)
lm.set_problem_data(x0, xref)
lm.initialize_trajectories()
status, chi2 = lm.solve(new_run=True)
print("quadrotor8 pipeline chi2[0] =", chi2[0], flush=True)
lm.clear()
# pinned host buffers: ingest / export kernels reading and writing host memory in place
import torch  # noqa: E402

ocp = problems.van_der_pol(20)
B = 70
x0, xref = problems.instance_data(ocp, B, seed=3)
lm = solver.BatchedLevenbergMarquardt(ocp, B)
lm.setIterations(3)
hx, hr = torch.from_numpy(x0.copy()).pin_memory(), torch.from_numpy(xref.copy()).pin_memory()
hu, hc, hs = torch.zeros((B, 1), dtype=torch.float64).pin_memory(), torch.zeros(B, dtype=torch.float64).pin_memory(), torch.zeros(B, dtype=torch.int32).pin_memory()
lm.mpc_step_raw(0, hx.data_ptr(), hr.data_ptr(), hu.data_ptr(), hc.data_ptr(), hs.data_ptr())
lm.mpc_step_raw(2, hx.data_ptr(), 0, hu.data_ptr(), hc.data_ptr(), hs.data_ptr())
print("pinned mpc_step ok", float(hc[0]))
lm.clear()

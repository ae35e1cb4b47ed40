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
    print(name, "chi2[0] =", chi2[0], "finite:", bool(np.isfinite(chi2).all()), flush=True)
    lm.clear()
A, B_ = solver.linearize_dynamics(0, [1.0], np.ones((70, 2)), np.ones((70, 1)), "central")
print("linearize ok", A.shape)

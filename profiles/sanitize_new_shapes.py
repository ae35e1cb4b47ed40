"""compute-sanitizer target for the stage-block shapes added with the rest of the benchmark systems and the further compiled
combinations: nx=3/nu=1 (rocket: block 4, time-optimal block 5 with state bounds and a partially fixed goal), time-optimal cart-pole
(block 6), shooting unicycle / cart-pole (Euler), ragged batch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np  # noqa: E402

import cases  # noqa: E402
from control_box_rst_b200 import problems, solver  # noqa: E402

for name in ("rocket20_cn", "rocket20_timeopt", "cartpole20_timeopt", "unicycle20_ms_rk4", "cartpole20_ms_euler", "toy20_cn", "unicycle20_backward"):
    make, w, _ = cases.CASES[name]
    ocp = make()
    B = 70  # ragged: not a multiple of 32
    x0, xref = problems.instance_data(ocp, B, seed=3)
    lm = solver.BatchedLevenbergMarquardt(ocp, B)
    lm.setIterations(3)
    lm.setPenaltyWeights(*w)
    lm.set_problem_data(x0, xref)
    lm.initialize_trajectories()
    status, chi2 = lm.solve(new_run=True)
    lm.evaluate(w)
    lm.get_first_controls()
    print(name, "chi2[0] =", chi2[0], "finite:", bool(np.isfinite(chi2).all()), flush=True)
    lm.clear()
xn = solver.plant_step(7, [], np.ones((70, 3)), np.ones((70, 1)), 0.05, "rk4")
A, B_ = solver.linearize_dynamics(7, [], np.ones((70, 3)), np.ones((70, 1)), "central")
print("rocket plant / linearize ok", xn.shape, A.shape)

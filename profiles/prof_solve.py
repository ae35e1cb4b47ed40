"""ncu target: a few cold batched solves of one workload (not a bench; numbers printed under ncu are never reported).
usage: python profiles/prof_solve.py [--cfg 1] [--batch 4096] [--T 0] [--solves 3]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from control_box_rst_b200 import problems, solver  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cfg", type=int, default=1)
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--T", type=int, default=0)
ap.add_argument("--solves", type=int, default=3)
ap.add_argument("--precision", default="f64")
a = ap.parse_args()
ocp, kw, _ = problems.config(a.cfg)
x0, xref = problems.instance_data(ocp, a.batch, seed=1234 + a.cfg)
lm = solver.BatchedLevenbergMarquardt(ocp, a.batch)
lm.setIterations(kw["iterations"])
lm.setPenaltyWeights(*kw["weights"])
lm.set_problem_data(x0, xref)
lm.set_threads_per_instance(a.T)
if a.precision != "f64":
    lm.set_precision(a.precision)
for _ in range(a.solves):
    lm.initialize_trajectories()
    lm.solve(new_run=True, fetch=False)
    lm.synchronize()
print("kernel ms", lm.last_solve_ms())
lm.clear()

import sys, numpy as np
sys.path.insert(0, ".")
from control_box_rst_b200 import problems, solver
ocp, kw, _ = problems.config(4)
B = 4096
x0, xref = problems.instance_data(ocp, B, seed=1238)
res = {}
for prec in ("f64", "f32"):
    lm = solver.BatchedLevenbergMarquardt(ocp, B)
    lm.setIterations(10); lm.set_precision(prec); lm.set_problem_data(x0, xref)
    ts = []
    for rep in range(3):
        lm.initialize_trajectories(); lm.solve(new_run=True, fetch=False); lm.synchronize(); ts.append(lm.last_solve_ms())
    st = lm.statistics(); p = lm.get_params()
    res[prec] = p
    print(prec, "ms", min(ts), "inner", st["inner_passes"].mean(), st["inner_passes"].max(), "rejects", st["rejects"].mean(), "lin", st["relinearizations"].mean())
    lm.clear()
err = np.abs(res["f32"] - res["f64"]).max(axis=1) / np.maximum(1, np.abs(res["f64"]).max(axis=1))
print("f32 vs f64 traj err median %.2e p99 %.2e max %.2e" % (np.median(err), np.percentile(err, 99), err.max()))

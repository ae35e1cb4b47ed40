"""Ad-hoc timing probe (not a test, not the bench): device time of one batched solve for a few batch sizes."""
import sys
import numpy as np
from control_box_rst_b200 import problems, solver

Bs_arg = tuple(int(a) for a in sys.argv[1:]) or (1024, 4096, 16384, 65536)
for cfg, Bs in ((1, Bs_arg),):
    ocp, kw, _ = problems.config(cfg)
    for B in Bs:
        x0, xref = problems.instance_data(ocp, B)
        lm = solver.BatchedLevenbergMarquardt(ocp, B)
        lm.setIterations(10)
        lm.setPenaltyWeights(*kw["weights"])
        lm.set_problem_data(x0, xref)
        ts = []
        for rep in range(5):
            lm.initialize_trajectories()
            lm.solve(new_run=True, fetch=False)
            lm.synchronize()
            ts.append(lm.last_solve_ms())
        st = lm.statistics()
        t = min(ts)
        d = lm.dims
        print(f"cfg{cfg} B={B} ms={t:.3f} (all {['%.3f' % x for x in ts]}) iters/s={B*10/t*1e3:.3e} "
              f"roofline_frac={B*10/t*1e3*d.algorithmic_bytes_per_iteration/6.45e12:.4f} inner/inst={st['inner_passes'].mean():.2f} "
              f"rejects/inst={st['rejects'].mean():.2f} lin/inst={st['relinearizations'].mean():.2f}")
        lm.clear()

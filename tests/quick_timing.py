"""Ad-hoc timing probe (not a test, not the bench): device time of one batched solve for a few batch sizes / thread mappings.
usage: python tests/quick_timing.py [B ...] [--T 1,2,4,8,0]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from control_box_rst_b200 import problems, solver

args = sys.argv[1:]
Ts = (0,)
if "--T" in args:
    j = args.index("--T")
    Ts = tuple(int(x) for x in args[j + 1].split(","))
    args = args[:j] + args[j + 2:]
cfg = 1
if "--cfg" in args:
    j = args.index("--cfg")
    cfg = int(args[j + 1])
    args = args[:j] + args[j + 2:]
Bs = tuple(int(a) for a in args) or (1024, 4096, 16384, 65536)
ocp, kw, _ = problems.config(cfg)
for B in Bs:
    x0, xref = problems.instance_data(ocp, B)
    lm = solver.BatchedLevenbergMarquardt(ocp, B)
    lm.setIterations(10)
    lm.setPenaltyWeights(*kw["weights"])
    lm.set_problem_data(x0, xref)
    ref = None
    for T in Ts:
        lm.set_threads_per_instance(T)
        ts = []
        for rep in range(4):
            lm.initialize_trajectories()
            lm.solve(new_run=True, fetch=False)
            lm.synchronize()
            ts.append(lm.last_solve_ms())
        st = lm.statistics()
        lm.set_phase_profile(True)
        lm.initialize_trajectories()
        lm.solve(new_run=True, fetch=False)
        lm.synchronize()
        pc = lm.phase_cycles()
        lm.set_phase_profile(False)
        p = lm.get_params()
        if ref is None:
            ref = p
        t = min(ts)
        d = lm.dims
        print(f"cfg{cfg} B={B} T={T} ms={t:.3f} iters/s={B*10/t*1e3:.3e} "
              f"roofline_frac={B*10/t*1e3*d.algorithmic_bytes_per_iteration/6.45e12:.4f} inner/inst={st['inner_passes'].mean():.2f} "
              f"rejects/inst={st['rejects'].mean():.2f} lin/inst={st['relinearizations'].mean():.2f} "
              f"maxdiff_vs_first_T={np.abs(p-ref).max():.2e} phase_kcycles={ {k: round(v/1e3,1) for k,v in pc.items()} }", flush=True)
    lm.clear()

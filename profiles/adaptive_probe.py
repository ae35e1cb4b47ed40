"""Timing probe of the grid-adaptation front-end (b200sqp_adaptive_step): per-step host time as the grids spread over the buckets.
usage: adaptive_probe.py [batch] [spread|same] [ocp_iterations]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from control_box_rst_b200 import problems, solver  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
mode = sys.argv[2] if len(sys.argv) > 2 else "spread"
m = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ocp, kw, _ = problems.config(2)
rng = np.random.default_rng(77)
r = rng.uniform(0.5, 9.0, B) if mode == "spread" else np.full(B, 4.9)  # 4.9 m at 1 m/s over 49 intervals of 0.1 s: the grid stays
ang = rng.uniform(-1.0, 1.0, B)
x0 = np.zeros((B, 3))
xf = np.stack([r * np.cos(ang), r * np.sin(ang), ang + rng.uniform(-0.5, 0.5, B)], axis=1)
ad = solver.AdaptiveGridBatch(ocp, B, 5, 100, 0.1, warm_start=True)
ad.setIterations(kw["iterations"])
ad.setPenaltyWeights(*kw["weights"])
t0 = time.perf_counter()
ad.reserve()
print(f"reserve {time.perf_counter() - t0:.3f} s")
for it in range(12):
    t0 = time.perf_counter()
    u0, chi2, status, n = ad.step(x0 + 0.01 * it * (xf - x0), xf, num_ocp_iterations=m)
    dt = time.perf_counter() - t0
    st = ad.statistics()
    print(f"step {it}: {dt * 1e3:8.2f} ms = {B * kw['iterations'] * m / dt:.3e} iters/s  N {n.min()}..{n.max()} mean {n.mean():.1f}  "
          f"buckets {st['occupied_buckets']}  launches {st['launches']}")

"""Timing probe of the grid-adaptation front-end (b200sqp_adaptive_step): per-step host time as the grids spread over the buckets."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from control_box_rst_b200 import problems, solver  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reserve = len(sys.argv) <= 2 or sys.argv[2] != "lazy"
ocp, kw, _ = problems.config(2)
rng = np.random.default_rng(77)
r, ang = rng.uniform(0.5, 9.0, B), rng.uniform(-1.0, 1.0, B)
x0 = np.zeros((B, 3))
xf = np.stack([r * np.cos(ang), r * np.sin(ang), ang + rng.uniform(-0.5, 0.5, B)], axis=1)
ad = solver.AdaptiveGridBatch(ocp, B, 5, 100, 0.1, warm_start=True)
ad.setIterations(kw["iterations"])
ad.setPenaltyWeights(*kw["weights"])
t0 = time.perf_counter()
if reserve:
    ad.reserve()
print(f"reserve {time.perf_counter() - t0:.3f} s")
for it in range(12):
    t0 = time.perf_counter()
    u0, chi2, status, n = ad.step(x0 + 0.01 * it * (xf - x0), xf, num_ocp_iterations=3)
    dt = time.perf_counter() - t0
    st = ad.statistics()
    print(f"step {it}: {dt * 1e3:8.2f} ms  N {n.min()}..{n.max()} mean {n.mean():.1f}  buckets {st['occupied_buckets']}  launches {st['launches']}")

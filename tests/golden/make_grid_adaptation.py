"""Generates tests/golden/grid_adaptation.npz from the compiled reference (oracle/_ref, needs /root/reference at build time):
the adaptive time-optimal MPC loop of tests/test_gpu_grid_adaptation.py for every case / warm-start mode / strategy -- per instance the
grid size after every controller step, the first controls and the final trajectories (padded), and whether the reference survived.

    python tests/golden/make_grid_adaptation.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_grid_adaptation as T  # noqa: E402
from oracle import bindings  # noqa: E402

ref = bindings.Reference()
out = {}
for name, warm, redundant in T.GOLDEN_RUNS:
    inputs = T._inputs(name)
    expected = T._reference_run(ref, *inputs, warm, T.M, redundant)
    out.update(T._pack(T._key(name, warm, redundant), expected, inputs[0]))
    print(T._key(name, warm, redundant), "instances without a reference answer:", sum(e is None for e in expected))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "grid_adaptation.npz"), **out)

"""Noise-floor equivalence for the trigonometric models (cart-pole, unicycle, quadrotor).

The reference's Jacobians are central differences with delta = 1e-9, so its own result is one realisation of ~5e-7 relative noise
in every Jacobian entry: changing the start state by ONE ULP changes the converged trajectory by far more than 1e-6 for these
models.  A trajectory tolerance for them can therefore not be a constant picked by hand -- it has to be the reference's own
sensitivity.  This test measures both distributions in the same run, on the same instances, with the same sample size:

    floor  = | checker(x0 + k ulp) - checker(x0 + (k-1) ulp) |,  k = 1..4   (the checker against itself one ulp apart)
    device = | device(x0 + (k-1) ulp) - checker(x0 + (k-1) ulp) |, k = 1..4

with |.| = max-norm of the trajectory difference relative to max(1, |trajectory|), 4 x 256 samples each, and asserts that the
device error has the SAME distribution as the floor: median, 95th and 99th percentile and the maximum within a factor 2 (a fixed
additive 1e-9 covers quantities that are exactly reproducible).  chi2 is compared the same way.  The checker is the compiled
reference when oracle/_ref/libcorbo_ref.so is present (it travels to the GPU box), else the oracle port.
"""
import numpy as np
import pytest

from control_box_rst_b200 import _abi as abi
from control_box_rst_b200 import problems, solver
from oracle import bindings

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def checker(oracle):
    return bindings.Reference() if bindings.Reference.available() else oracle


def _rel(p, q):
    return np.abs(p - q).max(axis=1) / np.maximum(1.0, np.abs(q).max(axis=1))


CASES = [
    ("cartpole40_rk4", lambda: problems.cart_pole_shooting(40), (10.0, 10.0, 10.0), 256),
    ("unicycle30_timeopt", lambda: problems.unicycle_time_optimal(30), (2.0, 2.0, 2.0), 256),
    ("quadrotor12_cn", lambda: problems.quadrotor(12), (2.0, 2.0, 2.0), 256),
    ("vdp50_cn", lambda: problems.van_der_pol(50), (2.0, 2.0, 2.0), 256),  # the polynomial model for comparison: same statement holds
]
SHIFTS = 4


@pytest.mark.parametrize("name,make,weights,B", CASES, ids=[c[0] for c in CASES])
def test_device_error_is_the_references_own_noise_floor(checker, name, make, weights, B):
    ocp = make()
    iters = 10
    x0, xref = problems.instance_data(ocp, B, seed=2024)
    opts = abi.LmOptions.defaults(iterations=iters, weights=weights)
    starts = [x0]
    for _ in range(SHIFTS):
        starts.append(np.nextafter(starts[-1], np.inf))  # every component one more ulp up
    ref = [checker.solve_batch(ocp, opts, xs, xref, threads=8)[:2] for xs in starts]
    lm = solver.BatchedLevenbergMarquardt(ocp, B)
    lm.setIterations(iters)
    lm.setPenaltyWeights(*weights)
    floor, floor_chi2, gpu, gpu_chi2 = [], [], [], []
    for k in range(1, SHIFTS + 1):
        floor.append(_rel(ref[k][0], ref[k - 1][0]))
        floor_chi2.append(np.abs(ref[k][1] / ref[k - 1][1] - 1.0))
        lm.set_problem_data(starts[k - 1], xref)
        lm.initialize_trajectories()
        _, chi2 = lm.solve(new_run=True)
        gpu.append(_rel(lm.get_params(), ref[k - 1][0]))
        gpu_chi2.append(np.abs(chi2 / ref[k - 1][1] - 1.0))
    lm.clear()

    def stats(e):
        e = np.concatenate(e)
        return np.array([np.median(e), np.percentile(e, 95), np.percentile(e, 99), e.max()])

    s_floor, s_gpu, s_fc, s_gc = stats(floor), stats(gpu), stats(floor_chi2), stats(gpu_chi2)
    print(f"{name} ({type(checker).__name__}): trajectory median/p95/p99/max  floor {s_floor}  device {s_gpu};  chi2 floor {s_fc}  device {s_gc}")
    assert np.all(s_gpu <= 2.0 * s_floor + 1e-9), (name, s_gpu, s_floor)
    assert np.all(s_gc <= 2.0 * s_fc + 1e-12), (name, s_gc, s_fc)

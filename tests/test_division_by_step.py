"""The device evaluates x / dt as q = RN(x z), r = x - q dt (one fma, exact), RN(q + r z) with z = RN(1/dt)
(control_box_rst_b200/csrc/dynamics.cuh StepSize) instead of the reference's IEEE division
(numerics/include/corbo-numerics/finite_differences_collocation.h:119-240, `(x2 - x1) / dt`).  This test restates the
three-operation sequence in exact rational arithmetic (Fraction -> float conversion is correctly rounded) and checks it against
the correctly rounded quotient on random and on hard-to-round dividends (quotients next to rounding midpoints)."""
import math
import random
from fractions import Fraction

import pytest


def rn(fr):
    return float(fr)  # int/int true division inside Fraction.__float__ is correctly rounded


def fma(a, b, c):
    return rn(Fraction(a) * Fraction(b) + Fraction(c))


def div_by_step(x, dt, z):
    q = x * z
    r = fma(-q, dt, x)
    return fma(r, z, q)


def ulp(v):
    return math.ulp(v)


@pytest.mark.parametrize("dt", [0.1, 0.05, 0.02, 0.1 + 1e-9, 0.1 - 1e-9, 0.3141592653589793, 1.0 / 3.0, 0.9999999999999999, 1e-3])
def test_three_operation_quotient_is_correctly_rounded(dt):
    rng = random.Random(int(dt * 1e12) & 0xFFFFFFFF)
    z = 1.0 / dt
    assert z == rn(Fraction(1) / Fraction(dt))
    cases = []
    for _ in range(4000):
        cases.append(rng.uniform(-4, 4) * 10.0 ** rng.randint(-12, 2))
    for _ in range(4000):
        # quotient next to a rounding midpoint: x ~ (m + 1/2 ulp) * dt
        qm = rng.uniform(1, 2) * 2.0 ** rng.randint(-30, 10)
        mid = Fraction(qm) + Fraction(ulp(qm)) / 2
        x = rn(mid * Fraction(dt))
        cases.extend([x, math.nextafter(x, math.inf), math.nextafter(x, -math.inf), -x])
    for _ in range(2000):
        # exactly representable quotients and their neighbours
        qe = float(rng.randint(1, 1 << 40)) * 2.0 ** rng.randint(-60, 0)
        x = rn(Fraction(qe) * Fraction(dt))
        cases.extend([x, math.nextafter(x, math.inf)])
    cases.extend([0.0, 1e-9, -1e-9, 2e-9])
    bad = 0
    for x in cases:
        want = rn(Fraction(x) / Fraction(dt))
        got = div_by_step(x, dt, z)
        bad += got != want
    assert bad == 0, f"{bad} of {len(cases)} quotients differ from RN(x/dt) for dt={dt!r}"

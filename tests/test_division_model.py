"""Executable form of the exactness claim behind `div_pair` (hope_kernels.cu; used by the LiDAR raycast and the
Reeds-Shepp trajectory check, where the reference divides two numerators by the same determinant,
lidar_simulator.py:112-113, car_parking_base.py:512-513): with r = RN(1/den), a = RN(n r), the sequence
    q = fma(fma(-den, a, n), r, a)
returns the IEEE-754 quotient RN(n / den).  Modelled with exact rationals (every fma rounded once), over operands of
the magnitudes the env produces plus adversarial ones (significands next to a power of two, quotients next to a
rounding boundary)."""
from fractions import Fraction as F

import numpy as np


def _rn(x):
    return float(x)   # Fraction -> nearest double, ties to even


def _div_pair_model(n, den):
    r = _rn(F(1) / F(den))            # __drcp_rn
    a = n * r                         # __dmul_rn
    res = _rn(F(n) - F(den) * F(a))   # __fma_rn(-den, a, n): exact residual, rounded once
    return _rn(F(res) * F(r) + F(a))  # __fma_rn(res, r, a)


def test_shared_reciprocal_division_is_the_ieee_quotient():
    rng = np.random.default_rng(1)
    cases = []
    for _ in range(12000):            # lidar / box-edge scale
        cases.append((float(rng.uniform(-200, 200)), float(rng.uniform(-30, 30))))
    for _ in range(12000):            # wide dynamic range
        cases.append((float(rng.standard_normal() * 10.0 ** rng.integers(-12, 6)), float(rng.standard_normal() * 10.0 ** rng.integers(-12, 6))))
    for _ in range(8000):             # denominators just below a power of two, numerators one ulp off a multiple
        den = float(np.nextafter(2.0 ** int(rng.integers(-8, 8)), 0.0)) * float(rng.choice([1, -1]))
        for _k in range(int(rng.integers(0, 4))):
            den = float(np.nextafter(den, 0.0))
        n = float(np.nextafter(den * float(rng.integers(1, 2000)), rng.choice([-np.inf, np.inf])))
        cases.append((n, den))
    for _ in range(8000):             # quotients next to rounding boundaries
        den, q = float(rng.uniform(0.5, 8)), float(rng.uniform(-10, 10))
        n = den * q
        cases += [(n, den), (float(np.nextafter(n, np.inf)), den), (float(np.nextafter(n, -np.inf)), den)]
    checked = 0
    for n, den in cases:
        if den == 0.0:
            continue
        assert _div_pair_model(n, den) == n / den, (n.hex(), den.hex())
        checked += 1
    assert checked > 50000

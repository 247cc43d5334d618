"""Loop constants computed on the host exactly as the reference does and passed into the
library: Common/calcLoopCoef.m:41-45, Common/calcLoopCoefCarr.m:41-56,
BDS-3_B1C/include/CalcWeighingFactor.m:43-81."""
from __future__ import annotations

import math

import numpy as np


def calcLoopCoef(LBW, zeta, k):
    Wn = LBW * 8 * zeta / (4 * zeta ** 2 + 1)
    return k / (Wn * Wn), 2.0 * zeta / Wn


def calcLoopCoefCarr(settings):
    Wn = 1.2 * settings.pllNoiseBandwidth
    T = settings.intTime
    return Wn ** 3 * T ** 2, 2 * Wn ** 2 * T, 2 * Wn   # pf3, pf2, pf1


_GLX, _GLW = np.polynomial.legendre.leggauss(48)


def _integrate(fn, a, b, breaks):
    """Composite 48-point Gauss-Legendre between the removable singularities of the PSDs."""
    pts = [a] + sorted(p for p in breaks if a < p < b) + [b]
    tot = 0.0
    for lo, hi in zip(pts[:-1], pts[1:]):
        n = max(1, int(math.ceil((hi - lo) / 0.25e6)))
        edges = np.linspace(lo, hi, n + 1)
        for l, h in zip(edges[:-1], edges[1:]):
            x = 0.5 * (h - l) * _GLX + 0.5 * (h + l)
            tot += 0.5 * (h - l) * float(np.dot(_GLW, fn(x)))
    return tot


def CalcWeighingFactor(settings):
    """Data/pilot DLL weight for wide-band tracking: ratio of 11*P*beta^2 of BOC(1,1) to that of the
    29/33 BOC(1,1) + 4/33 BOC(6,1) pilot over the front-end bandwidth FEBW."""
    fc = settings.codeFreqBasis
    Tc = 1 / fc
    Br = settings.FEBW

    def g(f, m):  # BOC(m/… ) PSD shape used by the reference: m = 2 -> BOC(1,1), m = 12 -> BOC(6,1)
        return Tc * (np.sin(np.pi / m * f / fc) * np.sin(np.pi * f / fc) / np.cos(np.pi / m * f / fc) * fc / f / np.pi) ** 2

    g11 = lambda f: g(f, 2)
    gp = lambda f: 29 / 33 * g(f, 2) + 4 / 33 * g(f, 12)
    # cos(pi/2 f/fc) = 0 at odd multiples of fc; cos(pi/12 f/fc) = 0 at odd multiples of 6 fc; f = 0 is 0/0
    br = [0.0] + [s * k * fc for k in range(1, 60, 2) for s in (-1, 1)] + [s * k * 6 * fc for k in (1, 3, 5) for s in (-1, 1)]
    lo, hi = -Br / 2, Br / 2
    P11_2 = _integrate(lambda f: g11(f) * f * f, lo, hi, br)
    P11 = _integrate(g11, lo, hi, br)
    Pp_2 = _integrate(lambda f: gp(f) * f * f, lo, hi, br)
    Pp = _integrate(gp, lo, hi, br)
    t1 = 11 * P11 * (P11_2 / P11)
    t2 = 33 * Pp * (Pp_2 / Pp)
    return t1 / (t1 + t2)

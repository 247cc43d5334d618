"""CPU model of the chip-synchronous B2a correlator (csrc/bds_track_b2a.cuh) against the float64 oracle.

The kernel only runs on a GPU.  Its decision logic is small enough to restate in Python line by line: the per-epoch
table (thresholds, rank sort, masks, bins: fastb_build_tab_warp), the per-unit bookkeeping of fastb_unit (first
sample, sub-sample phase, rank search, guard band, block-edge checks), the rotated code-bit array (b2a_load_bits /
fastb_code12), the generated body + combination (executed from the .inc as in test_fast_body_emulation.py), the unit
rotation and the B2a I/Q convention.  One whole epoch evaluated that way must
  * cover every sample of the block exactly once (fast units + exact-path units + the t = 0 sample), and
  * reproduce the oracle's twelve sums (tracking.m:260-331) within the parity tolerance 1e-4.
A sign, index or boundary mistake in the design shows up here, before any GPU time is spent on the kernel."""
import math

import numpy as np
import pytest

import bds_oracle as O
from test_fast_body_emulation import GENB, _s32  # noqa: F401  (generated body runner)

FS, FC, L = 99.375e6, 10.23e6, 10230
NSEG, CHIPS, UNITS = 20, 10, 1023
GUARD = 16
TWO32, TWO64 = 1 << 32, 1 << 64


class Settings:
    dllCorrelatorSpacing = 0.5
    codeLength = L
    samplingFreq = FS
    pilotTRKflag = 1


def build_tab(rem, step, carrFreq, remCarr):
    """fastb_build_tab_warp"""
    S = 1.0 / (2.0 * step)
    r = carrFreq / FS
    r -= math.floor(r)
    dphi = int(round(r * TWO64)) % TWO64
    r0 = remCarr / 6.283185307179586476925286766559
    r0 -= math.floor(r0)
    phi0 = int(round(r0 * TWO64)) % TWO64
    w = []
    for t in range(4 * (GENB.nwords + 1)):
        ph = (t * dphi) % TWO64
        hi = ph >> 32
        a = (hi - TWO32 if hi & 0x80000000 else hi) * 4.656612873077392578125e-10 * math.pi
        w.append((int(round(math.cos(a) * 32767.0)), int(round(-math.sin(a) * 32767.0))))
    thr, ok = [], True
    for k in range(1, NSEG + 1):
        th = GENB.beta[k] * S - GENB.R[k]
        ok &= 1e-6 < th < 1 - 1e-6
        th = min(max(th, 0.0), 1.0)
        thr.append(int(min(th * 4294967296.0, 4294967295.0)))
    pos = [sum((thr[j] < v) or (thr[j] == v and j < t) for j in range(NSEG)) for t, v in enumerate(thr)]
    srt = [0xFFFFFFFF] * 24
    for t, v in enumerate(thr):
        srt[pos[t]] = v
    mask = [sum(1 << (k - 1) for k in range(1, NSEG + 1) if pos[k - 1] >= j) for j in range(NSEG + 1)]
    bins = [sum((v >> 25) < t for v in thr) for t in range(129)]
    ok &= all(sum((v >> 25) == t for v in thr) <= 4 for t in range(129))
    return dict(w=w, thr=srt, mask=mask, bins=bins, u0=2.0 * rem, S=S, dphi=dphi, phi0=phi0, valid=ok)


def rotated_bits(code_pm):
    """b2a_load_bits: ext bit n = chip n-1 (bit set <=> chip is -1), ext bit 0 = last chip, ext bit 10231 = first chip"""
    neg = [1 if v < 0 else 0 for v in code_pm]
    ext = [neg[-1]] + neg + [neg[0]]
    return ext


def code12(ext, u):
    return sum(ext[CHIPS * u + b] << b for b in range(12))


def exact_range(x, codes, rem, step, carrFreq, remCarr, k0, k1, lo, hi, acc, cover):
    """fastb_exact_range: per-sample evaluation with the oracle's own expressions"""
    n = x.size
    d = 0.5
    t = {nm: O.colon((rem + off) if off else rem, step, ((n - 1) * step + rem + off) if off else ((n - 1) * step + rem), n)
         for nm, off in (("E", -d), ("P", 0.0), ("L", d))}
    for k in range(k0, k1 + 1):
        ip = int(math.ceil(t["P"][k]))
        if ip < lo or ip > hi:
            continue
        cover[k] += 1
        th = carrFreq * 2.0 * math.pi * (k / FS) + remCarr
        qB, iB = x[k] * math.cos(th), x[k] * math.sin(th)
        for nm in ("E", "P", "L"):
            idx = int(math.ceil(t[nm][k]))
            for fam in ("d", "p"):
                cv = codes[fam][idx]                       # padded [code(end) code code(1)]
                acc[f"{fam}_I_{nm}"] += cv * iB
                acc[f"{fam}_Q_{nm}"] += cv * qB


def model_epoch(x, B0, code_d, code_p, rem, step, carrFreq, remCarr, guard=GUARD):
    """b2a_correlate for one epoch; returns (sums, coverage, n_fast, n_exact)"""
    blk = x.size
    tab = build_tab(rem, step, carrFreq, remCarr)
    codes = {"d": np.concatenate([code_d[-1:], code_d, code_d[:1]]), "p": np.concatenate([code_p[-1:], code_p, code_p[:1]])}
    extd, extp = rotated_bits(code_d), rotated_bits(code_p)
    acc = {f"{fam}_{iq}_{nm}": 0.0 for fam in "dp" for nm in "EPL" for iq in "IQ"}
    cover = np.zeros(blk, dtype=np.int64)
    if rem == 0.0:
        exact_range(x, codes, rem, step, carrFreq, remCarr, 0, 0, -100, 0, acc, cover)
    tileBase = B0 & ~15
    xb = np.concatenate([np.zeros(B0 - tileBase, dtype=np.int64), x.astype(np.int64), np.zeros(256, dtype=np.int64)])
    nfast = nexact = 0
    for u in range(UNITS):
        q = (float(2 * CHIPS * u) - tab["u0"]) * tab["S"]
        nc = int(math.floor(q)) + 1
        psi = nc - q
        Psi = int(min(psi * 4294967296.0, 4294967295.0))
        j = tab["bins"][Psi >> 25]
        for _ in range(4):
            j += tab["thr"][j] < Psi
        mk = tab["mask"][j]
        below = Psi - tab["thr"][j - 1] if j > 0 else Psi
        above = tab["thr"][j] - Psi if j < NSEG else 0xFFFFFFFF - Psi
        exact = (not tab["valid"]) or below <= guard or above <= guard or Psi >= 0xFFFFFFFF - guard
        ln = GENB.R[NSEG] + ((mk >> (NSEG - 1)) & 1)
        if nc < 0 or nc + ln > blk:
            exact = True
        if exact:
            nexact += 1
            qe = (float(2 * CHIPS * (u + 1)) - tab["u0"]) * tab["S"]
            k0, k1 = max(0, nc - 2), min(blk - 1, int(math.floor(qe)) + 3)
            exact_range(x, codes, rem, step, carrFreq, remCarr, k0, k1, CHIPS * u + 1, CHIPS * u + CHIPS, acc, cover)
            continue
        nfast += 1
        cover[nc:nc + ln] += 1
        o = B0 + nc - tileBase
        b = (xb[(o & ~3):(o & ~3) + 4 * (GENB.nwords + 2)] & 0xFF).astype(np.uint64)
        words = [int(b[4 * i] | (b[4 * i + 1] << np.uint64(8)) | (b[4 * i + 2] << np.uint64(16)) | (b[4 * i + 3] << np.uint64(24)))
                 for i in range(GENB.nwords + 2)]
        pk = lambda a, c: (a & 0xFFFF) | ((c & 0xFFFF) << 16)
        w = tab["w"]
        table = [(pk(w[4 * i][0], w[4 * i + 1][0]), pk(w[4 * i + 2][0], w[4 * i + 3][0]),
                  pk(w[4 * i][1], w[4 * i + 1][1]), pk(w[4 * i + 2][1], w[4 * i + 3][1])) for i in range(GENB.nwords + 1)]
        db, pb = code12(extd, u), code12(extp, u)
        env = GENB.run(words, table, 8 * (o & 3), mk,
                       {"FASTB_CD": lambda jj: 1 - 2 * ((db >> (jj + 1)) & 1), "FASTB_CP": lambda jj: 1 - 2 * ((pb >> (jj + 1)) & 1)})
        ph = (tab["phi0"] + nc * tab["dphi"]) % TWO64
        hi = ph >> 32
        ang = (hi - TWO32 if hi & 0x80000000 else hi) * 1.4629180792671596e-9
        rr, ri = math.cos(ang) / 32767.0, -math.sin(ang) / 32767.0
        for fam, F in (("d", "D"), ("p", "P")):
            for nm in "EPL":
                nr, ni = env[F + nm + "r"], env[F + nm + "i"]
                acc[f"{fam}_I_{nm}"] -= nr * ri + ni * rr
                acc[f"{fam}_Q_{nm}"] += nr * rr - ni * ri
    return acc, cover, nfast, nexact


def make_epoch(rng, rem, codeFreq, carrFreq, remCarr):
    step = codeFreq / FS
    blk = int(math.ceil((L - rem) / step))
    code_d = rng.choice([-1, 1], size=L).astype(np.int64)
    code_p = rng.choice([-1, 1], size=L).astype(np.int64)
    k = np.arange(blk)
    t = rem + k * step
    idx = np.clip(np.ceil(t).astype(np.int64) - 1, 0, L - 1)
    th = carrFreq * 2.0 * np.pi * (k / FS) + remCarr
    x = 22.0 * rng.standard_normal(blk) + 18.0 * (code_d[idx] * np.sin(th) + code_p[idx] * np.cos(th))   # SURVEY 8d B2a model
    x = np.clip(np.rint(x), -127, 127).astype(np.int8)
    return x, step, code_d, code_p


@pytest.mark.parametrize("case", [
    dict(rem=0.0, codeFreq=10.23e6, carrFreq=13.55e6 + 1234.5, remCarr=0.0, B0=7),            # first epoch: t = 0 sample
    dict(rem=0.0731, codeFreq=10.23e6 - 17.3, carrFreq=13.55e6 - 4321.0, remCarr=2.5, B0=123456),
    dict(rem=0.0049, codeFreq=10.23e6 + 31.9, carrFreq=13.55e6 + 77.7, remCarr=5.9, B0=99375 * 3 + 11),
])
def test_model_epoch_equals_oracle(case):
    rng = np.random.default_rng(5)
    x, step, code_d, code_p = make_epoch(rng, case["rem"], case["codeFreq"], case["carrFreq"], case["remCarr"])
    acc, cover, nfast, nexact = model_epoch(x, case["B0"], code_d, code_p, case["rem"], step, case["carrFreq"], case["remCarr"])
    assert np.all(cover == 1), (np.flatnonzero(cover != 1)[:5], cover[cover != 1][:5])
    # rem = 0 at exactly the nominal code rate (the first epoch of a B2a channel: preRun gives no code-Doppler aiding) puts
    # (20 u + k) S on an integer for every 1364th half chip: ~18 exact ties, all sent through the exact path
    assert nfast + nexact == UNITS and nexact <= (25 if case["rem"] == 0.0 else 3)
    codes = {"data": np.concatenate([code_d[-1:], code_d, code_d[:1]]), "pilot": np.concatenate([code_p[-1:], code_p, code_p[:1]])}
    ref, _, _ = O.correlate_epoch("B2a", Settings, x.astype(np.float64), codes, case["rem"], step, case["carrFreq"], case["remCarr"])
    for fam in "dp":
        scale = max(abs(ref[f"{fam}_I_P"]), abs(ref[f"{fam}_Q_P"]))
        for nm in "EPL":
            for iq in "IQ":
                k = f"{fam}_{iq}_{nm}"
                assert abs(acc[k] - ref[k]) <= 1e-4 * scale, (k, acc[k], ref[k], scale)


def test_model_epoch_wide_guard_mixes_exact_and_fast_units():
    """a wide guard band (the cfg.reserved test hook) sends many units through the exact path; the seam between the two
    paths must still cover every sample once and give the same sums"""
    rng = np.random.default_rng(9)
    rem, cf, fc_, rc = 0.0387, 10.23e6 + 5.0, 13.55e6 + 2500.0, 1.0
    x, step, code_d, code_p = make_epoch(rng, rem, cf, fc_, rc)
    acc, cover, nfast, nexact = model_epoch(x, 40, code_d, code_p, rem, step, fc_, rc, guard=1 << 24)
    assert np.all(cover == 1)
    assert 0.05 < nexact / UNITS < 0.6
    codes = {"data": np.concatenate([code_d[-1:], code_d, code_d[:1]]), "pilot": np.concatenate([code_p[-1:], code_p, code_p[:1]])}
    ref, _, _ = O.correlate_epoch("B2a", Settings, x.astype(np.float64), codes, rem, step, fc_, rc)
    for k, v in ref.items():
        fam = k[0]
        scale = max(abs(ref[f"{fam}_I_P"]), abs(ref[f"{fam}_Q_P"]))
        assert abs(acc[k] - v) <= 1e-4 * scale, (k, acc[k], v)

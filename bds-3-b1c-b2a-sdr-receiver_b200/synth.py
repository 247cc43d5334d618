"""Synthetic int8 IF records; the reference has no generator.

``make_sats`` draws the scenario (numpy.random.default_rng(20260101) by default);
``synth_numpy`` renders it on the host in float64 (tests, small records);
``synth_device`` renders it on the GPU through ``bds_synth_if`` (bench, 30 s records).
Signal model: see csrc/bds_synth.cu header.  C/N0 -> amplitude:  A = sqrt(2 * 10^(CN0/10) * sigma^2 / (fs/2)).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib as L
from . import codes
from .settings import Struct, samples_per_code


def cn0_to_amplitude(cn0_dbhz, sigma, fs):
    n0 = sigma ** 2 / (fs / 2)
    return math.sqrt(2 * 10 ** (cn0_dbhz / 10) * n0)


def make_sats(n_sats, settings, signal="B1C", seed=20260101, cn0=45.0, sigma=25.0, max_doppler=4500.0, prns=None):
    rng = np.random.default_rng(seed)
    spc = samples_per_code(settings)
    prns = list(range(1, n_sats + 1)) if prns is None else list(prns)
    amp = cn0_to_amplitude(cn0, sigma, settings.samplingFreq) if sigma > 0 else 4.0
    sats = []
    for p in prns:
        sats.append(Struct(PRN=int(p), doppler=float(rng.uniform(-max_doppler, max_doppler)),
                           codeDelay=float(rng.uniform(0, spc)), carrPhase=float(rng.uniform(0, 2 * np.pi)),
                           amplitude=float(amp)))
    return sats


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def _sym(seed, s, period, k):
    with np.errstate(over="ignore"):
        h = _splitmix64(np.uint64(seed) ^ (np.uint64(0xD1B54A32D192ED03) * np.uint64(s * 2 + k + 1)) ^
                        (period.astype(np.int64).astype(np.uint64) * np.uint64(0x2545F4914F6CDD1D)))
    return np.where((h >> np.uint64(40)) & np.uint64(1), -1.0, 1.0)


def synth_numpy(signal, settings, sats, n, sigma=25.0, seed=20260101, first_sample=0, chunk=1 << 22, primary_codes=None, noise_seed=None):
    """Float64 host rendering of the same model as bds_synth_if (noise from numpy's generator).
    ``primary_codes(prn) -> (data, pilot)`` +-1 primary codes; default: the library's host code generator.  (bench.py's
    reference arm passes the oracle's generators so that libbdsgpu.so is never loaded in that process.)"""
    fs = settings.samplingFreq
    b1c = signal == "B1C"
    out = np.empty(n, dtype=np.int8)
    rng = np.random.default_rng(seed + 7919 if noise_seed is None else noise_seed)
    prim = []
    for st in sats:
        if primary_codes is not None:
            d_, p_ = primary_codes(st.PRN)
            prim.append((np.asarray(d_, dtype=np.float64), np.asarray(p_, dtype=np.float64)))
        elif b1c:
            prim.append((codes.gen_code(L.CODE_B1C_DATA_PRIMARY, st.PRN).astype(np.float64),
                         codes.gen_code(L.CODE_B1C_PILOT_PRIMARY, st.PRN).astype(np.float64)))
        else:
            prim.append((codes.gen_code(L.CODE_B2A_DATA, st.PRN).astype(np.float64),
                         codes.gen_code(L.CODE_B2A_PILOT, st.PRN).astype(np.float64)))
    for o in range(0, n, chunk):
        m = min(chunk, n - o)
        smp = np.arange(first_sample + o, first_sample + o + m, dtype=np.float64)
        acc = np.zeros(m)
        for s, st in enumerate(sats):
            f = settings.IF + st.doppler
            ph = np.mod(smp * (f / fs), 1.0) * (2 * np.pi) + st.carrPhase
            cs, sn = np.cos(ph), np.sin(ph)
            rate = settings.codeFreqBasis * (1.0 - st.doppler / settings.carrFreqBasis) if b1c else settings.codeFreqBasis
            tc = (smp - st.codeDelay) * (rate / fs)
            per = np.floor(tc / 10230.0)
            tin = tc - per * 10230.0
            chip = np.clip(np.floor(tin).astype(np.int64), 0, 10229)
            frac = tin - chip
            cd, cp = prim[s][0][chip], prim[s][1][chip]
            D, S = _sym(seed, s, per, 0), _sym(seed, s, per, 1)
            if b1c:
                sc1 = np.where(frac < 0.5, -1.0, 1.0)
                i6 = np.minimum((frac * 12).astype(np.int64), 11)
                sc6 = np.where(i6 & 1, 1.0, -1.0)
                acc += st.amplitude * (0.5 * D * cd * sc1 * cs -
                                       S * (math.sqrt(29 / 44) * cp * sc1 * sn + math.sqrt(4 / 44) * cp * sc6 * cs))
            else:
                acc += st.amplitude * (D * cd * sn + S * cp * cs)
        if sigma > 0:
            acc += sigma * rng.standard_normal(m)
        out[o:o + m] = np.clip(np.rint(acc), -127, 127).astype(np.int8)
    return out


def _sat_array(sats):
    arr = (L.bds_sat * len(sats))()
    for i, st in enumerate(sats):
        arr[i].PRN = int(st.PRN)
        arr[i].doppler = float(st.doppler)
        arr[i].codeDelay = float(st.codeDelay)
        arr[i].carrPhase = float(st.carrPhase)
        arr[i].amplitude = float(st.amplitude)
    return arr


def synth_device(signal, settings, sats, n, sigma=25.0, seed=20260101, first_sample=0, out_ptr=None):
    """Render on the GPU.  With ``out_ptr`` (device address) the record stays in HBM; otherwise a host
    int8 array is returned."""
    sig = L.SIG_B1C if signal == "B1C" else L.SIG_B2A
    arr = _sat_array(sats)
    args = (sig, float(settings.samplingFreq), float(settings.IF), float(settings.carrFreqBasis),
            float(settings.codeFreqBasis), arr, len(sats), float(sigma), int(seed), int(first_sample), int(n))
    if out_ptr is not None:
        L.check(L.lib().bds_synth_if(*args, C.c_void_p(out_ptr), L.LOC_DEVICE))
        return None
    out = np.empty(n, dtype=np.int8)
    L.check(L.lib().bds_synth_if(*args, L.ptr(out), L.LOC_HOST))
    return out


def channels_from_sats(sats, settings, signal="B1C", freq_error=0.0, n_channels=None):
    """The ``channel`` struct array a perfect acquisition would hand to tracking (preRun.m:61-76):
    codePhase = 1-based index of the first sample of a code period, acquiredFreq = IF + Doppler."""
    spc = samples_per_code(settings)
    nch = len(sats) if n_channels is None else n_channels
    ch = [Struct(PRN=0, acquiredFreq=0.0, codePhase=0, codeFreq=0.0, status="-") for _ in range(nch)]
    for i, st in enumerate(sats[:nch]):
        ch[i].PRN = int(st.PRN)
        ch[i].acquiredFreq = float(settings.IF + st.doppler + freq_error)
        ch[i].codePhase = int(math.ceil(st.codeDelay)) % spc + 1
        if signal == "B1C":
            ch[i].codeFreq = settings.codeFreqBasis - (ch[i].acquiredFreq - settings.IF) / settings.carrFreqBasis * settings.codeFreqBasis
        else:
            ch[i].codeFreq = float(settings.codeFreqBasis)
        ch[i].status = "T"
    return ch

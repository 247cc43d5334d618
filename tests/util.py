"""Shared helpers for the parity tests."""
from __future__ import annotations

import functools

import numpy as np

import bds_oracle as O
import c_oracle
import bds3_b200 as B
from bds3_b200 import synth

FS = 99.375e6


def settings_for(mode, **over):
    if mode in ("WB", "NB"):
        s = O.initSettings_B1C(samplingFreq=FS, pilotTRKflag=2 if mode == "WB" else 1)
    else:
        s = O.initSettings_B2a()
    s.update(over)
    return s


def product_settings(s):
    return B.Settings(dict(s))


@functools.lru_cache(maxsize=8)
def record(mode, n_sats, seconds, seed=11, sigma=25.0, max_doppler=None):
    """(settings, sats, x int8, channels) for a small synthetic record."""
    s = settings_for(mode, numberOfChannels=n_sats)
    sig = "B2a" if mode == "B2a" else "B1C"
    if max_doppler is None:
        max_doppler = 100.0 if sig == "B2a" else 4500.0
    sats = synth.make_sats(n_sats, s, sig, seed=seed, sigma=sigma, max_doppler=max_doppler)
    x = synth.synth_numpy(sig, s, sats, int(seconds * s.samplingFreq), sigma=sigma, seed=seed)
    ch = synth.channels_from_sats(sats, s, sig, freq_error=2.0)
    return s, sats, x, ch


def ochannels(ch):
    return [O.Settings(dict(c)) for c in ch]


RAW_NAMES = [f"{fam}_{iq}_{epl}" for fam in ("d", "p", "p61") for epl in ("E", "P", "L") for iq in ("I", "Q")]


def oracle_track(mode, s, x, ch, n_epochs, record_nco=True):
    """Oracle closed loop with the C correlator; also returns raw 18 sums per channel-epoch."""
    raws = []

    def corr(mode_, st_, raw, codes, rem, step, cf, rc):
        out = c_oracle.correlate_epoch(mode_, st_, raw, codes, rem, step, cf, rc)
        raws.append(np.array([out[0].get(k, 0.0) for k in RAW_NAMES]))
        return out

    tr, _ = O.tracking(mode, x, ochannels(ch), s, n_epochs=n_epochs, record_nco=record_nco, correlator=corr)
    nact = sum(1 for c in ch if c.PRN != 0)
    raw = np.array(raws).reshape(nact, -1, 18) if raws and len(raws) == nact * n_epochs else np.array(raws)
    return tr, raw


def family_scale(raw):
    """max(|I_P|,|Q_P|) of the same replica family and epoch -> shape like raw."""
    sc = np.empty_like(raw)
    for fam in range(3):
        ip = np.abs(raw[..., fam * 6 + 2])
        qp = np.abs(raw[..., fam * 6 + 3])
        m = np.maximum(ip, qp)
        m = np.where(m > 0, m, np.inf)   # family absent in this mode: sums are exactly 0 on both sides
        for k in range(6):
            sc[..., fam * 6 + k] = m
    return sc


def one_step_parity(mode, s, x, ch_c, g, n_epochs, tol_sums=1e-4):
    """Closed-loop parity that is insensitive to trajectory chaos (a single sample crossing a chip edge
    changes the following epochs at the 1e-3 level for ~1/loop-bandwidth seconds).  For every epoch of the
    device's own trajectory ``g``:
      (a) the 18 sums equal the oracle correlator evaluated at the device's NCO state (tol_sums * scale);
      (b) feeding the device's sums to the oracle's loop closure reproduces the device's next state
          (carrFreq, codeFreq, discriminators, remCodePhase, remCarrPhase) to float64 rounding.
    Returns the worst relative sum error."""
    import math
    codes = O.make_track_codes(mode, s, ch_c.PRN)
    coef = O.loop_coefficients(mode, s)
    st = O.LoopState(codeFreq=ch_c.codeFreq, carrFreq=ch_c.acquiredFreq, carrFreqBasis=ch_c.acquiredFreq,
                     pos=int(s.skipNumberOfBytes + ch_c.codePhase - 1))
    worst = 0.0
    for e in range(n_epochs):
        assert g.absoluteSample[e] == st.pos
        np.testing.assert_allclose([g.codeFreq[e], g.carrFreq[e]], [st.codeFreq, st.carrFreq], rtol=1e-13)
        np.testing.assert_allclose([g.remCodePhase[e], g.remCarrPhase[e]], [st.remCodePhase, st.remCarrPhase],
                                   rtol=0, atol=1e-9)
        # take the device's state (identical up to rounding) so the comparison stays one-step
        st.codeFreq, st.carrFreq = float(g.codeFreq[e]), float(g.carrFreq[e])
        st.remCodePhase, st.remCarrPhase = float(g.remCodePhase[e]), float(g.remCarrPhase[e])
        step = st.codeFreq / s.samplingFreq
        blk = int(math.ceil((s.codeLength - st.remCodePhase) / step))
        out, rc, rp = c_oracle.correlate_epoch(mode, s, x[st.pos: st.pos + blk], codes, st.remCodePhase, step,
                                               st.carrFreq, st.remCarrPhase)
        ref = np.array([out.get(k, 0.0) for k in RAW_NAMES])
        err = np.abs(g.raw[e] - ref) / family_scale(ref[None, :])[0]
        worst = max(worst, float(np.max(err)))
        assert np.max(err) <= tol_sums, (e, float(np.max(err)))
        sums = {k: float(g.raw[e][i]) for i, k in enumerate(RAW_NAMES) if k in out}
        st.remCodePhase, st.remCarrPhase = rc, rp
        st.pos += blk
        o = O.close_loops(mode, s, sums, st, coef, ch_c.codeFreq)
        for f in ("dllDiscr", "dllDiscrFilt", "pllDiscr", "pllDiscrFilt"):
            np.testing.assert_allclose(g[f][e], o[f], rtol=1e-9, atol=1e-13)
        for f in [k for k in o if k.startswith("Pilot_") or k in ("I_P", "Q_P", "I_E", "Q_E", "I_L", "Q_L")]:
            np.testing.assert_allclose(g[f][e], o[f], rtol=1e-12, atol=1e-6)
    return worst

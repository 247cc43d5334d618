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


@functools.lru_cache(maxsize=4)
def record_iq(mode, n_sats, seconds, seed=11, sigma=25.0):
    """The same scenario as ``record`` sampled as I/Q (settings.fileType == 2): complex x = I + 1i*Q, the quadrature
    rail being the same satellites a quarter cycle behind (Re / Im of the analytic signal) with independent noise."""
    s, sats, xr, ch = record(mode, n_sats, seconds, seed=seed, sigma=sigma)
    sig = "B2a" if mode == "B2a" else "B1C"
    quad = [B.Settings(dict(st, carrPhase=st.carrPhase - np.pi / 2)) for st in sats]
    xq = synth.synth_numpy(sig, s, quad, xr.size, sigma=sigma, seed=seed, noise_seed=seed + 104729)
    s = O.Settings(dict(s, fileType=2))
    x = xr.astype(np.float64) + 1j * xq.astype(np.float64)
    # the trackers wipe the carrier off with exp(-i theta) for B1C (WB_tracking.m:341) and exp(+i theta) for B2a
    # (tracking.m:309): a B1C tracker locks to exp(+i theta), the B2a tracker to its conjugate
    return s, sats, (np.conj(x) if mode == "B2a" else x), ch


def ochannels(ch):
    return [O.Settings(dict(c)) for c in ch]


RAW_NAMES = [f"{fam}_{iq}_{epl}" for fam in ("d", "p", "p61") for epl in ("E", "P", "L") for iq in ("I", "Q")]


def oracle_track(mode, s, x, ch, n_epochs, record_nco=True):
    """Oracle closed loop with the C correlator; also returns raw 18 sums per channel-epoch."""
    raws = []

    def corr(mode_, st_, raw, codes, rem, step, cf, rc):
        # the C restatement takes real int8 blocks; complex (fileType 2) blocks go through the numpy oracle
        fn = O.correlate_epoch if np.iscomplexobj(raw) else c_oracle.correlate_epoch
        out = fn(mode_, st_, raw, codes, rem, step, cf, rc)
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


def replay_loop_chain(mode, s, chans, planes, n_epochs):
    """Whole-trajectory check of the closed loop's bookkeeping, vectorised over channels (cheap: no correlation).

    From the device's own discriminator history (pllDiscr, dllDiscr) the reference's loop filters
    (WB_tracking.m:399-406, 422-430) are replayed in float64 and must reproduce the device's carrFreq / codeFreq /
    filtered discriminators of EVERY epoch; from (codeFreq, remCodePhase, carrFreq, remCarrPhase) of epoch e the
    reference's block bookkeeping (WB:258-261, 327, 335-337, 254) must give absoluteSample / remCodePhase /
    remCarrPhase of epoch e+1.  Returns the filter memories per epoch (for one_step_parity_sampled)."""
    tau1, tau2, pf3, pf2, pf1, _ = O.loop_coefficients(mode, s)
    nch = len(chans)
    fs, L, PDI = s.samplingFreq, float(s.codeLength), s.intTime
    basis = np.array([c.acquiredFreq for c in chans], dtype=np.float64)
    chcf = np.array([c.codeFreq for c in chans], dtype=np.float64)
    d2 = np.zeros(nch); d1 = np.zeros(nch); old_nco = np.zeros(nch); old_err = np.zeros(nch)
    mem = np.zeros((n_epochs, 4, nch))
    g = planes
    np.testing.assert_allclose(g["carrFreq"][:, 0], basis, rtol=1e-15)
    np.testing.assert_allclose(g["codeFreq"][:, 0], chcf, rtol=1e-15)
    b1c = mode != "B2a"
    for e in range(n_epochs):
        mem[e] = (d2, d1, old_nco, old_err)
        ce = g["pllDiscr"][:, e]
        d1_prev = d1
        d2 = d2 + ce * pf3
        d1 = d2 + ce * pf2 + d1
        nco = d1 + ce * pf1
        np.testing.assert_allclose(g["pllDiscrFilt"][:, e], nco, rtol=1e-12, atol=1e-13)
        # continue from the DEVICE's filter memories (recovered from its outputs), so that every epoch is checked as one
        # step and the rounding differences of 30 000 accumulations (fma on the device, two roundings here) do not add up
        nco = g["pllDiscrFilt"][:, e]
        d1_dev = nco - ce * pf1
        d2 = d1_dev - ce * pf2 - d1_prev
        d1 = d1_dev
        de = g["dllDiscr"][:, e]
        cn = old_nco + (tau2 / tau1) * (de - old_err) + de * (PDI / tau1)
        np.testing.assert_allclose(g["dllDiscrFilt"][:, e], cn, rtol=1e-12, atol=1e-13)
        cn = g["dllDiscrFilt"][:, e]
        old_nco, old_err = cn, de
        step = g["codeFreq"][:, e] / fs
        rem = g["remCodePhase"][:, e]
        blk = np.ceil((L - rem) / step)
        if e + 1 < n_epochs:
            np.testing.assert_allclose(g["carrFreq"][:, e + 1], basis + nco, rtol=1e-13)
            np.testing.assert_allclose(g["codeFreq"][:, e + 1], chcf - cn, rtol=1e-13)
            np.testing.assert_array_equal(g["absoluteSample"][:, e + 1], g["absoluteSample"][:, e] + blk)
            base = (blk - 1) * step + rem
            rem_next = (base * 2) / 2 + step - L if b1c else (base + step) - L
            np.testing.assert_allclose(g["remCodePhase"][:, e + 1], rem_next, rtol=0, atol=1e-9)
            trig = ((g["carrFreq"][:, e] * 2.0 * np.pi) * (blk / fs)) + g["remCarrPhase"][:, e]
            diff = np.abs(np.fmod(trig, 2 * np.pi) - g["remCarrPhase"][:, e + 1])
            assert np.all(np.minimum(diff, 2 * np.pi - diff) <= 1e-8), float(diff.max())
    return mem


def one_step_parity_sampled(mode, s, get_block, ch_c, g, raw_c, mem_c, epochs, tol_sums=1e-4):
    """one_step_parity for selected epochs of a long run.  ``g``: the channel's planes (dict of 1-D arrays), ``raw_c``
    [N,18] device sums, ``mem_c`` [N,4] filter memories from replay_loop_chain, ``get_block(pos, n)`` -> int8 samples.
    (a) device sums == oracle correlator at the device's NCO state (tol_sums x family scale);
    (b) oracle loop closure fed with the device's sums reproduces the device's discriminators and stored prompts."""
    import math
    codes = O.make_track_codes(mode, s, ch_c.PRN)
    coef = O.loop_coefficients(mode, s)
    worst = 0.0
    for e in epochs:
        step = float(g["codeFreq"][e]) / s.samplingFreq
        rem = float(g["remCodePhase"][e])
        blk = int(math.ceil((s.codeLength - rem) / step))
        pos = int(g["absoluteSample"][e])
        out, rc, rp = c_oracle.correlate_epoch(mode, s, get_block(pos, blk), codes, rem, step, float(g["carrFreq"][e]),
                                               float(g["remCarrPhase"][e]))
        ref = np.array([out.get(k, 0.0) for k in RAW_NAMES])
        err = np.abs(raw_c[e] - ref) / family_scale(ref[None, :])[0]
        worst = max(worst, float(np.max(err)))
        assert np.max(err) <= tol_sums, (ch_c.PRN, e, float(np.max(err)))
        st = O.LoopState(codeFreq=float(g["codeFreq"][e]), carrFreq=float(g["carrFreq"][e]),
                         carrFreqBasis=ch_c.acquiredFreq, d2CarrError=float(mem_c[e][0]), dCarrError=float(mem_c[e][1]),
                         oldCodeNco=float(mem_c[e][2]), oldCodeError=float(mem_c[e][3]))
        sums = {k: float(raw_c[e][i]) for i, k in enumerate(RAW_NAMES) if k in out}
        o = O.close_loops(mode, s, sums, st, coef, ch_c.codeFreq)
        for f in ("dllDiscr", "dllDiscrFilt", "pllDiscr", "pllDiscrFilt"):
            np.testing.assert_allclose(g[f][e], o[f], rtol=1e-9, atol=1e-13)
        for f in [k for k in o if k.startswith("Pilot_") or k in ("I_P", "Q_P", "I_E", "Q_E", "I_L", "Q_L")]:
            np.testing.assert_allclose(g[f][e], o[f], rtol=1e-12, atol=1e-6)
    return worst


def lock_report(mode, s, planes, n_epochs, seconds_tail=10.0, injected_cn0=45.0):
    """Per-channel lock over the tail of a run: DataPLD / PilotPLD (Calc_CNo_PLD.m:70-73) and the total C/N0
    (WB_tracking.m:467-477).  The stored C/N0 includes the other satellites of the record as noise."""
    per = s.intTime * s.CNoInterval
    nc = n_epochs // int(s.CNoInterval)
    k = max(1, min(nc - 1, int(round(seconds_tail / per))))   # never the first (half-scale) point
    sl = slice(nc - k, nc)
    dpld = planes["DataPLD"][:, sl].min(axis=1)
    ppld = planes["PilotPLD"][:, sl].min(axis=1)
    cno = np.median(planes["TotalCNo"][:, sl], axis=1)
    return {"data_pld_min": dpld, "pilot_pld_min": ppld, "cno_median": cno, "points": k}

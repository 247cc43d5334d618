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

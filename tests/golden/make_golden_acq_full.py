#!/usr/bin/env python
"""Generates tests/golden/acq_b1c_full_grid.npz: the float64 oracle's B1C acquisition on the FULL reference grid
(+-5 kHz in 50 Hz steps = 201 Doppler bins, 10 ms coherent, data + pilot; acquisition.m:129-307) for three PRNs of a
seeded synthetic record (two present, one absent).  The oracle needs ~2 000 transforms of 1 987 500 points for this -
minutes of CPU - so the GPU test compares against these stored results instead of running the oracle on the GPU box;
the record itself is re-rendered from its seed by the test (synth.synth_numpy is deterministic).

    python tests/golden/make_golden_acq_full.py      (CPU only; ~5 min)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]

import bds_oracle as O  # noqa: E402
from bds3_b200 import synth  # noqa: E402

SEED, PRNS, SECONDS, CN0 = 41, [7, 23, 40], 0.0305, 47.0


def scenario():
    s = O.initSettings_B1C(samplingFreq=99.375e6, acqSatelliteList=PRNS)
    sats = synth.make_sats(2, s, "B1C", seed=SEED, prns=PRNS[:2], cn0=CN0, max_doppler=4500.0)
    x = synth.synth_numpy("B1C", s, sats, int(SECONDS * s.samplingFreq), seed=SEED)
    return s, sats, x


def main():
    s, sats, x = scenario()
    acq, dbg = O.acquisition_B1C(x, s, return_debug=True)
    np.savez(os.path.join(HERE, "acq_b1c_full_grid.npz"), prns=np.array(PRNS), carrFreq=acq.carrFreq, codePhase=acq.codePhase,
             peakMetric=acq.peakMetric, bin=np.array([dbg[p]["bin"] for p in PRNS]),
             coarseCodePhase=np.array([dbg[p]["codePhase"] for p in PRNS]), peak=np.array([dbg[p]["peak"] for p in PRNS]),
             record_sha_head=np.frombuffer(x[:4096].tobytes(), dtype=np.uint8).sum(),
             doppler=np.array([st.doppler for st in sats]), codeDelay=np.array([st.codeDelay for st in sats]))
    print(acq.carrFreq[[p - 1 for p in PRNS]], acq.codePhase[[p - 1 for p in PRNS]], acq.peakMetric[[p - 1 for p in PRNS]])


if __name__ == "__main__":
    main()

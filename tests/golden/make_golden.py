#!/usr/bin/env python
"""Generates tests/golden/golden.npz — regression fixtures for the oracle and the CUDA path.

The reference (pure MATLAB) ships no golden vectors and cannot be executed in this image
(no MATLAB/Octave), so these vectors are produced by the float64 oracle restatement
(oracle/bds_oracle.py) on a small committed int8 IF record.  They pin the oracle against
regressions and give the GPU tests an input/output pair that does not depend on the synthetic
generator; they do NOT pin the oracle to MATLAB ("parity unpinned").

    python tests/golden/make_golden.py        (CPU only; needs libbdsgpu.so for the host code generator
                                                used by the synthetic record)
"""
import hashlib
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]

import bds_oracle as O  # noqa: E402
import util  # noqa: E402
from bds3_b200 import synth  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.int8).tobytes()).hexdigest()


def main():
    out = {}
    names, digs = [], []
    for comp, fn in (("b1c_data", O.b1c_data_primary), ("b1c_pilot", O.b1c_pilot_primary),
                     ("b2a_data", O.generateB2aDataCode), ("b2a_pilot", O.generateB2aPilotCode)):
        for prn in range(1, 64):
            names.append(f"{comp}:{prn}")
            digs.append(sha(fn(prn)))
    out["code_names"] = np.array(names)
    out["code_sha256"] = np.array(digs)
    out["head24_b1c_data"] = np.array([O.b1c_data_primary(p)[:24] for p in range(1, 5)], dtype=np.int8)
    out["head24_b2a_data"] = np.array([O.generateB2aDataCode(p)[:24] for p in range(1, 5)], dtype=np.int8)

    prn = 7
    out["trk_prn"] = np.int32(prn)
    # ---- B1C: one 10 ms epoch
    s = O.initSettings_B1C(samplingFreq=util.FS)
    sats = synth.make_sats(2, s, "B1C", seed=77, prns=[prn, 23])
    sats[0].codeDelay = 1234.5
    x = synth.synth_numpy("B1C", s, sats, 1_000_000, seed=77)
    out["if_b1c"] = x
    ch = synth.channels_from_sats(sats, s, "B1C", freq_error=1.5)[0]
    rem, remcarr = 0.0173, 1.25
    step = ch.codeFreq / s.samplingFreq
    blk = int(math.ceil((s.codeLength - rem) / step))
    pos = int(ch.codePhase - 1)
    out["trk_nco"] = np.array([pos, blk, rem, step, ch.acquiredFreq, remcarr])
    for mode in ("WB", "NB"):
        s.pilotTRKflag = 2 if mode == "WB" else 1
        cod = O.make_track_codes(mode, s, prn)
        o, rc, rp = O.correlate_epoch(mode, s, x[pos:pos + blk], cod, rem, step, ch.acquiredFreq, remcarr)
        out[f"trk_sums_{mode}"] = np.array([o.get(k, 0.0) for k in util.RAW_NAMES])
        out[f"trk_next_{mode}"] = np.array([rc, rp])
    # ---- B2a: one 1 ms epoch
    s2 = O.initSettings_B2a()
    sats2 = synth.make_sats(2, s2, "B2a", seed=78, prns=[prn, 23], max_doppler=100.0)
    sats2[0].codeDelay = 321.25
    xb = synth.synth_numpy("B2a", s2, sats2, 101_000, seed=78)
    out["if_b2a"] = xb
    ch2 = synth.channels_from_sats(sats2, s2, "B2a", freq_error=1.5)[0]
    step2 = ch2.codeFreq / s2.samplingFreq
    rem2 = 0.31
    blk2 = int(math.ceil((s2.codeLength - rem2) / step2))
    pos2 = int(ch2.codePhase - 1)
    out["trk_nco_b2a"] = np.array([pos2, blk2, rem2, step2, ch2.acquiredFreq, 0.4])
    cod = O.make_track_codes("B2a", s2, prn)
    o, rc, rp = O.correlate_epoch("B2a", s2, xb[pos2:pos2 + blk2], cod, rem2, step2, ch2.acquiredFreq, 0.4)
    out["trk_sums_B2a"] = np.array([o.get(k, 0.0) for k in util.RAW_NAMES])
    out["trk_next_B2a"] = np.array([rc, rp])
    np.savez_compressed(os.path.join(HERE, "golden.npz"), **out)
    print("wrote", os.path.join(HERE, "golden.npz"), os.path.getsize(os.path.join(HERE, "golden.npz")), "bytes")


if __name__ == "__main__":
    main()

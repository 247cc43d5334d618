"""Cross-check of the two independently written CPU restatements of the reference (the reference ships no
golden vectors, MATLAB cannot run here, so a second derivation is what guards the first against a misreading).

  oracle/bds_oracle.py       numpy float64, vectorised, written from the MATLAB by the round-1 builder
  oracle/c/bds_ref_model.cpp scalar per-sample C++ loops, written separately from the MATLAB text in round 2

Bit-exact on every primary code (63 PRNs x 4 components); closed-loop trackers (WB_tracking.m, NB_tracking.m, B2a
tracking.m) agree to 1e-12 over >= 100 epochs including the C/N0 and lock-detector outputs.  When /root/reference is
present the per-PRN constant tables are also parsed out of the MATLAB files themselves and compared with the oracle's."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import bds_oracle as O
import util

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE = os.path.join(os.path.dirname(HERE), "oracle")
REF = "/root/reference/BDS3_B1C_B2a"


@pytest.fixture(scope="module")
def ref():
    subprocess.check_call(["make", "-s", "-C", ORACLE])
    lib = C.CDLL(os.path.join(ORACLE, "_build", "librefmodel.so"))
    lib.ref_track.restype = C.c_int
    lib.ref_track.argtypes = [C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                              C.c_double, C.c_int, C.c_void_p, C.c_void_p]
    return lib


def _matlab_matrix(path, first_line, last_line):
    """integers of a MATLAB matrix literal spanning the given (1-based, inclusive) lines"""
    txt = "".join(open(path, encoding="latin-1").readlines()[first_line - 1:last_line])
    txt = re.sub(r"%.*", "", txt)
    return [int(v) for v in re.findall(r"-?\d+", txt.split("=", 1)[1])]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
def test_constant_tables_equal_the_matlab_literals():
    wp = np.array(_matlab_matrix(f"{REF}/BDS-3_B1C/include/generateDataBOC11.m", 43, 58)).reshape(63, 2)
    assert list(wp[:, 0]) == list(O.B1C_DATA_W) and list(wp[:, 1]) == list(O.B1C_DATA_P)
    for f in ("generatePilotBOC11.m", "generatePilotBOC61.m"):
        wp = np.array(_matlab_matrix(f"{REF}/BDS-3_B1C/include/{f}", 44, 59)).reshape(63, 2)
        assert list(wp[:, 0]) == list(O.B1C_PILOT_W) and list(wp[:, 1]) == list(O.B1C_PILOT_P)
    for f, tab in (("generateB2aDataCode.m", O.B2A_DATA_G2), ("generateB2aPilotCode.m", O.B2A_PILOT_G2)):
        m = np.array(_matlab_matrix(f"{REF}/BDS-3_B2a/include/{f}", 39, 101)).reshape(63, 13)
        assert np.array_equal(m, np.array(_g2_bits(tab)))


def _g2_bits(tab):
    """the oracle keeps the register-2 initial states in whatever form it likes; as 63 x 13 logic values"""
    a = np.asarray(tab)
    if a.ndim == 2:
        return a.tolist()
    return [[(int(v) >> i) & 1 for i in range(13)] for v in a]   # bit i = stage i + 1


def test_codes_bit_exact(ref):
    out = np.zeros(10230, dtype=np.int8)
    for prn in range(1, 64):
        ref.ref_weil_primary(int(O.B1C_DATA_W[prn - 1]), int(O.B1C_DATA_P[prn - 1]), out.ctypes.data_as(C.c_void_p))
        assert np.array_equal(out, O.b1c_data_primary(prn).astype(np.int8)), prn
        ref.ref_weil_primary(int(O.B1C_PILOT_W[prn - 1]), int(O.B1C_PILOT_P[prn - 1]), out.ctypes.data_as(C.c_void_p))
        assert np.array_equal(out, O.b1c_pilot_primary(prn).astype(np.int8)), prn
        for pilot, tab, gen in ((0, O.B2A_DATA_G2, O.generateB2aDataCode), (1, O.B2A_PILOT_G2, O.generateB2aPilotCode)):
            ini = np.array(_g2_bits(tab)[prn - 1], dtype=np.int32)
            ref.ref_b2a_code(pilot, ini.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
            assert np.array_equal(out, np.asarray(gen(prn)).astype(np.int8)), (prn, pilot)


FIELDS = ["absoluteSample", "codeFreq", "carrFreq", "I_P", "I_E", "I_L", "Q_E", "Q_P", "Q_L", "Pilot_I_P", "Pilot_I_E",
          "Pilot_I_L", "Pilot_Q_E", "Pilot_Q_P", "Pilot_Q_L", "dllDiscr", "dllDiscrFilt", "pllDiscr", "pllDiscrFilt",
          "remCodePhase", "remCarrPhase"]


def _run_ref(ref, mode, s, x, ch, n_epochs):
    tau1, tau2, pf3, pf2, pf1, factor = O.loop_coefficients(mode, s)
    S = np.array([s.samplingFreq, s.codeFreqBasis, s.codeLength, s.dllCorrelatorSpacing, s.intTime, s.pilotTRKflag,
                  s.CNoInterval, tau1, tau2, pf3, pf2, pf1, factor, 1, s.skipNumberOfBytes], dtype=np.float64)
    if mode == "B2a":
        d, p = O.generateB2aDataCode(ch.PRN), O.generateB2aPilotCode(ch.PRN)
    else:
        d, p = O.b1c_data_primary(ch.PRN), O.b1c_pilot_primary(ch.PRN)
    d, p = np.ascontiguousarray(d, dtype=np.int8), np.ascontiguousarray(p, dtype=np.int8)
    out = np.zeros((22, n_epochs))
    cno = np.zeros((5, max(1, n_epochs // int(s.CNoInterval))))
    pv = lambda a: a.ctypes.data_as(C.c_void_p)
    done = ref.ref_track({"WB": 1, "NB": 2, "B2a": 3}[mode], pv(x), x.size, pv(S), pv(d), pv(p), float(ch.codeFreq),
                         float(ch.acquiredFreq), float(ch.codePhase), n_epochs, pv(out), pv(cno))
    return done, out, cno


@pytest.mark.parametrize("mode,epochs,interval", [("WB", 100, 25), ("NB", 100, 25), ("B2a", 120, 40)])
def test_closed_loop_trackers_agree_to_1e_12(ref, mode, epochs, interval):
    spc_s = 0.001 if mode == "B2a" else 0.01
    s, sats, x, ch = util.record(mode, 1, spc_s * (epochs + 2.5), seed=23)
    s = s.copy()
    s.CNoInterval = interval
    tr, _ = O.tracking(mode, x, util.ochannels(ch), s, n_epochs=epochs)
    done, out, cno = _run_ref(ref, mode, s, x, ch[0], epochs)
    assert done == epochs and tr[0].status == "T"
    t = tr[0]
    np.testing.assert_array_equal(out[0], t.absoluteSample)
    scale = np.maximum(np.abs(t.I_P), np.abs(t.Q_P))
    for i, f in enumerate(FIELDS):
        if f not in t or i == 0:
            continue
        a, b = out[i], np.asarray(t[f])
        if f in ("codeFreq", "carrFreq"):
            np.testing.assert_allclose(a, b, rtol=1e-13, err_msg=f)
        elif f.startswith(("I_", "Q_")):
            assert np.max(np.abs(a - b) / scale) <= 1e-12, (f, float(np.max(np.abs(a - b) / scale)))
        elif f.startswith("Pilot_"):
            ps = np.maximum(np.abs(t.Pilot_I_P), np.abs(t.Pilot_Q_P))
            assert np.max(np.abs(a - b) / ps) <= 1e-12, (f, float(np.max(np.abs(a - b) / ps)))
        else:   # discriminators, filtered discriminators, phase remainders
            np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-11, err_msg=f)
    nc = epochs // interval
    names = ["DataCNo", "DataPLD", "PilotCNo", "PilotPLD", "B2a_CNo" if mode == "B2a" else "B1C_CNo"]
    for i, f in enumerate(names):
        np.testing.assert_allclose(cno[i, :nc], np.asarray(t[f])[:nc], rtol=1e-9, atol=1e-9, err_msg=f)


def test_short_read_stops_both_models_at_the_same_epoch(ref):
    s, sats, x, ch = util.record("NB", 1, 0.055, seed=5)
    tr, _ = O.tracking("NB", x, util.ochannels(ch), s, n_epochs=8)
    done, out, _ = _run_ref(ref, "NB", s, x, ch[0], 8)
    n_or = int(np.count_nonzero(np.isfinite(tr[0].codeFreq)))
    assert done == n_or and tr[0].status == "-"
    np.testing.assert_array_equal(out[0, :done + 1], tr[0].absoluteSample[:done + 1])

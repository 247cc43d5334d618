"""CPU tests (no GPU): the oracle against structural known answers, its independent C restatement,
the committed golden fixtures, and the library's host-side integer code (bds_gen_code /
bds_make_code_table run on the host and must equal the oracle bit for bit).

The reference ships no golden vectors: these checks are what pins the oracle.
"""
import math
import os

import numpy as np
import pytest
from hypothesis import given, settings as hsettings, strategies as hst

import bds_oracle as O
import c_oracle
import util
import bds3_b200 as B
from bds3_b200 import _lib as L, codes, loopcoef, synth

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "golden.npz")


# ---- a7: B1C codes ---------------------------------------------------------------------------------
def test_legendre_sequence_by_euler_criterion():
    leg = O.legendre_sequence()
    assert leg[0] == 0 and leg.size == 10243
    want = np.array([0] + [1 if pow(k, (10243 - 1) // 2, 10243) == 1 else 0 for k in range(1, 10243)])
    np.testing.assert_array_equal(leg, want)
    assert leg.sum() == (10243 - 1) // 2


@pytest.mark.parametrize("prn", [1, 19, 20, 63])
def test_weil_code_structure(prn):
    d, p = O.b1c_data_primary(prn), O.b1c_pilot_primary(prn)
    for c in (d, p):
        assert c.size == 10230 and set(np.unique(c)) == {-1.0, 1.0}
        assert abs(c.sum()) < 300            # Weil codes are nearly balanced
    # definition, written independently: chip n = L(k) xor L(k+w), k = (n+p-1) mod N  (generateDataBOC11.m:76-79)
    leg = O.legendre_sequence()
    w, pp = O.B1C_DATA_W[prn - 1], O.B1C_DATA_P[prn - 1]
    for n in (0, 1, 5000, 10229):
        k = (n + pp - 1) % 10243
        assert d[n] == 1 - 2 * (leg[k] ^ leg[(k + w) % 10243])
    assert abs(np.dot(d, p)) < 500           # data / pilot are different codes
    # periodic autocorrelation: peak = length, sidelobes small
    f = np.fft.fft(d)
    ac = np.fft.ifft(f * np.conj(f)).real
    assert round(ac[0]) == 10230 and np.max(np.abs(ac[1:])) < 500


def test_boc_expansion_pattern():
    prim = O.b1c_pilot_primary(7)
    b11 = O.generatePilotBOC11(None, 7)
    b61 = O.generatePilotBOC61(None, 7)
    assert b11.size == 20460 and b61.size == 122760
    np.testing.assert_array_equal(b11[0::2], -prim)          # chip c -> [-c, +c]
    np.testing.assert_array_equal(b11[1::2], prim)
    sub = b61.reshape(10230, 12)
    for ii in range(12):                                      # chip c -> (-1)^ii c, ii = 1..12
        np.testing.assert_array_equal(sub[:, ii], prim * (-1.0) ** (ii + 1))


# ---- a8: B2a codes ---------------------------------------------------------------------------------
def _b2a_lfsr_bits(g2_init, taps1, taps2):
    """Independent logic-level (0/1, xor) implementation of the two 13-stage registers."""
    r1 = [1] * 13
    r2 = [(g2_init >> k) & 1 for k in range(13)]
    out = []
    for ind in range(1, 10231):
        out.append(r1[12] ^ r2[12])
        f1 = 0
        for t in taps1:
            f1 ^= r1[t - 1]
        f2 = 0
        for t in taps2:
            f2 ^= r2[t - 1]
        r1 = [f1] + r1[:12]
        r2 = [f2] + r2[:12]
        if ind == 8190:
            r1 = [1] * 13
    return np.array(out)


@pytest.mark.parametrize("prn", [1, 30, 61, 63])
def test_b2a_codes_against_logic_level_lfsr(prn):
    d = O.generateB2aDataCode(prn)
    p = O.generateB2aPilotCode(prn)
    np.testing.assert_array_equal(d, 1 - 2 * _b2a_lfsr_bits(O.B2A_DATA_G2[prn - 1], (1, 5, 11, 13), (3, 5, 9, 11, 12, 13)))
    np.testing.assert_array_equal(p, 1 - 2 * _b2a_lfsr_bits(O.B2A_PILOT_G2[prn - 1], (3, 6, 7, 13), (1, 5, 7, 8, 12, 13)))
    # register 1 has period 8191 and is reset after chip 8190: chips 8190.. replay register 1 from its start
    assert d.size == 10230 and abs(d.sum()) < 400 and abs(np.dot(d, p)) < 600


def test_b2a_reg2_tables_rows_61_63_differ():
    assert O.B2A_DATA_G2[:60] == O.B2A_PILOT_G2[:60]
    assert all(a != b for a, b in zip(O.B2A_DATA_G2[60:], O.B2A_PILOT_G2[60:]))


# ---- library host code == oracle, bit exact (all 63 PRNs) --------------------------------------------
def test_library_codegen_bit_exact_all_prns():
    for prn in range(1, 64):
        np.testing.assert_array_equal(codes.gen_code(L.CODE_B1C_DATA_PRIMARY, prn), O.b1c_data_primary(prn))
        np.testing.assert_array_equal(codes.gen_code(L.CODE_B1C_PILOT_PRIMARY, prn), O.b1c_pilot_primary(prn))
        np.testing.assert_array_equal(codes.gen_code(L.CODE_B2A_DATA, prn), O.generateB2aDataCode(prn))
        np.testing.assert_array_equal(codes.gen_code(L.CODE_B2A_PILOT, prn), O.generateB2aPilotCode(prn))
    for prn in (1, 33, 63):
        np.testing.assert_array_equal(codes.generateDataBOC11(None, prn), O.generateDataBOC11(None, prn))
        np.testing.assert_array_equal(codes.generatePilotBOC11(None, prn), O.generatePilotBOC11(None, prn))
        np.testing.assert_array_equal(codes.generatePilotBOC61(None, prn), O.generatePilotBOC61(None, prn))


def test_library_codegen_rejects_bad_arguments():
    out = np.empty(10230, dtype=np.int8)
    assert L.lib().bds_gen_code(L.CODE_B2A_DATA, 64, L.ptr(out), out.size) == -1
    assert L.lib().bds_gen_code(L.CODE_B2A_DATA, 1, L.ptr(out), 100) == -1
    assert b"" != L.lib().bds_last_error()


@pytest.mark.parametrize("fs", [99.375e6, 53e6])
def test_library_code_tables_bit_exact(fs):
    s1 = O.initSettings_B1C(samplingFreq=fs)
    p1 = B.Settings(dict(s1))
    for prn in (1, 20):
        t = O.makeDataTable(s1, prn)
        assert t.size == O.samples_per_code(s1)
        np.testing.assert_array_equal(codes.makeDataTable(p1, prn), t)
        np.testing.assert_array_equal(codes.makePilotTable(p1, prn), O.makePilotTable(s1, prn))
    s2 = O.initSettings_B2a(samplingFreq=fs)
    p2 = B.Settings(dict(s2))
    for prn in (4, 62):
        np.testing.assert_array_equal(codes.makeB2aDataTable(prn, p2), O.makeB2aDataTable(prn, s2))
        np.testing.assert_array_equal(codes.makeB2aPilotTable(prn, p2), O.makeB2aPilotTable(prn, s2))


def test_code_table_quirks():
    """makeDataTable.m:49-63: index = ceil(ts*k/tc), first forced to 1, last forced to 20460."""
    s = O.initSettings_B1C(samplingFreq=99.375e6)
    t = O.makeDataTable(s, 1)
    c = O.generateDataBOC11(s, 1)
    assert t.size == 993750 and t[0] == c[0] and t[-1] == c[-1]
    # exact-integer boundary every 6625 samples (fc/fs = 682/6625): sample k=6625 -> ceil(1364.0) = 1364
    assert t[6624] == c[1363]


# ---- a13: loop constants -----------------------------------------------------------------------------
def test_loop_constants_match_survey_values():
    s = O.initSettings_B1C(samplingFreq=99.375e6)
    tau1, tau2 = O.calcLoopCoef(s.dllNoiseBandwidth, s.dllDampingRatio, 1.0)
    assert abs(tau1 - 0.279388) < 1e-6 and abs(tau2 - 0.74) < 1e-12
    np.testing.assert_allclose(O.calcLoopCoefCarr(s), (0.298598, 4.1472, 28.8), rtol=2e-6)
    s2 = O.initSettings_B2a()
    tau1, tau2 = O.calcLoopCoef(s2.dllNoiseBandwidth, s2.dllDampingRatio, 1.0)
    assert abs(tau1 - 0.069847) < 1e-6 and abs(tau2 - 0.37) < 1e-12
    np.testing.assert_allclose(O.calcLoopCoefCarr(s2), (0.013824, 1.152, 48), rtol=1e-9)
    assert abs(O.CalcWeighingFactor(s) - 0.163502) < 2e-6


def test_product_loop_constants_equal_oracle():
    for s in (O.initSettings_B1C(samplingFreq=99.375e6), O.initSettings_B2a()):
        p = B.Settings(dict(s))
        assert loopcoef.calcLoopCoef(p.dllNoiseBandwidth, p.dllDampingRatio, 1.0) == \
            O.calcLoopCoef(s.dllNoiseBandwidth, s.dllDampingRatio, 1.0)
        assert loopcoef.calcLoopCoefCarr(p) == O.calcLoopCoefCarr(s)
    s = O.initSettings_B1C(samplingFreq=99.375e6)
    # two different quadratures (scipy quad vs composite Gauss-Legendre) of the same integrals
    assert abs(loopcoef.CalcWeighingFactor(B.Settings(dict(s))) - O.CalcWeighingFactor(s)) < 1e-9


# ---- MATLAB colon ------------------------------------------------------------------------------------
@given(rem=hst.floats(0.0, 0.999), dcode=hst.floats(-6.0, 6.0))
@hsettings(max_examples=60, deadline=None)
def test_blksize_and_remcodephase_invariants(rem, dcode):
    """0 <= remCodePhase_next < step, and the colon vector has exactly blksize elements whose last
    element is the stop expression (SURVEY §8c item 4, quirk ii)."""
    fs, L_ = 99.375e6, 10230
    step = (1.023e6 + dcode) / fs
    rem = rem * step
    blk = int(math.ceil((L_ - rem) / step))
    stop = ((blk - 1) * step + rem) * 2
    t = O.colon(rem * 2, step * 2, stop, blk)
    assert t.size == blk and t[-1] == stop and t[0] == rem * 2
    nxt = t[-1] / 2 + step - L_
    assert -1e-9 <= nxt < step + 1e-9
    assert np.all(np.diff(t) > 0)


# ---- a9-a11: C restatement == numpy oracle -----------------------------------------------------------
@pytest.mark.parametrize("mode,seconds", [("WB", 0.025), ("NB", 0.025), ("B2a", 0.004)])
def test_c_oracle_equals_numpy_oracle(mode, seconds):
    s, sats, x, ch = util.record(mode, 2, seconds)
    c = ch[0]
    cod = O.make_track_codes(mode, s, c.PRN)
    pos = int(c.codePhase - 1)
    for rem, carr_err in ((0.0, 0.0), (0.0123, 3.0)):
        step = c.codeFreq / s.samplingFreq
        blk = int(math.ceil((s.codeLength - rem) / step))
        raw = x[pos: pos + blk]
        a, ra, pa = O.correlate_epoch(mode, s, raw, cod, rem, step, c.acquiredFreq + carr_err, 0.3)
        b, rb, pb = c_oracle.correlate_epoch(mode, s, raw, cod, rem, step, c.acquiredFreq + carr_err, 0.3)
        assert set(a) == set(b) and len(a) == (18 if mode == "WB" else 12)
        scale = max(abs(v) for v in a.values())
        for k in a:
            assert abs(a[k] - b[k]) <= 1e-9 * scale, (k, a[k], b[k])
        assert ra == rb and abs(pa - pb) < 1e-9


def test_sign_conventions_b1c_vs_b2a():
    """B1C mixes with exp(-i theta), I = real; B2a with exp(+i theta), I = imag."""
    s, sats, x, ch = util.record("NB", 1, 0.012, sigma=0.0)
    c = ch[0]
    step = c.codeFreq / s.samplingFreq
    blk = int(math.ceil(s.codeLength / step))
    pos = int(c.codePhase - 1)
    out, _, _ = O.correlate_epoch("NB", s, x[pos:pos + blk], O.make_track_codes("NB", s, c.PRN), 0.0, step,
                                  c.acquiredFreq - 2.0, 0.0)
    # noise-free, frequency-exact: data power lands in (d_I, d_Q), pilot BOC(1,1) 90 degrees away
    d = complex(out["d_I_P"], out["d_Q_P"])
    p = complex(out["p_I_P"], out["p_Q_P"])
    assert abs(d) > 1e5 and abs(p) > 1e5
    ang = np.angle(p / d)
    assert abs(abs(ang) - np.pi / 2) < 0.05


# ---- acquisition oracle recovers what was injected ---------------------------------------------------
def test_b2a_acquisition_oracle_recovers_injected():
    s = O.initSettings_B2a(acqSatelliteList=[4, 9])
    sats = synth.make_sats(1, s, "B2a", seed=3, prns=[4], cn0=47.0)
    x = synth.synth_numpy("B2a", s, sats, 17 * 99375, seed=3)
    acq, dbg = O.acquisition_B2a(x, s, return_debug=True)
    assert acq.carrFreq[3] != 0 and acq.carrFreq[8] == 0
    assert acq.peakMetric[3] > s.acqThreshold > acq.peakMetric[8]
    assert abs(acq.carrFreq[3] - (s.IF + sats[0].doppler)) <= 25
    spc = O.samples_per_code(s)
    d = (acq.codePhase[3] - 1 - sats[0].codeDelay) % spc
    assert min(d, spc - d) <= 1.5


def test_b1c_acquisition_oracle_recovers_injected():
    s = O.initSettings_B1C(samplingFreq=util.FS, acqSearchBand=100, acqSatelliteList=[1, 2])
    sats = synth.make_sats(1, s, "B1C", seed=5, max_doppler=40.0, cn0=47.0)
    x = synth.synth_numpy("B1C", s, sats, int(0.0305 * util.FS), seed=5)
    acq = O.acquisition_B1C(x, s)
    assert acq.carrFreq[0] != 0 and acq.carrFreq[1] == 0
    assert abs(acq.carrFreq[0] - (s.IF + sats[0].doppler)) <= 25
    spc = O.samples_per_code(s)
    d = (acq.codePhase[0] - 1 - sats[0].codeDelay) % spc
    assert min(d, spc - d) <= 1.5


# ---- tracking oracle: loops lock on the synthetic signal ---------------------------------------------
def test_tracking_oracle_locks_and_decodes_symbols():
    s, sats, x, ch = util.record("WB", 1, 0.42)
    s = s.copy()
    s.CNoInterval = 10
    tr, _ = util.oracle_track("WB", s, x, ch, 40, record_nco=False)
    r = tr[0]
    assert r.status == "T"
    assert np.all(np.abs(r.dllDiscr[10:]) < 0.2)
    assert np.abs(r.carrFreq[-1] - (s.IF + sats[0].doppler)) < 1.0      # 2 Hz initial error pulled in
    assert r.PilotPLD[-1] > 0.9 and r.DataPLD[-1] > 0.5
    assert 40.0 < r.B1C_CNo[-1] < 50.0                                    # injected 45 dB-Hz
    assert np.all(np.hypot(r.I_P[20:], r.Q_P[20:]) > 1e5)


def test_cno_first_point_is_half_scale():
    """WB_tracking.m:467-481: stored value = 0.5*new + 0.5*previous interval's (0 for the first)."""
    s, sats, x, ch = util.record("NB", 1, 0.13)
    s = s.copy()
    s.CNoInterval = 5
    tr, _ = util.oracle_track("NB", s, x, ch, 10, record_nco=False)
    r = tr[0]
    cno, _ = O.Calc_CNo_PLD(r, s, 5, "NB")
    assert r.DataCNo[0] == pytest.approx(0.5 * cno[0])


# ---- golden regression fixtures (generated by tests/golden/make_golden.py from the oracle) -----------
def test_golden_fixtures():
    g = np.load(GOLDEN)
    import hashlib

    def h(a):
        return hashlib.sha256(np.ascontiguousarray(a, dtype=np.int8).tobytes()).hexdigest()

    names = list(g["code_names"])
    digests = list(g["code_sha256"])
    for name, dig in zip(names, digests):
        comp, prn = name.split(":")
        fn = {"b1c_data": O.b1c_data_primary, "b1c_pilot": O.b1c_pilot_primary,
              "b2a_data": O.generateB2aDataCode, "b2a_pilot": O.generateB2aPilotCode}[comp]
        assert h(fn(int(prn))) == dig, name
    # first 24 chips of PRN 1..4 (the quantity the reference's commented-out self-test prints,
    # generatePilotBOC61.m:98-106)
    for comp, fn in (("b1c_data", O.b1c_data_primary), ("b2a_data", O.generateB2aDataCode)):
        for prn in range(1, 5):
            np.testing.assert_array_equal(g[f"head24_{comp}"][prn - 1], fn(prn)[:24])
    # one tracking epoch per mode on the committed 12 ms int8 record
    x = g["if_b1c"]
    s = O.initSettings_B1C(samplingFreq=util.FS)
    for mode in ("WB", "NB"):
        s.pilotTRKflag = 2 if mode == "WB" else 1
        cod = O.make_track_codes(mode, s, int(g["trk_prn"]))
        nco = g["trk_nco"]
        out, rc, rp = O.correlate_epoch(mode, s, x[int(nco[0]): int(nco[0]) + int(nco[1])], cod, nco[2], nco[3],
                                        nco[4], nco[5])
        got = np.array([out.get(k, 0.0) for k in util.RAW_NAMES])
        np.testing.assert_allclose(got, g[f"trk_sums_{mode}"], rtol=0, atol=1e-6)
        np.testing.assert_allclose([rc, rp], g[f"trk_next_{mode}"], rtol=0, atol=1e-12)
    xb = g["if_b2a"]
    s2 = O.initSettings_B2a()
    cod = O.make_track_codes("B2a", s2, int(g["trk_prn"]))
    nco = g["trk_nco_b2a"]
    out, rc, rp = O.correlate_epoch("B2a", s2, xb[int(nco[0]): int(nco[0]) + int(nco[1])], cod, nco[2], nco[3],
                                    nco[4], nco[5])
    got = np.array([out.get(k, 0.0) for k in util.RAW_NAMES])
    np.testing.assert_allclose(got, g["trk_sums_B2a"], rtol=0, atol=1e-6)


def test_secondary_code_and_frame_sync_restatement():
    """generate2ndCode.m / BCNAV1decoding.m:66-91 / BCNAV2decoding.m:69-97 as restated by the oracle: the secondary
    codes are 1800 +-1 chips, distinct per PRN, with the Weil structure; a prompt sequence that carries the pattern
    (either polarity) at a known epoch is found there and nowhere else."""
    codes = [O.generate2ndCode(p) for p in (1, 2, 19, 63)]
    for c in codes:
        assert c.size == 1800 and set(np.unique(c)) == {-1.0, 1.0} and abs(c.sum()) < 120
    assert all(abs(np.dot(codes[0], c)) < 200 for c in codes[1:])
    # Weil structure: chip = L(k) xor L(k + w), against a Legendre sequence built from the squares
    N = 3607
    leg = np.zeros(N, dtype=int)
    leg[(np.arange(1, N) ** 2) % N] = 1
    w, p = O.B1C_2ND_WP[18]
    k = (np.arange(1800) + p - 1) % N
    assert np.array_equal(O.generate2ndCode(19), 1 - 2 * (leg[k] ^ leg[(k + w) % N]))
    rng = np.random.default_rng(4)
    tr = O.Settings(PRN=19, Pilot_I_P=rng.standard_normal(4000) * 3.0)
    tr.Pilot_I_P[700:2500] = -O.generate2ndCode(19) * np.abs(tr.Pilot_I_P[700:2500] + 5.0)     # inverted polarity
    s = O.initSettings_B1C(pilotTRKflag=2)
    X, idx = O.frame_sync_B1C(tr, s)
    assert list(idx) == [701] and X[700] == -1800.0 and X.size == 4000
    ip = rng.standard_normal(3000)
    pre = np.kron([-1, -1, -1, 1, 1, 1, -1, 1, 1, -1, 1, 1, -1, -1, 1, -1, -1, -1, -1, 1, -1, 1, 1, 1], [1, 1, 1, -1, 1])
    ip[1234:1354] = pre * 2.0
    X, idx = O.frame_sync_B2a(ip)
    assert 1235 in idx and X[1234] == 120.0


def test_lock_loss_extension_is_off_by_default_and_drops_a_dead_channel():
    """settings.lockLossPLD (extension): without it the trajectory is the reference's; with it a channel whose signal
    disappears is dropped at the end of the C/N0 interval in which its lock detector has stayed low for
    lockLossIntervals intervals, and the next channel is still tracked."""
    FS = util.FS
    s = O.initSettings_B1C(samplingFreq=FS, pilotTRKflag=1, numberOfChannels=2, CNoInterval=5)
    sats = synth.make_sats(2, s, "B1C", seed=11, cn0=50.0)
    n1 = int(0.3 * FS)
    x = np.concatenate([synth.synth_numpy("B1C", s, sats, n1, seed=11),
                        synth.synth_numpy("B1C", s, [sats[1]], int(0.35 * FS), seed=12, first_sample=n1)])   # PRN 1 vanishes
    ch = [O.Settings(dict(c)) for c in synth.channels_from_sats(sats, s, "B1C", freq_error=0.0)]
    N = 60
    base, _ = O.tracking("NB", x, ch, s, n_epochs=N, correlator=c_oracle.correlate_epoch)
    s2 = s.copy()
    s2.lockLossPLD, s2.lockLossIntervals = 0.9, 3
    got, _ = O.tracking("NB", x, ch, s2, n_epochs=N, correlator=c_oracle.correlate_epoch)
    assert base[0].status == "T" and "lockLostEpoch" not in base[0]
    lost = got[0].lockLostEpoch
    assert got[0].status == "-" and 30 < lost < N and lost % 5 == 0
    # the rule, from the channel's own lock detector values
    low = 0
    for c, v in enumerate(base[0].PilotPLD):
        low = low + 1 if v < 0.9 else 0
        if low >= 3:
            assert lost == (c + 1) * 5
            break
    np.testing.assert_array_equal(got[0].I_P[:lost], base[0].I_P[:lost])
    assert np.all(got[0].I_P[lost:] == 0) and np.all(np.isinf(got[0].carrFreq[lost:]))
    assert got[1].status == "T" and np.array_equal(got[1].I_P, base[1].I_P)       # the other channel is untouched


# ---- resampling pre-conditioner (acquisition.m:56-123): third-party arithmetic (fir1 / filtfilt), restated ---------
def test_fir1_and_filtfilt_restatements_against_scipy():
    """fir1 == scipy.signal.firwin (an independent implementation of the same published window method), filtfilt ==
    scipy.signal.filtfilt with MATLAB's padding length, and the zero-phase result on the original samples equals the
    correlation of the odd-extended record with b (*) flip(b) - the form the device kernel evaluates."""
    from scipy import signal
    wp = [(14.58e6 - 4.5e6) * 2 / 99.375e6 - 0.002, (14.58e6 + 4.5e6) * 2 / 99.375e6 + 0.002]     # acquisition.m:64-67
    b = O.fir1_bandpass(700, wp)
    np.testing.assert_allclose(b, signal.firwin(701, wp, window="hamming", pass_zero=False, scale=True), rtol=0, atol=1e-14)
    assert abs(abs(np.sum(b * np.exp(-1j * np.pi * np.mean(wp) * np.arange(701)))) - 1) < 1e-12
    x = np.rint(np.random.default_rng(1).normal(0, 25, 9000))
    y = O.filtfilt_fir(b, x)
    np.testing.assert_allclose(y, signal.filtfilt(b, [1.0], x, padlen=2100), rtol=0, atol=1e-11)
    g = np.correlate(b, b, mode="full")
    xt = np.concatenate([2 * x[0] - x[2100:0:-1], x, 2 * x[-1] - x[-2:-2102:-1]])
    np.testing.assert_allclose(y, np.correlate(xt, g, mode="same")[2100:-2100], rtol=0, atol=1e-11)
    with pytest.raises(ValueError):
        O.filtfilt_fir(b, x[:2100])


def test_resampling_rates_and_index_vector():
    """acquisition.m:79-122 at the BASELINE sampling rate: B1C 19.62 MHz (IF above fs/2: the mirrored branch of
    :327-333), B2a 48.06 MHz"""
    s = O.initSettings_B1C(samplingFreq=99.375e6, resamplingflag=1)
    x = np.arange(5000, dtype=np.float64)
    y, st, oldFreq, oldIF = O.resample_for_acquisition(x, s, 9e6)
    assert st.samplingFreq == 19620000.0 and st.IF == s.IF and st.IF >= st.samplingFreq / 2 and (oldFreq, oldIF) == (99.375e6, s.IF)
    assert y.size == int(np.floor(4999 / 99.375e6 * 19.62e6))
    s2 = O.initSettings_B2a(resamplingflag=1)
    y2, st2, _, _ = O.resample_for_acquisition(x, s2, s2.codeFreqBasis * 2 + 0.5e6)
    assert st2.samplingFreq == 48060000.0 and st2.IF == s2.IF < st2.samplingFreq / 2
    same, st3, o3, _ = O.resample_for_acquisition(x, O.initSettings_B1C(samplingFreq=99.375e6), 9e6)
    assert same is x and o3 is None                                   # resamplingflag = 0: untouched

"""GPU parity of acquisition against the float64 oracle (run with -m gpu).
Bar: identical codePhase and carrFreq grid point, peakMetric within 1e-4 relative."""
import numpy as np
import pytest

import bds_oracle as O
import util
import bds3_b200 as B
from bds3_b200 import synth

pytestmark = pytest.mark.gpu


def _b1c_record(n_sats, seconds, seed, **over):
    s = O.initSettings_B1C(samplingFreq=util.FS, **over)
    sats = synth.make_sats(n_sats, s, "B1C", seed=seed, max_doppler=min(4500.0, s.acqSearchBand - 60), cn0=47.0)
    x = synth.synth_numpy("B1C", s, sats, int(seconds * s.samplingFreq), seed=seed)
    return s, sats, x


def test_b1c_acquisition_parity_reduced_band():
    s, sats, x = _b1c_record(2, 0.0305, 5, acqSearchBand=200, acqSatelliteList=[1, 2, 3])
    want, wd = O.acquisition_B1C(x, s, return_debug=True)
    got, gd = B.b1c.acquisition(x, B.Settings(dict(s)), return_debug=True)
    for prn in (1, 2, 3):
        assert gd[prn - 1, 0] == wd[prn]["bin"]
        assert gd[prn - 1, 1] == wd[prn]["codePhase"]
    np.testing.assert_allclose(got.peakMetric, want.peakMetric, rtol=1e-4)
    np.testing.assert_array_equal(got.codePhase, want.codePhase)
    np.testing.assert_array_equal(got.carrFreq, want.carrFreq)
    assert want.carrFreq[0] != 0 and want.carrFreq[1] != 0 and want.carrFreq[2] == 0
    # injected parameters are recovered
    for st in sats:
        assert abs(got.carrFreq[st.PRN - 1] - (s.IF + st.doppler)) <= 25
        spc = O.samples_per_code(s)
        d = (got.codePhase[st.PRN - 1] - 1 - st.codeDelay) % spc
        assert min(d, spc - d) <= 1.5


def test_b1c_acquisition_data_only():
    s, sats, x = _b1c_record(1, 0.0305, 6, acqSearchBand=100, acqSatelliteList=[1], pilotACQflag=0)
    want = O.acquisition_B1C(x, s)
    got = B.b1c.acquisition(x, B.Settings(dict(s)))
    np.testing.assert_allclose(got.peakMetric, want.peakMetric, rtol=1e-4)
    np.testing.assert_array_equal(got.codePhase, want.codePhase)
    np.testing.assert_array_equal(got.carrFreq, want.carrFreq)


def test_b2a_acquisition_parity_full_grid_config():
    """BASELINE config 1: B2a single-PRN acquisition, 17 ms record, reference defaults."""
    s = O.initSettings_B2a(acqSatelliteList=[4, 9])
    sats = synth.make_sats(1, s, "B2a", seed=3, prns=[4], cn0=47.0)
    x = synth.synth_numpy("B2a", s, sats, 17 * 99375, seed=3)
    want, wd = O.acquisition_B2a(x, s, return_debug=True)
    got, gd = B.b2a.acquisition(x, B.Settings(dict(s)), return_debug=True)
    for prn in (4, 9):
        assert gd[prn - 1, 0] == wd[prn]["bin"] and gd[prn - 1, 1] == wd[prn]["codePhase"]
    np.testing.assert_allclose(got.peakMetric, want.peakMetric, rtol=1e-4)
    np.testing.assert_array_equal(got.codePhase, want.codePhase)
    np.testing.assert_array_equal(got.carrFreq, want.carrFreq)
    assert want.carrFreq[3] != 0 and want.carrFreq[8] == 0
    assert abs(got.carrFreq[3] - (s.IF + sats[0].doppler)) <= 25


def test_acquisition_then_tracking_pipeline():
    """acquisition -> preRun -> WB_tracking through the reference-named call surface."""
    s, sats, x = _b1c_record(2, 0.23, 8, acqSearchBand=200, acqSatelliteList=[1, 2], numberOfChannels=2)
    ps = B.Settings(dict(s))
    acq = B.b1c.acquisition(x, ps)
    ch = B.b1c.preRun(acq, ps)
    assert sorted(c.PRN for c in ch) == [1, 2]
    tr, _ = B.b1c.WB_tracking(x, ch, ps, n_epochs=20)
    for r in tr:
        assert r.status == "T"
        # code is aligned and the carrier is within the loop's range: the prompt correlator sees the signal
        # (data component amplitude * N/2 ~ 2.8e5 at 47 dB-Hz); the reference PLL needs > 0.2 s to pull in
        # from the <= 12.5 Hz fine-search error, so lock quality is not asserted here
        assert np.all(np.hypot(r.I_P, r.Q_P) > 1.2e5)
    want = O.acquisition_B1C(x, s)
    och = O.preRun(want, s, "B1C")
    for a, b in zip(ch, och):
        assert (a.PRN, a.codePhase, a.acquiredFreq, a.codeFreq) == (b.PRN, b.codePhase, b.acquiredFreq, b.codeFreq)


def _iq(signal, s, sats, n, seed):
    """I/Q sampling of the scenario (settings.fileType == 2, postProcessing.m:96-99).  Both acquisitions mix with
    exp(+1i*f*t) (acquisition.m:198), so a complex record is found at +f when it rotates as exp(-i theta): the
    quadrature rail is a quarter cycle AHEAD (for real records the sign does not matter)."""
    quad = [B.Settings(dict(st, carrPhase=st.carrPhase - np.pi / 2)) for st in sats]
    xr = synth.synth_numpy(signal, s, sats, n, seed=seed)
    xq = synth.synth_numpy(signal, s, quad, n, seed=seed, noise_seed=seed + 104729)
    return xr.astype(np.float64) - 1j * xq.astype(np.float64)


def test_b1c_acquisition_parity_iq_record():
    s = O.initSettings_B1C(samplingFreq=util.FS, acqSearchBand=150, acqSatelliteList=[1, 2], fileType=2)
    sats = synth.make_sats(1, s, "B1C", seed=5, max_doppler=90.0, cn0=47.0)
    x = _iq("B1C", s, sats, int(0.0305 * s.samplingFreq), 5)
    want, wd = O.acquisition_B1C(x, s, return_debug=True)
    got, gd = B.b1c.acquisition(x, B.Settings(dict(s)), return_debug=True)
    for prn in (1, 2):
        assert gd[prn - 1, 0] == wd[prn]["bin"] and gd[prn - 1, 1] == wd[prn]["codePhase"]
    np.testing.assert_allclose(got.peakMetric, want.peakMetric, rtol=1e-4)
    np.testing.assert_array_equal(got.codePhase, want.codePhase)
    np.testing.assert_array_equal(got.carrFreq, want.carrFreq)
    assert want.carrFreq[0] != 0 and want.carrFreq[1] == 0
    assert abs(got.carrFreq[0] - (s.IF + sats[0].doppler)) <= 25


def test_b2a_acquisition_parity_iq_record():
    s = O.initSettings_B2a(acqSatelliteList=[4, 9], fileType=2)
    sats = synth.make_sats(1, s, "B2a", seed=3, prns=[4], cn0=47.0)
    x = _iq("B2a", s, sats, 17 * 99375, 3)
    want, wd = O.acquisition_B2a(x, s, return_debug=True)
    got, gd = B.b2a.acquisition(x, B.Settings(dict(s)), return_debug=True)
    for prn in (4, 9):
        assert gd[prn - 1, 0] == wd[prn]["bin"] and gd[prn - 1, 1] == wd[prn]["codePhase"]
    np.testing.assert_allclose(got.peakMetric, want.peakMetric, rtol=1e-4)
    np.testing.assert_array_equal(got.codePhase, want.codePhase)
    np.testing.assert_array_equal(got.carrFreq, want.carrFreq)
    assert want.carrFreq[3] != 0 and want.carrFreq[8] == 0


# ---- resampling pre-conditioner (acquisition.m:56-123 / B2a :56-124, results mapped back :321-338) -----------------
def test_b1c_acquisition_with_resampling_preconditioner():
    s, sats, x = _b1c_record(1, 0.0305, 5, acqSearchBand=150, acqSatelliteList=[1, 2], resamplingflag=1)
    assert s.samplingFreq > s.resamplingThreshold
    want, wd = O.acquisition_B1C(x, s, return_debug=True)
    got, gd = B.b1c.acquisition(x, B.Settings(dict(s)), return_debug=True)
    for prn in (1, 2):
        assert gd[prn - 1, 0] == wd[prn]["bin"] and gd[prn - 1, 1] == wd[prn]["codePhase"]     # at the resampled rate
    np.testing.assert_allclose(got.peakMetric, want.peakMetric, rtol=1e-4)
    np.testing.assert_array_equal(got.codePhase, want.codePhase)
    np.testing.assert_array_equal(got.carrFreq, want.carrFreq)
    assert want.carrFreq[0] != 0 and want.carrFreq[1] == 0
    spc = O.samples_per_code(s)
    d = (got.codePhase[0] - 1 - sats[0].codeDelay) % spc
    assert min(d, spc - d) <= 6            # one resampled sample = 5.07 original ones


def test_b2a_acquisition_with_resampling_preconditioner():
    s = O.initSettings_B2a(acqSatelliteList=[4, 9], resamplingflag=1)
    sats = synth.make_sats(1, s, "B2a", seed=3, prns=[4], cn0=47.0)
    x = synth.synth_numpy("B2a", s, sats, 17 * 99375, seed=3)
    want, wd = O.acquisition_B2a(x, s, return_debug=True)
    got, gd = B.b2a.acquisition(x, B.Settings(dict(s)), return_debug=True)
    for prn in (4, 9):
        assert gd[prn - 1, 0] == wd[prn]["bin"] and gd[prn - 1, 1] == wd[prn]["codePhase"]
    np.testing.assert_allclose(got.peakMetric, want.peakMetric, rtol=1e-4)
    np.testing.assert_array_equal(got.codePhase, want.codePhase)
    np.testing.assert_array_equal(got.carrFreq, want.carrFreq)
    assert want.carrFreq[3] != 0 and want.carrFreq[8] == 0
    assert abs(got.carrFreq[3] - (s.IF + sats[0].doppler)) <= 25


def test_b1c_acquisition_full_reference_grid_against_stored_oracle_results():
    """The FULL reference grid (+-5 kHz in 50 Hz steps = 201 bins, 10 ms coherent, data + pilot, 2^22-point transforms;
    acquisition.m:129-307) for three PRNs.  The float64 oracle needs minutes for this, so its results are stored
    (tests/golden/acq_b1c_full_grid.npz, written by tests/golden/make_golden_acq_full.py); the record is re-rendered
    from its seed here."""
    import os
    import sys
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, here)
    import make_golden_acq_full as G
    gold = np.load(os.path.join(here, "acq_b1c_full_grid.npz"))
    s, sats, x = G.scenario()
    assert int(np.frombuffer(x[:4096].tobytes(), dtype=np.uint8).sum()) == int(gold["record_sha_head"])   # same record
    got, gd = B.b1c.acquisition(x, B.Settings(dict(s)), return_debug=True)
    for i, prn in enumerate(gold["prns"]):
        assert gd[prn - 1, 0] == gold["bin"][i] and gd[prn - 1, 1] == gold["coarseCodePhase"][i]
        np.testing.assert_allclose(gd[prn - 1, 2], gold["peak"][i], rtol=1e-4)
    np.testing.assert_allclose(got.peakMetric, gold["peakMetric"], rtol=1e-4)
    np.testing.assert_array_equal(got.codePhase, gold["codePhase"])
    np.testing.assert_array_equal(got.carrFreq, gold["carrFreq"])
    assert got.carrFreq[6] != 0 and got.carrFreq[22] != 0 and got.carrFreq[39] == 0


@pytest.mark.parametrize("signal", ["B1C", "B2a"])
def test_specialised_inverse_passes_equal_the_generic_kernels(signal):
    """The compile-time specialised inverse passes (AUTO at the shipped transform shapes) against the generic kernels
    (cfg.tune bit 0) and the other row tiling (bit 1): same Doppler bin and code phase for every PRN, detected or not,
    peak sizes and normalisers equal to float rounding of two differently ordered transforms."""
    if signal == "B1C":
        s, sats, x = _b1c_record(2, 0.0305, 11, acqSearchBand=250, acqSatelliteList=[1, 2, 5, 9])
        run = B.b1c.acquisition
    else:
        s = O.initSettings_B2a(acqSatelliteList=[4, 9, 11])
        sats = synth.make_sats(2, s, "B2a", seed=12, prns=[4, 11], cn0=47.0)
        x = synth.synth_numpy("B2a", s, sats, 17 * 99375, seed=12)
        run = B.b2a.acquisition
    res = {}
    for tune in (0, 1, 2):
        d = dict(s)
        d["_tune"] = tune
        res[tune] = run(x, B.Settings(d), return_debug=True)
    ref, rd = res[1]
    assert np.count_nonzero(ref.carrFreq) == 2
    for tune in (0, 2):
        got, gd = res[tune]
        np.testing.assert_array_equal(gd[:, :2], rd[:, :2])            # bin, code phase
        np.testing.assert_allclose(gd[:, 2:], rd[:, 2:], rtol=2e-5)    # peak, normaliser
        np.testing.assert_array_equal(got.codePhase, ref.codePhase)
        np.testing.assert_array_equal(got.carrFreq, ref.carrFreq)
        np.testing.assert_allclose(got.peakMetric, ref.peakMetric, rtol=2e-5)


def test_b1c_acquisition_at_the_reference_shipped_53_mhz():
    """B1C/initSettings.m:57 ships samplingFreq = 53 MHz: N = 1 060 000, P = 2^21 = 2^9 x 2^12 - the third shape the inverse
    passes are specialised for.  Same bin, code phase and fine frequency as the oracle, metric within 1e-4; and equal to the
    generic kernels (cfg.tune bit 0)."""
    s = O.initSettings_B1C(acqSearchBand=200, acqSatelliteList=[1, 2, 3])
    assert s.samplingFreq == 53e6
    sats = synth.make_sats(2, s, "B1C", seed=7, max_doppler=140.0, cn0=47.0)
    x = synth.synth_numpy("B1C", s, sats, int(0.0305 * s.samplingFreq), seed=7)
    want, wd = O.acquisition_B1C(x, s, return_debug=True)
    got, gd = B.b1c.acquisition(x, B.Settings(dict(s)), return_debug=True)
    for prn in (1, 2, 3):
        assert gd[prn - 1, 0] == wd[prn]["bin"]
        assert gd[prn - 1, 1] == wd[prn]["codePhase"]
    np.testing.assert_allclose(got.peakMetric, want.peakMetric, rtol=1e-4)
    np.testing.assert_array_equal(got.codePhase, want.codePhase)
    np.testing.assert_array_equal(got.carrFreq, want.carrFreq)
    assert np.count_nonzero(want.carrFreq) == 2
    gen, dd = B.b1c.acquisition(x, B.Settings(dict(s, _tune=1)), return_debug=True)
    np.testing.assert_array_equal(dd[:, :2], gd[:, :2])
    np.testing.assert_allclose(dd[:, 2:], gd[:, 2:], rtol=2e-5)
    np.testing.assert_array_equal(gen.carrFreq, got.carrFreq)

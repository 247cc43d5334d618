"""GPU parity of the per-channel chip-synchronous B2a kernel (csrc/bds_track_b2a.cuh) against the float64 oracle.

KERNEL_AUTO selects it for B2a at fs = 99.375 MHz, d = 0.5 (first run on hardware in round 2: all green).  The arithmetic
of its generated body is covered on the CPU by tests/test_fast_body_emulation.py.  Tolerances as in
tests/test_gpu_tracking.py."""
import ctypes as C
import os

import numpy as np
import pytest

import util
from bds3_b200 import _lib as L, _track

pytestmark = [pytest.mark.gpu]


def _open_loop(s, x, prns, nco, reserved=0):
    cfg = _track.make_cfg("B2a", util.product_settings(s), L.KERNEL_AUTO)
    cfg.reserved = reserved
    nch, ne = nco.shape[0], nco.shape[1]
    sums = np.zeros((nch, ne, 18))
    prn = np.asarray(prns, dtype=np.int32)
    nco = np.ascontiguousarray(nco, dtype=np.float64)
    L.check(L.lib().bds_track_correlate_open_loop(L.TRK_B2A, C.byref(cfg), L.ptr(x), x.size, L.LOC_HOST, L.ptr(prn),
                                                  nch, ne, L.ptr(nco), L.ptr(sums)))
    return sums


@pytest.mark.parametrize("wide_guard", [0, 1])
def test_open_loop_parity(wide_guard):
    s, sats, x, ch = util.record("B2a", 2, 0.012)
    tr, raw = util.oracle_track("B2a", s, x, ch, 8)
    nco = np.stack([t.nco for t in tr])
    got = _open_loop(s, x, [c.PRN for c in ch], nco, wide_guard)
    err = np.abs(got - raw) / util.family_scale(raw)
    assert np.nanmax(err[np.isfinite(err)]) <= 1e-4, np.nanmax(err[np.isfinite(err)])
    assert np.all(got[..., 12:] == 0.0)
    fast_units, exact_units, general_slices, _ = _track.counters(None)
    assert general_slices == 0 and fast_units + exact_units == 2 * 8 * 1023
    if wide_guard:
        assert 0.05 < exact_units / (2 * 8 * 1023) < 0.6
    else:
        # the first epoch of a B2a channel starts at remCodePhase = 0 and exactly the nominal code rate (preRun gives no
        # code-Doppler aiding), which ties 18 unit phases to a threshold exactly: those go through the exact path
        assert exact_units <= 2 * 20 + 8


def test_closed_loop_matches_oracle_and_general_kernel():
    s, sats, x, ch = util.record("B2a", 2, 0.03)
    epochs = 25
    tr, raw = util.oracle_track("B2a", s, x, ch, epochs)
    ps = util.product_settings(s)
    got, _ = _track.run_tracking("B2a", x, ch, ps, n_epochs=epochs, raw=True)
    fast_units, exact_units, general_slices, _ = _track.run_tracking.last_counters
    assert general_slices == 0 and fast_units + exact_units == 2 * epochs * 1023
    for c in range(len(ch)):
        g, o = got[c], tr[c]
        assert g.status == "T" and g.PRN == o.PRN
        np.testing.assert_array_equal(g.absoluteSample, o.absoluteSample)
        sc = util.family_scale(raw[c])
        err = np.abs(g.raw - raw[c]) / sc
        err = err[np.isfinite(err)]
        # Trajectory against trajectory is not the north-star comparison (one_step_parity below is: <= 1e-4 at the device's
        # own NCO state).  The two closed loops differ by fp rounding in remCodePhase, so now and then ONE sample falls on
        # the other side of a chip edge: that moves a 1 ms sum by at most 2 * 127 LSB (|x| <= 127, |carrier| <= 1), and the
        # 2 Hz / 18 Hz loops pass the step on to the following epochs with gain < 1.  Bound = two such samples, derived
        # from the record, not calibrated; 99 % of all values must meet 1e-4 outright.
        absdiff = np.abs(g.raw - raw[c])[np.isfinite(np.abs(g.raw - raw[c]) / sc)]
        assert np.mean(err <= 1e-4) >= 0.99, float(np.mean(err <= 1e-4))
        assert np.max(absdiff) <= 2 * 2 * 127.0, float(np.max(absdiff))
        for f in ("carrFreq", "codeFreq"):
            np.testing.assert_allclose(g[f], o[f], rtol=1e-9)
        for f in ("remCodePhase", "remCarrPhase", "dllDiscr", "pllDiscr"):
            np.testing.assert_allclose(g[f], o[f], rtol=0, atol=2e-4)
        assert util.one_step_parity("B2a", s, x, ch[c], g, epochs) <= 1e-4


def test_short_read_and_resume_in_windows():
    """a record that ends inside an epoch stops the channel like tracking.m:246-251; state survives between launches"""
    s, sats, x, ch = util.record("B2a", 2, 0.0215)
    ps = util.product_settings(s)
    full, _ = _track.run_tracking("B2a", x, ch, ps, n_epochs=40, raw=True)
    ref, _ = _track.run_tracking("B2a", x, ch, ps, n_epochs=40, kernel=L.KERNEL_GENERAL, raw=True)
    for c in range(len(ch)):
        assert full[c].status == ref[c].status
        np.testing.assert_array_equal(full[c].absoluteSample, ref[c].absoluteSample)
        np.testing.assert_allclose(full[c].carrFreq, ref[c].carrFreq, rtol=1e-9)
    sess = _track.TrackSession("B2a", ps, ch, x)
    sess.run_async(7)
    sess.sync()
    sess.run_async(33)
    pl = sess.fetch(40)
    one = _track.TrackSession("B2a", ps, ch, x)
    one.run_async(40)
    np.testing.assert_array_equal(pl["I_P"], one.fetch(40)["I_P"])

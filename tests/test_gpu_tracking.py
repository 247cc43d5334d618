"""GPU parity of the tracking correlator against the float64 oracle (run with -m gpu).

Tolerance (BASELINE.md §3 / SURVEY §7.6): |x_gpu - x_ref| <= 1e-4 * max(|I_P|,|Q_P|) of the
same replica family and epoch, for the open-loop (teacher-forced) correlator.  Closed loop:
the GPU and oracle trajectories differ at the 1e-7 level (fp32 accumulation, libm vs CUDA
atan), which occasionally moves a sample across a chip edge (one-sample change ~ 2|x| in one
sum), so closed-loop comparisons use 1e-3 * scale for every value and 1e-4 for >= 99 %.
"""
import ctypes as C

import numpy as np
import pytest

import util
from bds3_b200 import _lib as L, _track

pytestmark = pytest.mark.gpu
MODES = {"WB": L.TRK_B1C_WB, "NB": L.TRK_B1C_NB, "B2a": L.TRK_B2A}


def open_loop(mode, s, x, prns, nco, kernel):
    cfg = _track.make_cfg(mode, util.product_settings(s), kernel)
    nch, ne = nco.shape[0], nco.shape[1]
    sums = np.zeros((nch, ne, 18))
    prn = np.asarray(prns, dtype=np.int32)
    nco = np.ascontiguousarray(nco, dtype=np.float64)
    L.check(L.lib().bds_track_correlate_open_loop(MODES[mode], C.byref(cfg), L.ptr(x), x.size, L.LOC_HOST, L.ptr(prn),
                                                  nch, ne, L.ptr(nco), L.ptr(sums)))
    return sums


@pytest.mark.parametrize("mode,seconds,epochs", [("WB", 0.06, 3), ("NB", 0.06, 3), ("B2a", 0.012, 8)])
def test_open_loop_parity_general(mode, seconds, epochs):
    s, sats, x, ch = util.record(mode, 2, seconds)
    tr, raw = util.oracle_track(mode, s, x, ch, epochs)
    nco = np.stack([t.nco for t in tr])
    got = open_loop(mode, s, x, [c.PRN for c in ch], nco, L.KERNEL_GENERAL)
    err = np.abs(got - raw) / util.family_scale(raw)
    assert np.nanmax(err[np.isfinite(err)]) <= 1e-4, np.nanmax(err[np.isfinite(err)])


@pytest.mark.parametrize("wide_guard", [0, 1])
def test_open_loop_parity_fast(wide_guard):
    """Chip-synchronous kernel vs oracle, strict tolerance; wide_guard=1 sends ~25 % of the chips through
    the exact per-sample path (cfg.reserved test hook), exercising the seam between the two paths."""
    s, sats, x, ch = util.record("WB", 2, 0.06)
    tr, raw = util.oracle_track("WB", s, x, ch, 3)
    nco = np.stack([t.nco for t in tr])
    cfg_s = util.product_settings(s)
    cfg = _track.make_cfg("WB", cfg_s, L.KERNEL_FAST)
    cfg.reserved = wide_guard
    sums = np.zeros((2, 3, 18))
    prn = np.asarray([c.PRN for c in ch], dtype=np.int32)
    nco = np.ascontiguousarray(nco)
    L.check(L.lib().bds_track_correlate_open_loop(L.TRK_B1C_WB, C.byref(cfg), L.ptr(x), x.size, L.LOC_HOST, L.ptr(prn),
                                                  2, 3, L.ptr(nco), L.ptr(sums)))
    err = np.abs(sums - raw) / util.family_scale(raw)
    assert np.max(err) <= 1e-4, np.max(err)
    fast_chips, exact_chips, general_slices, _ = _track.counters(None)
    assert general_slices == 0 and fast_chips + exact_chips == 2 * 3 * 10230
    if wide_guard:
        assert 0.1 < exact_chips / (2 * 3 * 10230) < 0.6
    else:
        assert exact_chips <= 2 * 3 + 2      # only chip 0 of a first epoch (rem = 0) and rare near-edge chips


def test_fast_kernel_rejects_unsupported_config():
    s, sats, x, ch = util.record("B2a", 2, 0.012)
    s = s.copy()
    s.dllCorrelatorSpacing = 0.4  # the B2a chip-synchronous body is generated for the reference's 0.5 chip
    with pytest.raises(L.BdsError):
        _track.run_tracking("B2a", x, ch, util.product_settings(s), n_epochs=2, kernel=L.KERNEL_FAST)
    s, sats, x, ch = util.record("NB", 2, 0.06)
    s = s.copy()
    s.pilotTRKflag = 0            # data-only narrow band has no fast path
    with pytest.raises(L.BdsError):
        _track.run_tracking("NB", x, ch, util.product_settings(s), n_epochs=2, kernel=L.KERNEL_FAST)


def test_narrow_band_fast_kernel():
    """NB_tracking.m on the chip-synchronous kernel: open-loop sums (BOC(6,1) family exactly zero) and the
    closed loop one step at a time against the oracle's NB loop closure."""
    s, sats, x, ch = util.record("NB", 2, 0.23)
    tr, raw = util.oracle_track("NB", s, x, ch, 3)
    nco = np.stack([t.nco for t in tr])
    got = open_loop("NB", s, x, [c.PRN for c in ch], nco, L.KERNEL_FAST)
    assert np.max(np.abs(got - raw) / util.family_scale(raw)) <= 1e-4
    assert np.all(got[..., 12:] == 0.0)
    ps = util.product_settings(s)
    res, _ = _track.run_tracking("NB", x, ch, ps, n_epochs=20, kernel=L.KERNEL_FAST, raw=True)
    fast_chips, exact_chips, general_slices, _ = _track.run_tracking.last_counters
    assert general_slices == 0 and fast_chips + exact_chips == 2 * 20 * 10230
    for c in range(2):
        assert res[c].status == "T" and "Pilot_I_E" not in res[c]
        assert util.one_step_parity("NB", s, x, ch[c], res[c], 20) <= 1e-4
    auto, _ = _track.run_tracking("NB", x, ch, ps, n_epochs=20, raw=True)       # AUTO now picks the fast kernel
    assert _track.run_tracking.last_counters[2] == 0
    np.testing.assert_array_equal(auto[0].raw, res[0].raw)


@pytest.mark.parametrize("kernel", ["general", "fast"])
def test_closed_loop_one_step_parity_wb(kernel):
    """Every epoch of the device's closed-loop trajectory: sums == oracle correlator at the device's NCO
    state (1e-4), and the oracle loop closure fed with the device's sums reproduces the device's next
    state to rounding (util.one_step_parity).  Trajectory-vs-trajectory comparison is only meaningful
    until the first chip-edge flip, see the module docstring."""
    s, sats, x, ch = util.record("WB", 2, 0.23)
    ps = util.product_settings(s)
    kern = L.KERNEL_GENERAL if kernel == "general" else L.KERNEL_FAST
    got, _ = _track.run_tracking("WB", x, ch, ps, n_epochs=20, kernel=kern, raw=True)
    fast_chips, exact_chips, general_slices, _ = _track.run_tracking.last_counters
    if kernel == "fast":
        assert general_slices == 0 and fast_chips + exact_chips == 2 * 20 * 10230 and exact_chips <= 64
    else:
        assert fast_chips == 0 and general_slices > 0
    for c in range(2):
        assert got[c].status == "T"
        worst = util.one_step_parity("WB", s, x, ch[c], got[c], 20)
        assert worst <= 1e-4


def test_closed_loop_fast_tracks_oracle_trajectory():
    """Trajectory-vs-trajectory: identical first epoch, then the two closed loops stay together up to the
    chaos introduced by single samples crossing chip edges (see util.one_step_parity for the strict test)."""
    s, sats, x, ch = util.record("WB", 2, 0.13)
    ps = util.product_settings(s)
    tr, raw = util.oracle_track("WB", s, x, ch, 10)
    fast, _ = _track.run_tracking("WB", x, ch, ps, n_epochs=10, kernel=L.KERNEL_FAST, raw=True)
    for c in range(2):
        err0 = np.abs(fast[c].raw[0] - raw[c][0]) / util.family_scale(raw[c][0][None, :])[0]
        assert np.max(err0) <= 1e-4
        np.testing.assert_allclose(fast[c].carrFreq, tr[c].carrFreq, rtol=0, atol=0.05)         # Hz
        np.testing.assert_allclose(fast[c].remCodePhase, tr[c].remCodePhase, rtol=0, atol=1e-3)  # chips
        np.testing.assert_allclose(fast[c].absoluteSample, tr[c].absoluteSample, rtol=0, atol=1)


def test_open_loop_first_epoch_t0_sample():
    """remCodePhase = 0: the t = 0 sample takes the previous period's last chip (SURVEY quirk i)."""
    s, sats, x, ch = util.record("WB", 2, 0.06)
    tr, raw = util.oracle_track("WB", s, x, ch, 1)
    nco = np.stack([t.nco for t in tr])
    assert np.all(nco[:, 0, 2] == 0.0)
    got = open_loop("WB", s, x, [c.PRN for c in ch], nco, L.KERNEL_GENERAL)
    assert np.max(np.abs(got - raw) / util.family_scale(raw)) <= 1e-4


@pytest.mark.parametrize("mode,seconds,epochs", [("WB", 0.13, 10), ("NB", 0.13, 10), ("B2a", 0.03, 25)])
def test_closed_loop_general(mode, seconds, epochs):
    s, sats, x, ch = util.record(mode, 2, seconds)
    tr, raw = util.oracle_track(mode, s, x, ch, epochs)
    ps = util.product_settings(s)
    got, _ = _track.run_tracking(mode, x, ch, ps, n_epochs=epochs, kernel=L.KERNEL_GENERAL, raw=True)
    for c in range(len(ch)):
        g, o = got[c], tr[c]
        assert g.status == "T" and g.PRN == o.PRN
        np.testing.assert_array_equal(g.absoluteSample, o.absoluteSample)
        sc = util.family_scale(raw[c])
        err = np.abs(g.raw - raw[c]) / sc
        assert np.max(err) <= 1e-3, np.max(err)
        assert np.mean(err <= 1e-4) >= 0.99
        for f in ("carrFreq", "codeFreq"):
            np.testing.assert_allclose(g[f], o[f], rtol=1e-9)
        for f in ("remCodePhase", "remCarrPhase", "dllDiscr", "pllDiscr"):
            np.testing.assert_allclose(g[f], o[f], rtol=0, atol=2e-4)
        scale = np.maximum(np.abs(o.I_P), np.abs(o.Q_P))
        for f in ("I_P", "Q_P", "I_E", "Q_E", "I_L", "Q_L"):
            assert np.max(np.abs(g[f] - o[f]) / scale) <= 1e-3
        if "Pilot_I_P" in o:
            ps_ = np.maximum(np.abs(o.Pilot_I_P), np.abs(o.Pilot_Q_P))
            for f in [k for k in o.keys() if k.startswith("Pilot_")]:
                assert np.max(np.abs(g[f] - o[f]) / ps_) <= 1e-3


def test_field_order_and_initial_values():
    s, sats, x, ch = util.record("WB", 2, 0.06)
    ps = util.product_settings(s)
    ps.numberOfChannels = 3
    ch3 = list(ch) + [type(ch[0])(PRN=0, acquiredFreq=0.0, codePhase=0, codeFreq=0.0, status="-")]
    got, _ = _track.run_tracking("WB", x, ch3, ps, n_epochs=3)
    want = ["status", "absoluteSample", "codeFreq", "carrFreq", "I_P", "I_E", "I_L", "Q_E", "Q_P", "Q_L",
            "Pilot_I_P", "Pilot_I_E", "Pilot_I_L", "Pilot_Q_E", "Pilot_Q_P", "Pilot_Q_L", "dllDiscr", "dllDiscrFilt",
            "pllDiscr", "pllDiscrFilt", "remCodePhase", "remCarrPhase", "DataCNo", "DataPLD", "PilotCNo", "PilotPLD",
            "B1C_CNo", "PRN"]
    assert list(got[0].keys())[:len(want)] == want
    unused = got[2]
    assert unused.status == "-" and "PRN" not in unused
    assert np.all(np.isinf(unused.carrFreq)) and np.all(unused.I_P == 0)


def test_short_read_stops_like_reference():
    """Record shorter than requested: the first starving channel keeps its partial data with status '-',
    later channels keep the template (WB_tracking.m:279-283)."""
    s, sats, x, ch = util.record("WB", 2, 0.06)
    ps = util.product_settings(s)
    got, _ = _track.run_tracking("WB", x, ch, ps, n_epochs=20)
    tr, _ = util.oracle_track("WB", s, x, ch, 20, record_nco=False)
    assert got[0].status == "-" and tr[0].status == "-"
    done = got[0].epochsDone
    assert 0 < done < 20
    assert np.all(np.isinf(got[0].carrFreq[done:])) and np.all(np.isfinite(got[0].carrFreq[:done]))
    np.testing.assert_array_equal(np.isfinite(got[0].carrFreq), np.isfinite(tr[0].carrFreq))
    assert got[1].status == "-" and np.all(np.isinf(got[1].carrFreq)) and "PRN" not in got[1]
    assert np.all(np.isinf(tr[1].carrFreq))


@pytest.mark.parametrize("mode,kernel", [("WB", "general"), ("WB", "fast"), ("NB", "general"), ("NB", "fast")])
def test_cno_and_lock_detector(mode, kernel):
    s, sats, x, ch = util.record(mode, 2, 0.13)
    s = s.copy()
    s.CNoInterval = 5
    tr, _ = util.oracle_track(mode, s, x, ch, 10)
    kern = L.KERNEL_GENERAL if kernel == "general" else L.KERNEL_FAST
    got, _ = _track.run_tracking(mode, x, ch, util.product_settings(s), n_epochs=10, kernel=kern)
    for c in range(2):
        for f in ("DataCNo", "PilotCNo", "B1C_CNo"):
            np.testing.assert_allclose(got[c][f], tr[c][f], atol=0.02)
        for f in ("DataPLD", "PilotPLD"):
            np.testing.assert_allclose(got[c][f], tr[c][f], atol=2e-3)


@pytest.mark.parametrize("kernel", ["general", "fast"])
def test_streamed_host_record_equals_resident_record(kernel, tmp_path):
    """bds_track_run_streamed (chunked H2D under the kernel; channels stop at the end of the arrived data and
    are resumed by the next launch) gives bit-identical results to tracking the fully resident record, for
    a host array, a file path (bds_track_open_file) and odd chunk sizes."""
    s, sats, x, ch = util.record("WB", 2, 0.13)
    ps = util.product_settings(s)
    kern = L.KERNEL_GENERAL if kernel == "general" else L.KERNEL_FAST
    N = 9
    with _track.TrackSession("WB", ps, ch, kernel=kern) as ses:        # resident: one copy, then one launch
        ses.feed(x)
        ses.run_async(N)
        want = ses.fetch(N, raw=True)
    outs = []
    for chunk in (1 << 20, 3 * 4096 + 4096 * 777):
        with _track.TrackSession("WB", ps, ch, kernel=kern) as ses:
            ses.run_streamed(x.ctypes.data, x.size, N, chunk_bytes=chunk)
            outs.append(ses.fetch(N, raw=True))
    path = tmp_path / "if.bin"
    x.tofile(path)
    with _track.TrackSession("WB", ps, ch, source=str(path), kernel=kern) as ses:
        ses.run_async(N)
        outs.append(ses.fetch(N, raw=True))
    got_api, _ = _track.run_tracking("WB", x, ch, ps, n_epochs=N, kernel=kern, raw=True)
    for o in outs:
        assert list(o["epochsDone"]) == [N, N]
        for k in want:
            np.testing.assert_array_equal(o[k], want[k], err_msg=k)
    np.testing.assert_array_equal(got_api[0].raw, want["raw"][0])


def test_golden_fixture_open_loop():
    """The committed golden record / sums (tests/golden, generated from the oracle) through both kernels."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden.npz"))
    x = g["if_b1c"]
    for mode, kernels in (("WB", (L.KERNEL_GENERAL, L.KERNEL_FAST)), ("NB", (L.KERNEL_GENERAL,))):
        s = util.settings_for(mode)
        want = g[f"trk_sums_{mode}"]
        for kern in kernels:
            got = open_loop(mode, s, x, [int(g["trk_prn"])], g["trk_nco"].reshape(1, 1, 6), kern)[0, 0]
            err = np.abs(got - want) / util.family_scale(want[None, :])[0]
            assert np.max(err) <= 1e-4, (mode, kern, float(np.max(err)))
    s = util.settings_for("B2a")
    got = open_loop("B2a", s, g["if_b2a"], [int(g["trk_prn"])], g["trk_nco_b2a"].reshape(1, 1, 6), L.KERNEL_GENERAL)[0, 0]
    want = g["trk_sums_B2a"]
    assert np.max(np.abs(got - want) / util.family_scale(want[None, :])[0]) <= 1e-4


def test_closed_loop_is_deterministic_under_load():
    """24 channels x 58 epochs repeated 60 times on one session (resident and streamed alternately): every channel
    completes every epoch and every run gives bit-identical outputs (the slice sums are exact integer additions, so
    the order in which slices arrive cannot matter).  Regression test for a shared-memory overlay bug that corrupted
    a closer warp's channel state about once per 10^7 loop closures."""
    import torch
    from bds3_b200 import synth
    st = util.product_settings(util.settings_for("WB", numberOfChannels=24))
    sats = synth.make_sats(24, st, "B1C", seed=5)
    ch = synth.channels_from_sats(sats, st, "B1C", freq_error=2.0)
    n = int(0.6 * util.FS)
    x = synth.synth_device("B1C", st, sats, n, seed=5)
    ne = 58
    ref = None
    with _track.TrackSession("WB", st, ch) as ses:
        for it in range(60):
            ses.reset()
            if it % 2:
                ses.run_streamed(x.ctypes.data, x.size, ne, chunk_bytes=8 << 20)
            else:
                ses.feed(x)
                ses.run_async(ne)
            pl = ses.fetch(ne)
            assert int(pl["epochsDone"].min()) == ne, (it, pl["epochsDone"])
            key = np.stack([pl["I_P"], pl["Q_P"], pl["carrFreq"], pl["codeFreq"]])
            if ref is None:
                ref = key
            else:
                np.testing.assert_array_equal(key, ref, err_msg=f"run {it} differs")


def test_run_window_on_an_arriving_device_record():
    """bds_track_run_window: successive launches over a growing valid prefix of a device buffer (how the multi-GPU
    end-to-end path tracks a record that arrives over NVLink) equal the one-launch result bit for bit."""
    import torch
    s, sats, x, ch = util.record("WB", 2, 0.13)
    ps = util.product_settings(s)
    N = 9
    with _track.TrackSession("WB", ps, ch) as ses:
        ses.feed(x)
        ses.run_async(N)
        want = ses.fetch(N, raw=True)
    xd = torch.zeros(x.size + 64, dtype=torch.int8, device="cuda")
    with _track.TrackSession("WB", ps, ch) as ses:
        for frac in (0.3, 0.55, 0.8, 1.0):
            n_avail = int(x.size * frac)
            lo = int(x.size * {0.3: 0.0, 0.55: 0.3, 0.8: 0.55, 1.0: 0.8}[frac])
            xd[lo:n_avail].copy_(torch.from_numpy(x[lo:n_avail]))
            torch.cuda.synchronize()
            ses.run_window(xd.data_ptr(), n_avail if frac < 1.0 else x.size + 32, N)
        got = ses.fetch(N, raw=True)
    assert list(got["epochsDone"]) == [N, N]
    for k in want:
        np.testing.assert_array_equal(got[k], want[k], err_msg=k)


@pytest.mark.parametrize("mode,kernel", [("WB", "fast"), ("WB", "general"), ("B2a", "fast")])
def test_device_record_that_ends_on_the_last_sample_of_an_epoch(mode, kernel):
    """A caller-owned device record is used in place, all n samples of it: the epoch whose block ends on the record's
    last sample completes (WB_tracking.m:265-283 reads exactly blksize samples), nothing is read past it, and the next
    epoch is the short read."""
    s, sats, x, ch = util.record(mode, 2, 0.13 if mode == "WB" else 0.02)
    ps = util.product_settings(s)
    kern = L.KERNEL_GENERAL if kernel == "general" else L.KERNEL_FAST
    N = 6
    ref, _ = _track.run_tracking(mode, x, ch[:1], ps, n_epochs=N + 1, kernel=kern, raw=True)
    # the first channel's block of epoch N-1 ends at absoluteSample[N] (start of the following block)
    n_exact = int(ref[0].absoluteSample[N])
    p = C.c_void_p()
    L.check(L.lib().bds_dev_alloc(C.byref(p), n_exact))          # no slack on purpose
    try:
        L.check(L.lib().bds_memcpy_h2d(p, L.ptr(x), n_exact))
        st1 = ps.copy()
        st1.numberOfChannels = 1
        with _track.TrackSession(mode, st1, ch[:1], kernel=kern, device_ptr=p.value, n_samples=n_exact) as ses:
            ses.run_async(N + 1)
            got = ses.fetch(N + 1, raw=True)
        assert int(got["epochsDone"][0]) == N
        for k in ("absoluteSample", "carrFreq", "codeFreq", "remCodePhase"):
            np.testing.assert_array_equal(got[k][0, :N], ref[0][k][:N], err_msg=k)
        np.testing.assert_array_equal(got["raw"][0, :N - 1], ref[0].raw[:N - 1])
        # the last epoch: with no slack behind the record the chip-synchronous kernels stage whole 16-byte pieces only and
        # take the final chips through their exact per-sample path (a few Q8 units of different rounding); the general
        # kernel is bit-identical
        sc = util.family_scale(ref[0].raw[N - 1][None, :])[0]
        err = np.abs(got["raw"][0, N - 1] - ref[0].raw[N - 1]) / sc
        assert np.max(err[np.isfinite(err)]) <= (0.0 if kernel == "general" else 1e-6)
    finally:
        L.lib().bds_dev_free(p)


def test_file_backed_session_reads_only_what_the_epochs_need(tmp_path):
    """bds_track_open_file behind a large skipNumberOfBytes and a long (sparse) tail: the requested epochs are tracked
    from the part of the file they touch (the reference reads one block per epoch), a second run resumes where the
    first stopped, and the result equals the resident-record run."""
    s, sats, x, ch = util.record("WB", 2, 0.13)
    ps = util.product_settings(s)
    skip = 3 * 4096 + 123
    ps_skip = ps.copy()
    ps_skip.skipNumberOfBytes = skip
    path = tmp_path / "if_skip.bin"
    with open(path, "wb") as f:
        f.write(bytes(skip))
        f.write(x.tobytes())
        f.truncate(skip + x.size + (6 << 30))       # 6 GiB of sparse zeros after the record
    want, _ = _track.run_tracking("WB", x, ch, ps, n_epochs=9, raw=True)
    with _track.TrackSession("WB", ps_skip, ch, source=str(path)) as ses:
        ses.run_async(4)
        ses.sync()
        ses.run_async(5)
        got = ses.fetch(9, raw=True)
    assert list(got["epochsDone"]) == [9, 9]
    for c in range(2):
        np.testing.assert_array_equal(got["absoluteSample"][c], want[c].absoluteSample + skip)
        for k in ("carrFreq", "codeFreq", "I_P", "Q_P"):
            np.testing.assert_array_equal(got[k][c], want[c][k], err_msg=k)


def test_frame_sync_on_device_equals_oracle():
    """bds_frame_sync against BCNAV1decoding.m:66-91 / BCNAV2decoding.m:69-97 as restated by the
    oracle: the whole correlation vector and the index list, bit for bit; both polarities, a pattern cut off by the end
    of the record, values that are exactly zero (which the reference maps to -1)."""
    import bds_oracle as O
    import bds3_b200 as B
    rng = np.random.default_rng(9)
    s = O.initSettings_B1C(pilotTRKflag=2)
    for prn in (1, 19, 63):
        assert np.array_equal(B.b1c.generate2ndCode(prn), O.generate2ndCode(prn))
        n = 5000 + prn
        v = rng.standard_normal(n) * 100.0
        sec = O.generate2ndCode(prn)
        v[300:2100] = sec * (1.0 + np.abs(v[300:2100]))
        v[2100:3900] = -sec * (1.0 + np.abs(v[2100:3900]))
        v[n - 900:] = sec[:900] * 7.0                      # truncated period at the end of the record
        v[rng.integers(0, n, 40)] = 0.0
        tr = O.Settings(PRN=prn, Pilot_I_P=v, Pilot_Q_P=-v)
        X, idx = O.frame_sync_B1C(tr, s)
        gX, gidx = B.b1c.frameSync(B.settings.Struct(PRN=prn, Pilot_I_P=v, Pilot_Q_P=-v), util.product_settings(s))
        np.testing.assert_array_equal(gX, X)
        np.testing.assert_array_equal(gidx, idx)
    ip = rng.standard_normal(30000)
    pre = np.kron([-1, -1, -1, 1, 1, 1, -1, 1, 1, -1, 1, 1, -1, -1, 1, -1, -1, -1, -1, 1, -1, 1, 1, 1], [1, 1, 1, -1, 1])
    for o in (17, 3017, 6017, 29950):
        ip[o:o + 120] = (pre * (1 if o % 2 else -1))[:30000 - o]
    X, idx = O.frame_sync_B2a(ip)
    gX, gidx = B.b2a.frameSync(ip)
    np.testing.assert_array_equal(gX, X)
    np.testing.assert_array_equal(gidx, idx)
    assert {18, 3018, 6018} <= set(gidx.tolist())


@pytest.mark.parametrize("mode,kernel", [("WB", "fast"), ("NB", "fast"), ("WB", "general"), ("B2a", "fast")])
def test_lock_loss_status_and_early_channel_drop(mode, kernel):
    """settings.lockLossPLD (an extension that is off by default): a channel whose signal vanishes
    is dropped at the end of the C/N0 interval in which its lock detector has been low for lockLossIntervals
    intervals - by the rule applied to the device's own lock-detector plane -, it keeps status '-', its later epochs
    keep the preallocation values, and the other channel's results are bit-identical to a run without the option."""
    from bds3_b200 import synth
    sig = "B2a" if mode == "B2a" else "B1C"
    s = util.settings_for(mode, numberOfChannels=2)
    s.CNoInterval = 5 if sig == "B1C" else 20
    ep = 0.01 if sig == "B1C" else 0.001
    sats = synth.make_sats(2, s, sig, seed=11, cn0=50.0, max_doppler=100.0 if sig == "B2a" else 4500.0)
    n_on, n_off, N = int(30 * ep * util.FS), int(66 * ep * util.FS), 90
    if sig == "B2a":
        n_on, n_off, N = int(300 * ep * util.FS), int(660 * ep * util.FS), 900
    ps0 = util.product_settings(s)
    x = np.concatenate([synth.synth_device(sig, ps0, sats, n_on, seed=11),      # rendered on the device (fast); any record will do
                        synth.synth_device(sig, ps0, [sats[1]], n_off, seed=12, first_sample=n_on)])   # PRN 1 vanishes
    ch = synth.channels_from_sats(sats, s, sig, freq_error=0.0)
    kern = L.KERNEL_GENERAL if kernel == "general" else L.KERNEL_FAST
    ps = util.product_settings(s)
    base, _ = _track.run_tracking(mode, x, ch, ps, n_epochs=N, kernel=kern)
    ps2 = ps.copy()
    ps2.lockLossPLD, ps2.lockLossIntervals = 0.95, 3     # three intervals: the two of the pull-in do not drop the channel
    got, _ = _track.run_tracking(mode, x, ch, ps2, n_epochs=N, kernel=kern)
    assert base[0].status == "T" and base[1].status == "T" and "lockLostEpoch" not in base[0]
    ci = int(s.CNoInterval)
    pld = base[0].PilotPLD
    low, want = 0, None
    for c, v in enumerate(pld):
        low = low + 1 if v < 0.95 else 0
        if low >= 3:
            want = (c + 1) * ci
            break
    assert want is not None and want > n_on / (ep * util.FS), pld
    assert got[0].status == "-" and got[0].lockLostEpoch == want and got[0].epochsDone == want
    for f in ("I_P", "carrFreq", "absoluteSample", "PilotPLD"):
        k = want if f != "PilotPLD" else want // ci
        np.testing.assert_array_equal(got[0][f][:k], base[0][f][:k], err_msg=f)
    assert np.all(got[0].I_P[want:] == 0) and np.all(np.isinf(got[0].carrFreq[want:])) and np.all(got[0].absoluteSample[want:] == 0)
    assert got[1].status == "T"
    for f in ("I_P", "Q_P", "carrFreq", "codeFreq", "PilotPLD"):
        np.testing.assert_array_equal(got[1][f], base[1][f], err_msg=f)


# ---- fileType 2: interleaved I/Q records (WB_tracking.m:155-159,270-274; B2a tracking.m) -------------------------
@pytest.mark.parametrize("mode,seconds,epochs", [("WB", 0.05, 2), ("NB", 0.05, 2), ("B2a", 0.012, 8)])
def test_open_loop_parity_iq_record(mode, seconds, epochs):
    """complex rawSignal = I + 1i*Q: the 18 sums against the float64 (numpy) oracle, teacher forced"""
    s, sats, x, ch = util.record_iq(mode, 2, seconds)
    tr, raw = util.oracle_track(mode, s, x, ch, epochs)
    nco = np.ascontiguousarray(np.stack([t.nco for t in tr]))
    cfg = _track.make_cfg(mode, util.product_settings(s), L.KERNEL_AUTO)
    assert cfg.fileType == 2
    xi = L.as_int8(x)                       # I0, Q0, I1, Q1, ... as the file holds them
    assert xi.size == 2 * x.size
    sums = np.zeros((2, epochs, 18))
    prn = np.asarray([c.PRN for c in ch], dtype=np.int32)
    L.check(L.lib().bds_track_correlate_open_loop(MODES[mode], C.byref(cfg), L.ptr(xi), x.size, L.LOC_HOST, L.ptr(prn),
                                                  2, epochs, L.ptr(nco), L.ptr(sums)))
    err = np.abs(sums - raw) / util.family_scale(raw)
    assert np.nanmax(err[np.isfinite(err)]) <= 1e-4, np.nanmax(err[np.isfinite(err)])
    assert _track.counters(None)[2] > 0     # I/Q records run on the general kernel
    cfg.kernel = L.KERNEL_FAST              # ... and the chip-synchronous kernels refuse them
    rc = L.lib().bds_track_correlate_open_loop(MODES[mode], C.byref(cfg), L.ptr(xi), x.size, L.LOC_HOST, L.ptr(prn), 2,
                                               epochs, L.ptr(nco), L.ptr(sums))
    assert rc == -6


@pytest.mark.parametrize("mode,seconds,epochs", [("WB", 0.075, 5), ("B2a", 0.02, 15)])
def test_closed_loop_iq_record_array_and_file(mode, seconds, epochs, tmp_path):
    """closed loop on an I/Q record: complex array == the fileType-2 file it came from (skip counted in samples,
    dataAdaptCoeff*(skipNumberOfBytes + codePhase - 1)), both against the oracle trajectory"""
    s, sats, x, ch = util.record_iq(mode, 2, seconds)
    tr, raw = util.oracle_track(mode, s, x, ch, epochs)
    ps = util.product_settings(s)
    got, _ = _track.run_tracking(mode, x, ch, ps, n_epochs=epochs, raw=True)
    skip = 4096 + 3
    path = tmp_path / "iq.bin"
    pad = np.zeros(2 * skip, dtype=np.int8)
    np.concatenate([pad, L.as_int8(x)]).tofile(path)
    pf = util.product_settings(s)
    pf.skipNumberOfBytes = skip
    with open(path, "rb") as fid:
        gotf, _ = _track.run_tracking(mode, fid, ch, pf, n_epochs=epochs, raw=True)
    for c in range(len(ch)):
        g, o, f = got[c], tr[c], gotf[c]
        assert g.status == "T" and f.status == "T"
        np.testing.assert_array_equal(g.absoluteSample, o.absoluteSample)
        np.testing.assert_array_equal(f.absoluteSample, o.absoluteSample + skip)     # ftell/dataAdaptCoeff
        # the file's samples sit at another 16-byte phase (skip): same arithmetic, another fp32 summation grouping
        assert np.max(np.abs(f.raw - g.raw) / util.family_scale(g.raw)) <= 1e-5
        sc = util.family_scale(raw[c])
        err = np.abs(g.raw - raw[c]) / sc
        assert np.max(err) <= 1e-3, np.max(err)
        assert np.mean(err <= 1e-4) >= 0.99
        for k in ("carrFreq", "codeFreq"):
            np.testing.assert_allclose(g[k], o[k], rtol=1e-9)


@pytest.mark.parametrize("mode", ["WB", "NB"])
def test_chip_synchronous_kernel_at_the_reference_shipped_53_mhz(mode):
    """The reference ships B1C with samplingFreq = 53 MHz (B1C/initSettings.m:57).  The chip-synchronous kernel has a
    second instantiation for that geometry (51.8 samples per chip, 1/1024-sample rank bins): AUTO picks it, open-loop
    sums equal the oracle's at 1e-4, the closed loop is checked one step at a time."""
    import bds_oracle as O
    from bds3_b200 import synth
    s = O.initSettings_B1C(pilotTRKflag=2 if mode == "WB" else 1, numberOfChannels=2)
    assert s.samplingFreq == 53e6 and s.IF == 1590e6 - 1575.42e6
    sats = synth.make_sats(2, s, "B1C", seed=21, sigma=25.0, max_doppler=4500.0)
    x = synth.synth_numpy("B1C", s, sats, int(0.23 * s.samplingFreq), sigma=25.0, seed=21)
    ch = synth.channels_from_sats(sats, s, "B1C", freq_error=2.0)
    tr, raw = util.oracle_track(mode, s, x, ch, 3)
    nco = np.stack([t.nco for t in tr])
    for wide_guard in (0, 1):
        cfg = _track.make_cfg(mode, util.product_settings(s), L.KERNEL_FAST)
        cfg.reserved = wide_guard
        sums = np.zeros((2, 3, 18))
        prn = np.asarray([c.PRN for c in ch], dtype=np.int32)
        L.check(L.lib().bds_track_correlate_open_loop(MODES[mode], C.byref(cfg), L.ptr(x), x.size, L.LOC_HOST, L.ptr(prn),
                                                      2, 3, L.ptr(np.ascontiguousarray(nco)), L.ptr(sums)))
        err = np.abs(sums - raw) / util.family_scale(raw)
        assert np.max(err[np.isfinite(err)]) <= 1e-4, np.max(err[np.isfinite(err)])
        fast_chips, exact_chips, general_slices, _ = _track.counters(None)
        assert general_slices == 0 and fast_chips + exact_chips == 2 * 3 * 10230
        # wide guard: 2^-8 sample on either side of every threshold (36 wide band, 6 narrow band) and of the chip edges
        lo, hi = (0.1, 0.6) if mode == "WB" else (0.03, 0.16)
        assert (lo < exact_chips / (2 * 3 * 10230) < hi) if wide_guard else exact_chips <= 2 * 3 + 2
    ps = util.product_settings(s)
    res, _ = _track.run_tracking(mode, x, ch, ps, n_epochs=20, raw=True)            # AUTO
    fast_chips, exact_chips, general_slices, _ = _track.run_tracking.last_counters
    assert general_slices == 0 and fast_chips + exact_chips == 2 * 20 * 10230 and exact_chips <= 64
    for c in range(2):
        assert res[c].status == "T"
        assert util.one_step_parity(mode, s, x, ch[c], res[c], 20) <= 1e-4
    gen, _ = _track.run_tracking(mode, x, ch, ps, n_epochs=20, kernel=L.KERNEL_GENERAL, raw=True)
    for c in range(2):   # the two kernels follow the same trajectory up to chip-edge chaos
        np.testing.assert_allclose(res[c].carrFreq, gen[c].carrFreq, rtol=0, atol=0.05)
        np.testing.assert_allclose(res[c].absoluteSample, gen[c].absoluteSample, rtol=0, atol=1)


def test_sixty_channels_400_epochs_closed_loop_proves_itself():
    """The headline configuration (60 B1C wide-band channels, fs = 99.375 MHz) over 4 s of signal, checked the way
    bench.py checks its timed 30 s run: every channel completes every epoch; the loop bookkeeping of every epoch of every
    channel follows from the device's own discriminators through the reference's loop filters (float64 replay); all 60
    channels hold lock over the last second (PLD > 0.9, C/N0 in the expected band); sampled one-step parity at the end of
    the record <= 1e-4; exact-path share < 1e-4."""
    import math
    import torch
    import bench
    from bds3_b200 import synth
    seconds, fs, spc = 4.05, 99.375e6, 993750
    st = bench.settings_for("track", 60, seconds)
    sats = synth.make_sats(60, st, "B1C", max_doppler=4500.0)
    chans = synth.channels_from_sats(sats, st, "B1C", freq_error=2.0)
    n = int(round(seconds * fs))
    n_epochs = int(math.floor((n - spc) / (spc * (1 + 1e-5)))) - 1
    assert n_epochs >= 400
    x_dev = torch.empty(n + 64, dtype=torch.int8, device="cuda")
    synth.synth_device("B1C", st, sats, n, out_ptr=x_dev.data_ptr())
    torch.cuda.synchronize()
    st.numberOfChannels_total = 60
    sess = _track.TrackSession("WB", st, chans, kernel=L.KERNEL_AUTO, device_ptr=x_dev.data_ptr(), n_samples=n)
    try:
        sess.run_async(n_epochs)
        sess.sync()
        chk = bench.self_check(sess, st, chans, n_epochs, x_dev, n, mode="WB", n_sampled_channels=3, n_sampled_epochs=10,
                               seconds_tail=1.0)
    finally:
        sess.close()
    assert chk["channels"] == 60 and chk["locked_channels"] == 60
    assert chk["parity_max_rel"] <= 1e-4 and chk["exact_chip_frac"] < 1e-4


def test_one_second_trajectory_against_the_stored_oracle_trajectory():
    """Trajectory against trajectory over 100 epochs = 1 s: the device's closed loop next to the float64 oracle's closed
    loop stored in tests/golden/wb_trajectory_100.npz (tests/golden/make_trajectory.py: five minutes of a host core, so it
    is not recomputed here).  Bounds as in test_closed_loop_fast_tracks_oracle_trajectory: the first epoch at the
    north-star tolerance, afterwards the two loops stay together up to the chaos of single samples crossing chip edges
    (0.05 Hz, 1e-3 chip, one sample); the measured divergence is printed (pytest -s) and recorded in profiles/r03."""
    import json
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wb_trajectory_100.npz"))
    n = int(g["n_epochs"])
    s, sats, x, ch = util.record("WB", 2, float(g["seconds"]))
    if np.frombuffer(x[:1 << 20].tobytes(), dtype=np.uint8).astype(np.uint64).sum() != g["x_sha_head"]:
        pytest.skip("numpy renders the synthetic record differently on this host (libm): the stored trajectory is of another record")
    assert int(ch[0].PRN) == int(g["prn"])
    res, _ = _track.run_tracking("WB", x, ch, util.product_settings(s), n_epochs=n, raw=True)
    fast_chips, exact_chips, general_slices, _ = _track.run_tracking.last_counters
    assert general_slices == 0 and fast_chips + exact_chips == 2 * n * 10230
    r = res[0]
    err0 = np.abs(r.raw[0] - g["raw"][0]) / util.family_scale(g["raw"][0][None, :])[0]
    assert np.max(err0) <= 1e-4
    prompt = np.hypot(r.I_P - g["I_P"], r.Q_P - g["Q_P"]) / np.hypot(g["I_P"], g["Q_P"])
    div = {"epochs": n, "max_abs_carrFreq_Hz": float(np.max(np.abs(r.carrFreq - g["carrFreq"]))),
           "max_abs_codeFreq_Hz": float(np.max(np.abs(r.codeFreq - g["codeFreq"]))),
           "max_abs_code_phase_chips": float(np.max(np.abs(r.remCodePhase - (r.absoluteSample - g["absoluteSample"])
                                                             * (g["codeFreq"] / s.samplingFreq) - g["remCodePhase"]))),
           "max_abs_absoluteSample": float(np.max(np.abs(r.absoluteSample - g["absoluteSample"]))),
           "epochs_with_identical_absoluteSample": int(np.sum(r.absoluteSample == g["absoluteSample"])),
           "max_rel_prompt": float(np.max(prompt)), "first_epoch_max_rel_18_sums": float(np.max(err0))}
    print("trajectory divergence over 1 s:", json.dumps(div))
    out = os.environ.get("BDS_TRAJECTORY_JSON")
    if out:
        os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
        with open(out, "w") as f:
            json.dump(div, f)
    np.testing.assert_allclose(r.carrFreq, g["carrFreq"], rtol=0, atol=0.05)          # Hz
    np.testing.assert_allclose(r.absoluteSample, g["absoluteSample"], rtol=0, atol=1)
    # an epoch whose block starts one sample later in one of the loops (blksize = ceil((L - rem) / step) at a boundary) carries
    # one codePhaseStep more remCodePhase: compare the code phase referred to the oracle's block start
    rem = r.remCodePhase - (r.absoluteSample - g["absoluteSample"]) * (g["codeFreq"] / s.samplingFreq)
    np.testing.assert_allclose(rem, g["remCodePhase"], rtol=0, atol=1e-3)             # chips


def test_narrow_band_body_against_oracle_and_wide_band_body():
    """NB_tracking on its own chip body (six segments per chip, the AUTO / FAST choice) and on the wide-band body
    (cfg.reserved bit 1, the path before the narrow-band bodies existed): both within the north-star tolerance of the
    oracle, the BOC(6,1) family exactly zero, and equal to each other to float rounding."""
    s, sats, x, ch = util.record("NB", 2, 0.06)
    tr, raw = util.oracle_track("NB", s, x, ch, 3)
    nco = np.ascontiguousarray(np.stack([t.nco for t in tr]))
    prn = np.asarray([c.PRN for c in ch], dtype=np.int32)
    got = {}
    for hook in (0, 2, 1, 3):       # bit 0: wide guard band (a quarter of the chips through the exact path)
        cfg = _track.make_cfg("NB", util.product_settings(s), L.KERNEL_FAST)
        cfg.reserved = hook
        sums = np.zeros((2, 3, 18))
        L.check(L.lib().bds_track_correlate_open_loop(L.TRK_B1C_NB, C.byref(cfg), L.ptr(x), x.size, L.LOC_HOST, L.ptr(prn),
                                                      2, 3, L.ptr(nco), L.ptr(sums)))
        err = np.abs(sums - raw) / util.family_scale(raw)
        assert np.max(err[np.isfinite(err)]) <= 1e-4, (hook, np.max(err[np.isfinite(err)]))
        assert np.all(sums[..., 12:] == 0.0)
        fast_chips, exact_chips, general_slices, _ = _track.counters(None)
        assert general_slices == 0 and fast_chips + exact_chips == 2 * 3 * 10230
        got[hook] = sums
    sc = util.family_scale(raw)
    d = np.abs(got[0] - got[2]) / sc
    assert np.max(d[np.isfinite(d)]) <= 2e-5

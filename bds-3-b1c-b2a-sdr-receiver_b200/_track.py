"""Host side of tracking: marshals ``settings``/``channel`` into the C ABI and assembles the
``trackResults`` struct array exactly as the reference lays it out:
field creation order, initial values (0 / Inf / '-') and 1xN double row vectors of
BDS-3_B1C/WB_tracking.m:53-112, NB_tracking.m:53-100 and BDS-3_B2a/tracking.m:48-96."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib as L
from . import loopcoef
from .settings import Struct, matlab_round

_MODE = {"WB": L.TRK_B1C_WB, "NB": L.TRK_B1C_NB, "B2a": L.TRK_B2A}


def num_to_process(mode, settings) -> int:
    if mode == "B2a":
        return int(settings.msToProcess)                                   # B2a/tracking.m:100
    return matlab_round(settings.msToProcess / 1000 / settings.intTime)    # WB_tracking.m:56


def has_pilot(mode, settings) -> bool:
    f = settings.pilotTRKflag
    return (mode == "WB" and f == 2) or (mode in ("NB", "B2a") and f == 1)


def make_cfg(mode, settings, kernel=L.KERNEL_AUTO, tuning=None) -> L.bds_trk_cfg:
    """``tuning``: optional dict of the bds_trk_cfg tuning / diagnostics fields (fwPassesPerTask, fwPrefetch, debug,
    traceTickets); the library itself never reads the environment."""
    tau1, tau2 = loopcoef.calcLoopCoef(settings.dllNoiseBandwidth, settings.dllDampingRatio, 1.0)
    pf3, pf2, pf1 = loopcoef.calcLoopCoefCarr(settings)
    factor = loopcoef.CalcWeighingFactor(settings) if mode == "WB" else 0.0   # WB_tracking.m:138
    cfg = L.bds_trk_cfg(samplingFreq=settings.samplingFreq, codeFreqBasis=settings.codeFreqBasis,
                        codeLength=int(settings.codeLength), dllCorrelatorSpacing=settings.dllCorrelatorSpacing,
                        intTime=settings.intTime, pilotTRKflag=int(settings.pilotTRKflag),
                        CNoInterval=int(settings.CNoInterval), tau1code=tau1, tau2code=tau2, pf3=pf3, pf2=pf2,
                        pf1=pf1, wbFactor=factor, kernel=int(kernel), reserved=0,
                        fileType=int(settings.get("fileType", 1)))
    # lock-loss status / early channel drop: an extension of the reference, off unless the settings carry it
    if settings.get("lockLossPLD", 0):
        cfg.lockLossPLD = float(settings.lockLossPLD)
        cfg.lockLossIntervals = int(settings.get("lockLossIntervals", 1))
    for k, v in (tuning or {}).items():
        setattr(cfg, k, int(v))
    return cfg


def make_channels(channel):
    arr = (L.bds_channel * len(channel))()
    for i, ch in enumerate(channel):
        arr[i].PRN = int(ch.PRN)
        arr[i].status = ord(ch.status[0]) if isinstance(ch.status, str) else int(ch.status)
        arr[i].acquiredFreq = float(ch.acquiredFreq)
        arr[i].codePhase = float(ch.codePhase)
        arr[i].codeFreq = float(ch.codeFreq)
    return arr


def template(mode, settings, N) -> Struct:
    tr = Struct()
    tr.status = "-"
    tr.absoluteSample = np.zeros(N)
    tr.codeFreq = np.full(N, np.inf)
    tr.carrFreq = np.full(N, np.inf)
    for f in ("I_P", "I_E", "I_L", "Q_E", "Q_P", "Q_L"):
        tr[f] = np.zeros(N)
    if has_pilot(mode, settings):
        names = (("Pilot_I_P", "Pilot_I_E", "Pilot_I_L", "Pilot_Q_E", "Pilot_Q_P", "Pilot_Q_L") if mode == "WB"
                 else ("Pilot_I_P", "Pilot_Q_P"))
        for f in names:
            tr[f] = np.zeros(N)
    for f in ("dllDiscr", "dllDiscrFilt", "pllDiscr", "pllDiscrFilt", "remCodePhase", "remCarrPhase"):
        tr[f] = np.full(N, np.inf)
    nc = N // int(settings.CNoInterval)
    tr.DataCNo = np.zeros(nc)
    tr.DataPLD = np.zeros(nc)
    if has_pilot(mode, settings):
        tr.PilotCNo = np.zeros(nc)
        tr.PilotPLD = np.zeros(nc)
        tr["B2a_CNo" if mode == "B2a" else "B1C_CNo"] = np.zeros(nc)
    return tr


class TrackSession:
    """Thin RAII wrapper over bds_trk* (bds_track_open / run / fetch / close)."""

    def __init__(self, mode, settings, channel, source=None, kernel=L.KERNEL_AUTO, device_ptr=None, n_samples=None,
                 tuning=None):
        """``source``: host samples (array / np.memmap) or a file / path (the reference's ``fid``); they are
        streamed to the device under the tracking kernel by the first ``run_async``.  ``device_ptr``: a record
        already resident in HBM (used in place)."""
        self.mode, self.settings = mode, settings
        self.nch = len(channel)
        self.cfg = make_cfg(mode, settings, kernel, tuning)
        if isinstance(source, np.ndarray) and np.iscomplexobj(source):
            self.cfg.fileType = 2     # rawSignal = I + 1i*Q (WB_tracking.m:270-274): uploaded as the file's I, Q byte pairs
        self.bps = 2 if self.cfg.fileType == 2 else 1   # bytes per sample
        self.chs = make_channels(channel)
        self.h = C.c_void_p()
        lib = L.lib()
        skip = int(settings.get("skipNumberOfBytes", 0))
        self._host = self._keep = None
        if device_ptr is not None:
            L.check(lib.bds_track_open(_MODE[mode], C.byref(self.cfg), C.c_void_p(device_ptr), int(n_samples),
                                       L.LOC_DEVICE, skip, self.chs, self.nch, C.byref(self.h)))
        elif isinstance(source, (str, os.PathLike)) or (hasattr(source, "name") and not isinstance(source, np.ndarray)):
            path = os.fspath(source if isinstance(source, (str, os.PathLike)) else source.name)
            L.check(lib.bds_track_open_file(_MODE[mode], C.byref(self.cfg), path.encode(), skip, 0, self.chs, self.nch,
                                            C.byref(self.h)))
        else:
            L.check(lib.bds_track_open(_MODE[mode], C.byref(self.cfg), None, 0, L.LOC_HOST, skip, self.chs,
                                       self.nch, C.byref(self.h)))
            if source is not None:
                self._host = L.as_int8(source)

    def close(self):
        if self.h:
            L.lib().bds_track_close(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def feed(self, x=None, first_sample=0, device_ptr=None, n=None):
        if device_ptr is not None:
            L.check(L.lib().bds_track_feed(self.h, C.c_void_p(device_ptr), int(n), L.LOC_DEVICE, int(first_sample)))
        else:
            x = L.as_int8(x)
            self._keep = x
            self._host = None
            L.check(L.lib().bds_track_feed(self.h, L.ptr(x), x.size // self.bps, L.LOC_HOST, int(first_sample)))

    def run_async(self, n_epochs):
        if self._host is not None:      # host record not uploaded yet: stream it under the kernel
            x, self._host, self._keep = self._host, None, self._host
            L.check(L.lib().bds_track_run_streamed(self.h, L.ptr(x), x.size // self.bps, 0, int(n_epochs)))
        else:
            L.check(L.lib().bds_track_run_async(self.h, int(n_epochs)))

    def run_streamed(self, host_ptr, n, n_epochs, chunk_bytes=0):
        """Track a host record given by address (e.g. pinned memory): H2D in chunks overlapped with tracking."""
        L.check(L.lib().bds_track_run_streamed(self.h, C.c_void_p(host_ptr), int(n), int(chunk_bytes), int(n_epochs)))

    def run_window(self, device_ptr, n_avail, epoch_limit):
        """One asynchronous launch over the first ``n_avail`` samples of a device record that is still arriving."""
        L.check(L.lib().bds_track_run_window(self.h, C.c_void_p(device_ptr), int(n_avail), int(epoch_limit)))

    def sync(self):
        L.check(L.lib().bds_track_sync(self.h))

    def reset(self):
        L.check(L.lib().bds_track_reset(self.h))

    def stats(self):
        cs, ep, ms = C.c_longlong(), C.c_int(), C.c_float()
        L.check(L.lib().bds_track_stats(self.h, C.byref(cs), C.byref(ep), C.byref(ms)))
        return cs.value, ep.value, ms.value

    def counters(self):
        return counters(self.h)

    def device_block(self):
        p, b, nf, cap = C.c_void_p(), C.c_size_t(), C.c_int(), C.c_int()
        L.check(L.lib().bds_track_device_block(self.h, C.byref(p), C.byref(b), C.byref(nf), C.byref(cap)))
        return p.value, b.value, nf.value, cap.value

    def fetch(self, N, raw=False, into=None):
        """-> dict of [nch, N] planes (+ CNo planes [nch, N//CNoInterval], raw [nch,N,18], epochsDone [nch]).
        ``into``: dict name -> preallocated float64 [nch, N] arrays (e.g. pinned memory) to receive the planes."""
        out = L.bds_trk_out()
        planes = {}
        for name in L.TRK_PLANES:
            a = into[name] if into is not None else np.empty((self.nch, N))
            assert a.shape == (self.nch, N) and a.dtype == np.float64 and a.flags.c_contiguous
            planes[name] = a
            setattr(out, name, a.ctypes.data_as(L._PD))
        nc = N // int(self.settings.CNoInterval)
        for name in L.CNO_PLANES:
            a = np.zeros((self.nch, nc))
            planes[name] = a
            if nc > 0:
                setattr(out, name, a.ctypes.data_as(L._PD))
        if raw:
            r = np.zeros((self.nch, N, 18))
            planes["raw"] = r
            out.raw = r.ctypes.data_as(L._PD)
        done = np.zeros(self.nch, dtype=np.int32)
        out.epochsDone = done.ctypes.data_as(C.POINTER(C.c_int32))
        planes["epochsDone"] = done
        lost = np.zeros(self.nch, dtype=np.int32)
        out.lockLostEpoch = lost.ctypes.data_as(C.POINTER(C.c_int32))
        planes["lockLostEpoch"] = lost
        L.check(L.lib().bds_track_fetch(self.h, C.byref(out), N))
        return planes


def counters(handle=None):
    """(fast-body chips, exact-path chips, general-kernel slices, 0); handle None = last open-loop call."""
    out = (C.c_longlong * 4)()
    L.check(L.lib().bds_track_counters(handle, out))
    return tuple(int(v) for v in out)


def assemble(mode, settings, channel, planes, N):
    """trackResults struct array from the fetched planes, mirroring the reference's control flow:
    channels are processed in order; a short read leaves that channel partially filled with
    status '-' and *returns*, so later channels keep the template (WB_tracking.m:279-283,485-488)."""
    results = []
    stopped = False
    pilot = has_pilot(mode, settings)
    cno_name = "B2a_CNo" if mode == "B2a" else "B1C_CNo"
    for c, ch in enumerate(channel):
        tr = template(mode, settings, N)
        if ch.PRN != 0 and not stopped:
            done = int(planes["epochsDone"][c])
            for f in list(tr.keys()):
                if f in L.TRK_PLANES:
                    tr[f] = planes[f][c].copy()
            ncd = min(tr.DataCNo.size, done // int(settings.CNoInterval))
            tr.DataCNo[:ncd] = planes["DataCNo"][c, :ncd]
            tr.DataPLD[:ncd] = planes["DataPLD"][c, :ncd]
            if pilot:
                tr.PilotCNo[:ncd] = planes["PilotCNo"][c, :ncd]
                tr.PilotPLD[:ncd] = planes["PilotPLD"][c, :ncd]
                tr[cno_name][:ncd] = planes["TotalCNo"][c, :ncd]
            tr.PRN = int(ch.PRN)                                        # WB_tracking.m:167
            lost = int(planes["lockLostEpoch"][c]) if "lockLostEpoch" in planes else 0
            if done >= N:
                tr.status = ch.status                                    # WB_tracking.m:485-488
            elif lost:
                tr.lockLostEpoch = lost      # extension (settings.lockLossPLD): dropped for loss of lock, status stays '-',
            else:                            # the following channels are tracked (not the reference's bare return)
                stopped = True
            if "raw" in planes:
                tr.raw = planes["raw"][c].copy()
            tr.epochsDone = done
        results.append(tr)
    return results


def run_tracking(mode, source, channel, settings, n_epochs=None, kernel=L.KERNEL_AUTO, raw=False):
    N = num_to_process(mode, settings) if n_epochs is None else int(n_epochs)
    if not any(ch.PRN != 0 for ch in channel):
        return [template(mode, settings, N) for _ in channel], channel
    with TrackSession(mode, settings, channel, source, kernel=kernel) as s:
        s.run_async(N)
        planes = s.fetch(N, raw=raw)
        run_tracking.last_counters = s.counters()
    return assemble(mode, settings, channel, planes, N), channel


def frame_sync(signal, prompt, prn=0, want_xcorr=True, index_cap=4096):
    """(XcorrResult for lags >= 0, index (1-based, ascending)) of bds_frame_sync: the correlation that opens the
    reference's nav decoding (BCNAV1decoding.m:66-91 / BCNAV2decoding.m:69-97), on the device."""
    v = np.ascontiguousarray(prompt, dtype=np.float64).reshape(-1)
    x = np.zeros(v.size) if want_xcorr else None
    idx = np.zeros(index_cap, dtype=np.int32)
    n = C.c_int32(0)
    L.check(L.lib().bds_frame_sync(int(signal), int(prn), L.ptr(v), v.size, L.LOC_HOST, L.ptr(x) if want_xcorr else None,
                                   L.ptr(idx), index_cap, C.byref(n)))
    if n.value > index_cap:
        return frame_sync(signal, prompt, prn, want_xcorr, n.value)
    return x, idx[:n.value].astype(np.int64)

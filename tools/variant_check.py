"""Developer A/B measurement of alternative builds of the tracking kernel (no torch: ctypes + numpy only, so a
process starts in a second or two on a fresh GPU box).

    BDS_LIB_NAME=libbds_x.so python tools/variant_check.py closed OUT_PREFIX [seconds]
        60-channel B1C WB closed loop on a synthetic record rendered on the device; prints the kernel time and
        x real time, writes OUT_PREFIX.npz (output planes + the NCO trajectory for the open-loop mode).
    BDS_LIB_NAME=libbds_x.so python tools/variant_check.py open NCO.npz OUT_PREFIX
    [BDS_TRK_B2A_UNIT=1] python tools/variant_check.py closed_b2a OUT_PREFIX [seconds]
        60-channel B2a closed loop (general kernel, or the per-channel kernel with BDS_TRK_B2A_UNIT=1)
        the same correlator teacher-forced with that trajectory (no loop closure): usable with ablation builds
        whose sums are wrong on purpose.

Build a variant with e.g.
    BDS_LIB_NAME=libbds_f2.so BDS_OBJ_SUFFIX=_f2 BDS_EXTRA_FLAGS=-DBDS_FAST_F32X2=1 python <package>/build.py
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bds3_b200 as B  # noqa: E402
from bds3_b200 import _lib as L, _track, synth  # noqa: E402

FS = 99.375e6
NCH = int(os.environ.get("BDS_NCH", 60))


def tuning():
    """developer knobs of these tools (the library itself never reads the environment): BDS_TRK_PASSES, BDS_TRK_AHEAD,
    BDS_TRK_TIMING -> bds_trk_cfg.fwPassesPerTask / fwPrefetch / debug"""
    t = {}
    if "BDS_TRK_PASSES" in os.environ:
        t["fwPassesPerTask"] = int(os.environ["BDS_TRK_PASSES"])
    if "BDS_TRK_AHEAD" in os.environ:
        t["fwPrefetch"] = int(os.environ["BDS_TRK_AHEAD"]) + 1
    if "BDS_B2A_CS" in os.environ:
        t["b2aClusterSize"] = int(os.environ["BDS_B2A_CS"])
    if os.environ.get("BDS_TRK_TIMING"):
        t["debug"] = L.DBG_TIMING
    return t


def render(seconds):
    st = B.b1c.initSettings(samplingFreq=FS, numberOfChannels=NCH, pilotTRKflag=2, msToProcess=int(seconds * 1000))
    sats = synth.make_sats(NCH, st, "B1C")
    ch = synth.channels_from_sats(sats, st, "B1C", freq_error=2.0)
    n = int(round(seconds * FS))
    p = C.c_void_p()
    L.check(L.lib().bds_dev_alloc(C.byref(p), n + 64))
    synth.synth_device("B1C", st, sats, n, out_ptr=p.value)
    L.check(L.lib().bds_dev_sync())
    return st, ch, n, p.value


def main():
    mode = sys.argv[1]
    t00 = time.time()
    L.init(0)
    res = {"lib": os.path.basename(L.lib_path()), "mode": mode}
    if mode == "closed_b2a":
        out = sys.argv[2]
        seconds = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
        st = B.b2a.initSettings(numberOfChannels=NCH, msToProcess=int(seconds * 1000))
        sats = synth.make_sats(NCH, st, "B2a", max_doppler=100.0)
        ch = synth.channels_from_sats(sats, st, "B2a", freq_error=2.0)
        n = int(round(seconds * FS))
        p = C.c_void_p()
        L.check(L.lib().bds_dev_alloc(C.byref(p), n + 64))
        synth.synth_device("B2a", st, sats, n, out_ptr=p.value)
        L.check(L.lib().bds_dev_sync())
        ne = max(1, int(n // 99376) - 2)
        s = _track.TrackSession("B2a", st, ch, device_ptr=p.value, n_samples=n, tuning=tuning())
        ms = []
        for _ in range(3):
            s.reset()
            s.run_async(ne)
            s.sync()
            ms.append(s.stats()[2])
        pl = s.fetch(ne)
        cnt = s.counters()
        res.update(seconds=seconds, epochs=ne, kernel_ms=[round(m, 3) for m in ms], unit_kernel=os.environ.get("BDS_TRK_B2A_UNIT", "0"),
                   x_realtime=round(ne * 0.001 / max(min(ms[1:]) * 1e-3, 1e-12), 1), epochs_done_min=int(pl["epochsDone"].min()),
                   fast_units=cnt[0], exact_units=cnt[1], general_slices=cnt[2])
        np.savez_compressed(out + ".npz", **{k: pl[k] for k in ("I_P", "Q_P", "carrFreq", "codeFreq", "absoluteSample", "epochsDone")})
    elif mode == "closed":
        out = sys.argv[2]
        seconds = float(sys.argv[3]) if len(sys.argv) > 3 else 5.0
        st, ch, n, xp = render(seconds)
        ne = max(1, int((n - 993750) // 993760) - 1)
        s = _track.TrackSession("WB", st, ch, device_ptr=xp, n_samples=n, tuning=tuning())
        ms = []
        for _ in range(4):
            s.reset()
            s.run_async(ne)
            s.sync()
            ms.append(s.stats()[2])
        pl = s.fetch(ne)
        cnt = s.counters()
        res.update(seconds=seconds, epochs=ne, kernel_ms=[round(m, 3) for m in ms],
                   x_realtime=round(ne * 0.01 / max(min(ms[1:]) * 1e-3, 1e-12), 1), epochs_done_min=int(pl["epochsDone"].min()),
                   fast_chips=cnt[0], exact_chips=cnt[1])
        step = pl["codeFreq"] / FS
        nco = np.zeros((NCH, ne, 6))
        nco[:, :, 0] = pl["absoluteSample"]
        nco[:, :, 1] = np.ceil((10230 - pl["remCodePhase"]) / step)
        nco[:, :, 2] = pl["remCodePhase"]
        nco[:, :, 3] = step
        nco[:, :, 4] = pl["carrFreq"]
        nco[:, :, 5] = pl["remCarrPhase"]
        keep = {k: pl[k] for k in ("I_P", "Q_P", "I_E", "Q_L", "p11_I_P", "p61_Q_E", "carrFreq", "codeFreq", "absoluteSample",
                                   "epochsDone") if k in pl}
        np.savez_compressed(out + ".npz", nco=nco, seconds=seconds, **keep)
    else:
        z = np.load(sys.argv[2])
        out = sys.argv[3]
        nco = np.ascontiguousarray(z["nco"])
        seconds = float(z["seconds"])
        ne = nco.shape[1]
        st, ch, n, xp = render(seconds)
        cfg = _track.make_cfg("WB", st, L.KERNEL_FAST)
        prn = np.asarray([c.PRN for c in ch], dtype=np.int32)
        sums = np.zeros((NCH, ne, 18))
        us = []
        for _ in range(4):
            L.check(L.lib().bds_track_correlate_open_loop(L.TRK_B1C_WB, C.byref(cfg), C.c_void_p(xp), n, L.LOC_DEVICE,
                                                          L.ptr(prn), NCH, ne, L.ptr(nco), L.ptr(sums)))
            us.append(_track.counters(None)[3])
        res.update(seconds=seconds, epochs=ne, kernel_ms=[round(u / 1e3, 3) for u in us],
                   x_realtime=round(ne * 0.01 / max(min(us[1:]) * 1e-6, 1e-12), 1))
        np.savez_compressed(out + ".npz", sums=sums)
    res["wall_s"] = round(time.time() - t00, 2)
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()

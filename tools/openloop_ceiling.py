"""Developer measurement: the chip-synchronous correlator without the closed loop (teacher-forced with the NCO
trajectory of a closed-loop run) = the compute/TMA ceiling of trk_fw_kernel for the same work."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bds3_b200 as B
from bds3_b200 import _lib as L, _track, synth

FS = 99.375e6
nch, seconds = 60, 1.05
L.init(0)
st = B.b1c.initSettings(samplingFreq=FS, numberOfChannels=nch, pilotTRKflag=2, msToProcess=1000)
sats = synth.make_sats(nch, st, "B1C")
ch = synth.channels_from_sats(sats, st, "B1C", freq_error=2.0)
n = int(seconds * FS)
x = torch.empty(n + 64, dtype=torch.int8, device="cuda")
synth.synth_device("B1C", st, sats, n, out_ptr=x.data_ptr())
torch.cuda.synchronize()
ne = 100
s = _track.TrackSession("WB", st, ch, device_ptr=x.data_ptr(), n_samples=n)
for _ in range(3):
    s.reset(); s.run_async(ne); s.sync()
cs, ep, ms = s.stats()
pl = s.fetch(ne)
print("closed loop: %d epochs, kernel %.3f ms, %.1f x real time" % (ep, ms, ne * 0.01 / (ms * 1e-3)))
nco = np.zeros((nch, ne, 6))
nco[:, :, 0] = pl["absoluteSample"]
blk = np.diff(np.concatenate([pl["absoluteSample"], pl["absoluteSample"][:, -1:] + 993750], axis=1), axis=1)
step = pl["codeFreq"] / FS
nco[:, :, 1] = np.ceil((10230 - pl["remCodePhase"]) / step)
nco[:, :, 2] = pl["remCodePhase"]; nco[:, :, 3] = step; nco[:, :, 4] = pl["carrFreq"]; nco[:, :, 5] = pl["remCarrPhase"]
assert np.all(nco[:, :-1, 1] == blk[:, :-1])
cfg = _track.make_cfg("WB", st, L.KERNEL_FAST)
prn = np.asarray([c.PRN for c in ch], dtype=np.int32)
sums = np.zeros((nch, ne, 18))
nco = np.ascontiguousarray(nco)
for _ in range(3):
    L.check(L.lib().bds_track_correlate_open_loop(L.TRK_B1C_WB, C.byref(cfg), C.c_void_p(x.data_ptr()), n, L.LOC_DEVICE,
                                                  L.ptr(prn), nch, ne, L.ptr(nco), L.ptr(sums)))
us = _track.counters(None)[3]
print("open loop  : kernel %.3f ms, %.1f x real time (ceiling of the correlator without loop closure)" % (us / 1e3, ne * 0.01 / (us * 1e-6)))

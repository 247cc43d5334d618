"""compute-sanitizer target: a small closed-loop WB run (resident + streamed), an open-loop call and a B2a acquisition."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import util
import bds3_b200 as B
from bds3_b200 import _lib as L, _track, synth
L.init(0)
s, sats, x, ch = util.record("WB", 2, 0.045)
ps = util.product_settings(s)
got, _ = _track.run_tracking("WB", x, ch, ps, n_epochs=4, raw=True)          # streamed host record, runs to the end of the record
print("wb fast   :", got[0].epochsDone, got[0].status)
got, _ = _track.run_tracking("NB", x, ch, ps, n_epochs=2, kernel=L.KERNEL_GENERAL)
print("nb general:", got[0].epochsDone)
st = B.b2a.initSettings(acqSatelliteList=[4, 9])
sats2 = synth.make_sats(1, st, "B2a", seed=3, prns=[4], cn0=47.0)
xb = synth.synth_device("B2a", st, sats2, 17 * 99375)
acq = B.b2a.acquisition(xb, st)
print("b2a acq   :", acq.carrFreq[3] != 0)

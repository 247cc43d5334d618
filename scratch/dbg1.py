import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle'); sys.path.insert(0,'/root/repo/tests')
import util
from bds3_b200 import _lib as L, _track
s, sats, x, ch = util.record("WB", 2, 0.13)
ps = util.product_settings(s)
tr, raw = util.oracle_track("WB", s, x, ch, 10)
runs=[]
for k in range(3):
    fast,_ = _track.run_tracking("WB", x, ch, ps, n_epochs=10, kernel=L.KERNEL_FAST, raw=True)
    runs.append(fast); print("counters", _track.run_tracking.last_counters)
gen,_ = _track.run_tracking("WB", x, ch, ps, n_epochs=10, kernel=L.KERNEL_GENERAL, raw=True)
for c in range(2):
    sc = util.family_scale(raw[c])
    print("ch",c,"fast vs oracle per-epoch max", np.round(np.max(np.abs(runs[0][c].raw-raw[c])/sc,axis=1),6))
    print("     gen vs oracle per-epoch max", np.round(np.max(np.abs(gen[c].raw-raw[c])/sc,axis=1),6))
    print("     fast run0 vs run1 identical:", np.array_equal(runs[0][c].raw, runs[1][c].raw), np.array_equal(runs[1][c].raw, runs[2][c].raw))
    print("     carrFreq diff fast-oracle", runs[0][c].carrFreq - tr[c].carrFreq)
    print("     remCode diff", runs[0][c].remCodePhase - tr[c].remCodePhase)
    print("     absSample diff", runs[0][c].absoluteSample - tr[c].absoluteSample)

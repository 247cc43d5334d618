import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle'); sys.path.insert(0,'/root/repo/tests')
import util, bds_oracle as O, c_oracle
from bds3_b200 import _lib as L, _track
s, sats, x, ch = util.record("WB", 2, 0.13)
ps = util.product_settings(s)
tr, raw = util.oracle_track("WB", s, x, ch, 10)
fast,_ = _track.run_tracking("WB", x, ch, ps, n_epochs=10, kernel=L.KERNEL_FAST, raw=True)
np.set_printoptions(linewidth=200, precision=3)
for c in (1,):
    g=fast[c]; codes=O.make_track_codes("WB", s, ch[c].PRN)
    for e in range(5,9):
        pos=int(g.absoluteSample[e]); blk=int(g.absoluteSample[e+1]-pos)
        step=g.codeFreq[e]/s.samplingFreq
        out,_,_=c_oracle.correlate_epoch("WB", s, x[pos:pos+blk], codes, g.remCodePhase[e], step, g.carrFreq[e], g.remCarrPhase[e])
        ref=np.array([out.get(k,0.0) for k in util.RAW_NAMES])
        sc=util.family_scale(ref[None,:])[0]
        print("epoch",e,"self-consistency max rel err", np.max(np.abs(g.raw[e]-ref)/sc), " vs closed-loop-oracle", np.max(np.abs(g.raw[e]-raw[c][e])/util.family_scale(raw[c][e][None,:])[0]))
        print("   err/scale", (g.raw[e]-raw[c][e])/util.family_scale(raw[c][e][None,:])[0])
    print("dll fast", g.dllDiscr, "\ndll orc ", tr[c].dllDiscr)
    print("pll fast", g.pllDiscr, "\npll orc ", tr[c].pllDiscr)

import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bds3_b200 as B
from bds3_b200 import _lib as L, _track, synth
FS = 99.375e6
L.init(0)
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
st = B.b1c.initSettings(samplingFreq=FS, numberOfChannels=60, pilotTRKflag=2, msToProcess=int(secs * 1000))
sats = synth.make_sats(60, st, "B1C")
ch = synth.channels_from_sats(sats, st, "B1C", freq_error=2.0)
n = int(round(secs * FS))
ne = max(1, int((n - 993750) // 993760) - 1)
x = torch.empty(n + 64, dtype=torch.int8, device="cuda")
synth.synth_device("B1C", st, sats, n, out_ptr=x.data_ptr())
xh = torch.empty(n, dtype=torch.int8).pin_memory(); xh.copy_(x[:n]); torch.cuda.synchronize()
s = _track.TrackSession("WB", st, ch)
for it in range(8):
    s.reset()
    t0 = time.perf_counter()
    s.run_streamed(xh.data_ptr(), n, ne)
    s.sync()
    dt = time.perf_counter() - t0
    cs, ep, ms = s.stats()
    pl = s.fetch(ne)
    d = pl["epochsDone"]
    print(it, "ms %.1f" % (dt * 1e3), "epochsDone min", d.min(), "max", d.max(), "short channels", list(np.nonzero(d < ne)[0])[:8], flush=True)

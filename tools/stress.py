"""Developer stress test: repeat the 60-channel closed-loop run (resident and streamed) and check that every channel
completes every epoch and that the outputs are identical from run to run."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bds3_b200 as B
from bds3_b200 import _lib as L, _track, synth
FS = 99.375e6
L.init(0)
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 6.0
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100
st = B.b1c.initSettings(samplingFreq=FS, numberOfChannels=60, pilotTRKflag=2, msToProcess=int(secs * 1000))
sats = synth.make_sats(60, st, "B1C")
ch = synth.channels_from_sats(sats, st, "B1C", freq_error=2.0)
n = int(round(secs * FS))
ne = max(1, int((n - 993750) // 993760) - 1)
x = torch.empty(n + 64, dtype=torch.int8, device="cuda")
synth.synth_device("B1C", st, sats, n, out_ptr=x.data_ptr())
xh = torch.empty(n, dtype=torch.int8).pin_memory(); xh.copy_(x[:n]); torch.cuda.synchronize()
sr = _track.TrackSession("WB", st, ch, device_ptr=x.data_ptr(), n_samples=n)
ss = _track.TrackSession("WB", st, ch)
ref = None
bad = 0
for it in range(iters):
    for name, s in (("resident", sr), ("streamed", ss)):
        s.reset()
        if name == "resident":
            s.run_async(ne)
        else:
            s.run_streamed(xh.data_ptr(), n, ne, chunk_bytes=int(os.environ.get("CHUNK", 32 << 20)))
        pl = s.fetch(ne, raw=(os.environ.get('RAW', '1') == '1'))
        d = pl["epochsDone"]
        if d.min() != ne:
            bad += 1
            c = int(np.argmin(d))
            print(it, name, "INCOMPLETE: channel", c, "epochs", int(d[c]), "of", ne, "| carrFreq tail", pl["carrFreq"][c, max(0, d[c] - 2): d[c] + 1],
                  "remCode", pl["remCodePhase"][c, max(0, d[c] - 2): d[c] + 1], "abs", pl["absoluteSample"][c, max(0, d[c] - 1): d[c] + 1], flush=True)
            if "raw" in pl:
                np.set_printoptions(linewidth=250, precision=1, suppress=True)
                for e in range(max(0, d[c] - 4), d[c]):
                    print("   epoch", e, "raw", pl["raw"][c, e], "pll", pl["pllDiscr"][c, e], "dll", pl["dllDiscr"][c, e], "codeFreq", pl["codeFreq"][c, e], flush=True)
        key = pl["I_P"].copy()
        if ref is None:
            ref = key
        elif not np.array_equal(ref, key) and d.min() == ne:
            bad += 1
            w = np.argwhere(ref != key)
            print(it, name, "DIFFERENT OUTPUT at", w[:3].tolist(), flush=True)
print("iterations", iters, "x2, failures", bad)

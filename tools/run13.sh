#!/bin/bash
B="timeout 90 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
run() { echo "== $*"; env $1 $B $2 $3 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('x_rt',round(d['config']['x_realtime'],1),'kernel_ms',round(d['roofline']['kernel_ms_per_launch'],2))
"; }
run BDS_TRK_PASSES=1 --channels 30
run BDS_TRK_PASSES=2 --channels 30
run BDS_TRK_PASSES=1 --channels 15
run BDS_TRK_PASSES=2 --channels 15
run BDS_TRK_AHEAD=1 --channels 15
run BDS_TRK_AHEAD=0 --channels 8
# other tracking modes at 60 channels, 3 s records (general kernel is slow)
python - <<'PY'
import sys, time, os
sys.path.insert(0, os.getcwd())
import torch
import bds3_b200 as B
from bds3_b200 import _lib as L, _track, synth
L.init(0)
for mode, sig, secs, ne, kern in (("NB", "B1C", 3.0, 290, L.KERNEL_AUTO), ("WB", "B1C", 0.6, 50, L.KERNEL_GENERAL), ("B2a", "B2a", 0.3, 280, L.KERNEL_AUTO)):
    st = (B.b2a.initSettings if sig == "B2a" else B.b1c.initSettings)(samplingFreq=99.375e6, numberOfChannels=60)
    if mode == "NB": st.pilotTRKflag = 1
    if mode == "WB": st.pilotTRKflag = 2
    sats = synth.make_sats(60, st, sig, max_doppler=100.0 if sig == "B2a" else 4500.0)
    ch = synth.channels_from_sats(sats, st, sig, freq_error=2.0)
    n = int(secs * 99.375e6)
    x = torch.empty(n + 64, dtype=torch.int8, device="cuda")
    synth.synth_device(sig, st, sats, n, out_ptr=x.data_ptr())
    s = _track.TrackSession(mode, st, ch, kernel=kern, device_ptr=x.data_ptr(), n_samples=n)
    for _ in range(2):
        s.reset(); s.run_async(ne); s.sync()
    cs, ep, ms = s.stats()
    T = ne * st.intTime
    print(f"{mode} 60 ch kernel={'general' if kern == L.KERNEL_GENERAL else 'auto'}: {ep} epochs in {ms:.2f} ms = {T / (ms * 1e-3):.1f} x real time")
    s.close()
PY

"""How far does the device's closed-loop trajectory drift from the float64 oracle's over seconds of signal?

Both loops start from the same channel state and track the same int8 record; every step of the device's loop is the
oracle's to rounding (tests/util.one_step_parity), but the two trajectories are not identical: fp32 accumulation and CUDA's
atan against libm's differ at the 1e-7 level, which now and then moves one sample across a chip edge.  This script runs
the oracle closed loop (C correlator, all host threads) next to the device for N epochs of one B1C wide-band channel and
prints the largest differences.  VERDICT r1, weak 3.   python tools/trajectory_divergence.py [epochs]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import bds3_b200 as B                      # noqa: E402
from bds3_b200 import _lib as L, _track, synth   # noqa: E402
import util                                # noqa: E402


def main():
    n_epochs = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    L.init(0)
    s = util.settings_for("WB", numberOfChannels=2)
    sats = synth.make_sats(2, s, "B1C", seed=11, sigma=25.0, max_doppler=4500.0)
    n = int((n_epochs + 2.2) * 993750)
    x_dev = torch.empty(n + 64, dtype=torch.int8, device="cuda")
    synth.synth_device("B1C", s, sats, n, out_ptr=x_dev.data_ptr())
    torch.cuda.synchronize()
    x = x_dev[:n].cpu().numpy()
    ch = synth.channels_from_sats(sats, s, "B1C", freq_error=2.0)
    got, _ = _track.run_tracking("WB", x, ch, util.product_settings(s), n_epochs=n_epochs, raw=True)
    s1 = s.copy()
    s1.numberOfChannels = 1
    tr, raw = util.oracle_track("WB", s1, x, ch[:1], n_epochs)
    g, o = got[0], tr[0]
    o = {k: np.asarray(o[k][:n_epochs], dtype=np.float64) for k in ("carrFreq", "codeFreq", "remCodePhase", "absoluteSample", "I_P", "Q_P")}
    o = B.Settings(o)
    ip = np.abs(np.asarray(o.I_P)) + np.abs(np.asarray(o.Q_P))
    out = {"epochs": n_epochs, "seconds": n_epochs * 0.01, "channel_prn": int(ch[0].PRN),
           "max_abs_carrFreq_Hz": float(np.max(np.abs(g.carrFreq - o.carrFreq))),
           "max_abs_codeFreq_Hz": float(np.max(np.abs(g.codeFreq - o.codeFreq))),
           "max_abs_code_phase_chips": float(np.max(np.abs(g.remCodePhase - (g.absoluteSample - o.absoluteSample)
                                                             * (o.codeFreq / s.samplingFreq) - o.remCodePhase))),
           "max_abs_absoluteSample": float(np.max(np.abs(g.absoluteSample - o.absoluteSample))),
           "max_rel_prompt": float(np.max(np.hypot(g.I_P - o.I_P, g.Q_P - o.Q_P) / np.hypot(o.I_P, o.Q_P))),
           "first_epoch_rel_err_18_sums": float(np.max(np.abs(g.raw[0] - raw[0][0]) / util.family_scale(raw[0][0][None, :])[0])),
           "epochs_with_identical_absoluteSample": int(np.sum(g.absoluteSample == o.absoluteSample)),
           "last_second_max_abs_carrFreq_Hz": float(np.max(np.abs(g.carrFreq[-100:] - o.carrFreq[-100:]))),
           "last_second_max_rel_prompt": float(np.max((np.hypot(g.I_P - o.I_P, g.Q_P - o.Q_P) / np.hypot(o.I_P, o.Q_P))[-100:]))}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

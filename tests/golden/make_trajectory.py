"""Writes tests/golden/wb_trajectory_100.npz: the float64 oracle's closed-loop trajectory (C correlator) of one B1C
wide-band channel over 100 epochs = 1 s of the deterministic synthetic record util.record("WB", 2, 1.03).  About five
minutes of one host core; the GPU test test_one_second_trajectory_against_the_stored_oracle_trajectory re-creates the
record and compares the device's trajectory with this one.      python tests/golden/make_trajectory.py"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import util  # noqa: E402

N_EPOCHS, SECONDS = 100, 1.03


def main():
    t0 = time.time()
    s, sats, x, ch = util.record("WB", 2, SECONDS)
    print("record", x.size, "samples in", round(time.time() - t0, 1), "s", flush=True)
    s1 = s.copy()
    s1.numberOfChannels = 1
    tr, raw = util.oracle_track("WB", s1, x, ch[:1], N_EPOCHS)
    o = tr[0]
    keys = ("absoluteSample", "carrFreq", "codeFreq", "remCodePhase", "remCarrPhase", "I_P", "Q_P", "I_E", "Q_E", "I_L", "Q_L",
            "dllDiscr", "pllDiscr")
    np.savez_compressed(os.path.join(HERE, "wb_trajectory_100.npz"), prn=int(ch[0].PRN), n_epochs=N_EPOCHS, seconds=SECONDS,
                        x_sha_head=np.frombuffer(x[:1 << 20].tobytes(), dtype=np.uint8).astype(np.uint64).sum(),
                        raw=np.asarray(raw[0]), **{k: np.asarray(o[k][:N_EPOCHS], dtype=np.float64) for k in keys})
    print("done in", round(time.time() - t0, 1), "s")


if __name__ == "__main__":
    main()

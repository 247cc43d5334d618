"""CPU tests of the host-side mirror of the reference interface: settings, preRun, the
trackResults template / assembly (field order, initial values, short-read control flow)."""
import numpy as np
import pytest

import bds_oracle as O
import bds3_b200 as B
from bds3_b200 import _lib as L, _track, _acq, _shard
from bds3_b200.settings import Struct


def test_init_settings_match_reference_defaults():
    s, o = B.b1c.initSettings(), O.initSettings_B1C()
    for k, v in o.items():
        assert s[k] == v, k
    s, o = B.b2a.initSettings(), O.initSettings_B2a()
    for k, v in o.items():
        assert s[k] == v, k
    assert B.b1c.initSettings().acqStep == 50 and B.b2a.initSettings().acqStep == 400


def test_prerun_matches_oracle():
    rng = np.random.default_rng(1)
    for sig, mod, st in (("B1C", B.b1c, O.initSettings_B1C(numberOfChannels=4)),
                         ("B2a", B.b2a, O.initSettings_B2a(numberOfChannels=4))):
        acq = Struct(carrFreq=np.zeros(12), codePhase=np.zeros(12), peakMetric=rng.uniform(1, 20, 12))
        found = [1, 4, 5, 8, 10, 11]
        acq.carrFreq[found] = st.IF + rng.uniform(-4000, 4000, len(found))
        acq.codePhase[found] = rng.integers(1, 99375, len(found))
        got = mod.preRun(acq, B.Settings(dict(st)))
        want = O.preRun(O.Settings(dict(acq)), st, sig)
        assert len(got) == 4
        for a, b in zip(got, want):
            assert (a.PRN, a.acquiredFreq, a.codePhase, a.codeFreq, a.status) == \
                   (b.PRN, b.acquiredFreq, b.codePhase, b.codeFreq, b.status)
    # fewer detections than channels: the rest keep PRN 0 / status '-'
    acq = Struct(carrFreq=np.array([0, 14.58e6, 0]), codePhase=np.array([0, 77, 0]), peakMetric=np.array([9., 3., 5.]))
    ch = B.b1c.preRun(acq, B.b1c.initSettings(numberOfChannels=3))
    assert [c.PRN for c in ch] == [1, 0, 0] or [c.PRN for c in ch][1:] == [0, 0]


@pytest.mark.parametrize("mode,flag", [("WB", 2), ("NB", 1), ("NB", 0), ("B2a", 1)])
def test_template_equals_oracle_template(mode, flag):
    st = (O.initSettings_B2a if mode == "B2a" else O.initSettings_B1C)(pilotTRKflag=flag)
    want = O._new_track_result(mode, st, 120)
    got = _track.template(mode, B.Settings(dict(st)), 120)
    assert list(got.keys()) == list(want.keys())
    for k in want:
        if isinstance(want[k], np.ndarray):
            np.testing.assert_array_equal(got[k], want[k])
        else:
            assert got[k] == want[k]


def _fake_planes(nch, N, nc, done):
    planes = {n: np.arange(nch * N, dtype=float).reshape(nch, N) + i for i, n in enumerate(L.TRK_PLANES)}
    for n in L.CNO_PLANES:
        planes[n] = np.ones((nch, nc))
    planes["epochsDone"] = np.asarray(done, dtype=np.int32)
    return planes


def test_assemble_short_read_control_flow():
    """WB_tracking.m:279-283,485-488: the first starving channel keeps its data with status '-',
    every later channel keeps the template and has no PRN field."""
    st = B.b1c.initSettings(samplingFreq=99.375e6, numberOfChannels=4, CNoInterval=5)
    ch = [Struct(PRN=p, acquiredFreq=1.0, codePhase=1, codeFreq=1.0, status="T") for p in (3, 0, 9, 11)]
    tr = _track.assemble("WB", st, ch, _fake_planes(4, 20, 4, [20, 0, 12, 20]), 20)
    assert tr[0].status == "T" and tr[0].PRN == 3 and tr[0].I_P[5] == L.TRK_PLANES.index("I_P") + 5
    assert tr[1].status == "-" and "PRN" not in tr[1]                # unused channel (PRN 0)
    assert tr[2].status == "-" and tr[2].PRN == 9 and tr[2].epochsDone == 12
    assert np.all(tr[2].DataCNo[:2] == 1) and np.all(tr[2].DataCNo[2:] == 0)
    assert tr[3].status == "-" and "PRN" not in tr[3] and np.all(np.isinf(tr[3].carrFreq))


def test_num_to_process():
    assert _track.num_to_process("WB", B.b1c.initSettings(msToProcess=37000)) == 3700     # WB_tracking.m:56
    assert _track.num_to_process("B2a", B.b2a.initSettings(msToProcess=49000)) == 49000   # tracking.m:100


def test_as_int8_validation():
    np.testing.assert_array_equal(L.as_int8(np.array([1.0, -127.0, 0.0])), np.array([1, -127, 0], dtype=np.int8))
    with pytest.raises(L.BdsError):
        L.as_int8(np.array([0.5]))
    # fileType 2: a complex vector becomes the file's interleaved I, Q byte pairs (postProcessing.m:96-99)
    np.testing.assert_array_equal(L.as_int8(np.array([1 + 2j, -3 - 127j])), np.array([1, 2, -3, -127], dtype=np.int8))
    with pytest.raises(L.BdsError):
        L.as_int8(np.array([1 + 2.5j]))
    assert L.is_iq(np.zeros(2, dtype=complex)) and not L.is_iq(np.zeros(2))
    assert L.is_iq(np.zeros(2, dtype=np.int8), B.Settings(fileType=2))


def test_resampling_settings_reach_the_library():
    """acquisition.m:56-57: the branch is the library's (bds_acq_cfg.resamplingflag / resamplingThreshold)"""
    assert {"resamplingThreshold", "resamplingflag", "fileType"} <= {f[0] for f in L.bds_acq_cfg._fields_}
    s = B.b1c.initSettings(samplingFreq=99.375e6, resamplingflag=1)
    if not L.device_ok():
        with pytest.raises(L.BdsError) as e:     # no CPU fallback: the call reaches the library and fails on the device check
            _acq.acquire(L.SIG_B1C, np.zeros(16, dtype=np.int8), s)
        assert e.value.code == -2


def test_shard_helpers():
    assert _shard.shard_indices(60, 1, 8) == list(range(1, 60, 8))
    assert sum(len(_shard.shard_indices(60, r, 8)) for r in range(8)) == 60
    rngs = [_shard.prn_range(63, r, 8) for r in range(8)]
    assert [hi - lo for lo, hi in rngs] == [8, 8, 8, 8, 8, 8, 8, 7]
    assert rngs[0][0] == 0 and rngs[-1][1] == 63 and all(a[1] == b[0] for a, b in zip(rngs, rngs[1:]))
    with pytest.raises(ValueError):
        _shard.shard_indices(4, 2, 2)


def test_bench_reference_arm_contract_and_no_cpu_fallback():
    """bench.py --impl reference runs the CPU restatement (never the product) and prints the contract's JSON line;
    the product arm refuses to run without a B200 instead of falling back to the CPU."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--channels", "2"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Msamples/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # nothing of the product is loaded in the reference arm: its input is rendered with numpy + the oracle's generators
    assert line["native_so_loaded"] and all(p.startswith("oracle/") for p in line["native_so_loaded"]), line["native_so_loaded"]
    assert "WB tracking" in line["config"]["workload"] and "epochs" in line["config"]["sample"]
    if not B.device_ok():
        r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                           timeout=300)
        assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)

"""CPU tests of the drop-in boundary: libbdsgpu.so loads, exports every symbol include/bdsgpu.h
declares, fails loudly without a B200 (no CPU fallback), and the product never touches oracle/."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import bds3_b200 as B
from bds3_b200 import _lib as L, _track

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "bdsgpu.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bds_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported():
    names = _declared_functions()
    assert len(names) >= 30
    lib = L.lib()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in bdsgpu.h but not exported by libbdsgpu.so"
    assert sorted(L.EXPORTS) == names, "ctypes binding table out of sync with the header"
    assert lib.bds_abi_version() == L.ABI_VERSION == 3


def test_struct_layouts_match_header():
    """ctypes mirrors have the sizes the C compiler gives the header's structs."""
    prog = r'''
#include <stdio.h>
#include "bdsgpu.h"
int main(void){printf("%zu %zu %zu %zu %zu\n",sizeof(bds_acq_cfg),sizeof(bds_trk_cfg),sizeof(bds_channel),sizeof(bds_trk_out),sizeof(bds_sat));return 0;}
'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(prog)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", os.path.join(d, "s")])
        sizes = [int(v) for v in subprocess.check_output([os.path.join(d, "s")]).split()]
    assert sizes == [C.sizeof(L.bds_acq_cfg), C.sizeof(L.bds_trk_cfg), C.sizeof(L.bds_channel),
                     C.sizeof(L.bds_trk_out), C.sizeof(L.bds_sat)]


def test_header_is_plain_c():
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", HEADER])


def test_mex_gateway_typechecks_against_stub():
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-fsyntax-only", "-DMATLAB_MEX_FILE",
                           "-I", os.path.join(ROOT, "tests", "stubs"), "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "matlab", "bds_mex.c")])


@pytest.mark.skipif(B.device_ok(), reason="a B200 is present: the no-device behaviour cannot be observed")
def test_no_cpu_fallback_without_device():
    """Every compute entry point must fail with BDS_ERR_NO_DEVICE when there is no sm_100 device."""
    lib = L.lib()
    assert lib.bds_device_ok() == 0
    assert lib.bds_init(0) == L.ERR_NO_DEVICE
    s = B.b1c.initSettings(samplingFreq=99.375e6)
    x = np.zeros(4096, dtype=np.int8)
    with pytest.raises(L.BdsError) as e:
        B.b1c.acquisition(x, s)
    assert e.value.code == L.ERR_NO_DEVICE
    ch = [B.settings.Struct(PRN=1, acquiredFreq=14.58e6, codePhase=1, codeFreq=1.023e6, status="T")]
    s.numberOfChannels = 1
    with pytest.raises(L.BdsError) as e:
        B.b1c.WB_tracking(x, ch, s, n_epochs=1)
    assert e.value.code == L.ERR_NO_DEVICE
    cfg = _track.make_cfg("WB", s)
    prn = np.array([1], dtype=np.int32)
    nco = np.zeros(6)
    sums = np.zeros(18)
    rc = lib.bds_track_correlate_open_loop(L.TRK_B1C_WB, C.byref(cfg), L.ptr(x), x.size, L.LOC_HOST, L.ptr(prn), 1, 1,
                                           L.ptr(nco), L.ptr(sums))
    assert rc == L.ERR_NO_DEVICE and b"device" in lib.bds_last_error().lower()
    out = np.zeros(16, dtype=np.int8)
    sat = (L.bds_sat * 1)()
    assert lib.bds_synth_if(L.SIG_B1C, 99.375e6, 14.58e6, 1575.42e6, 1.023e6, sat, 1, 25.0, 1, 0, 16, L.ptr(out),
                            L.LOC_HOST) == L.ERR_NO_DEVICE


def test_bad_arguments_are_rejected_before_any_device_work():
    lib = L.lib()
    assert lib.bds_track_run_async(None, 1) == -1
    assert lib.bds_track_fetch(None, None, 0) == -1
    assert lib.bds_track_counters(None, None) == -1


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package (or the MEX/ABI sources) may
    import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "bds-3-b1c-b2a-sdr-receiver_b200")
    bad = []
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                if re.search(r"bds_oracle|c_oracle|liboracle|oracle/", txt):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad
    so = subprocess.check_output(["ldd", L.lib_path()]).decode()
    assert "oracle" not in so


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "_LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(L.BdsError):
        L.lib()

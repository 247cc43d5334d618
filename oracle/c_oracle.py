"""ctypes wrapper of oracle/c/bds_oracle.c (TEST INFRASTRUCTURE ONLY, see that file's header)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_correlate_epoch.restype = None
        _lib.orc_correlate_epoch.argtypes = [C.c_int, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                             C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    return _lib


_MODE = {"WB": 1, "NB": 2, "B2a": 3}
_NAMES = [f"{fam}_{iq}_{epl}" for fam in ("d", "p", "p61") for epl in ("E", "P", "L") for iq in ("I", "Q")]


def correlate_epoch(mode, settings, raw, codes, remCodePhase, codePhaseStep, carrFreq, remCarrPhase):
    """Drop-in for bds_oracle.correlate_epoch (same signature / return)."""
    raw = np.ascontiguousarray(raw, dtype=np.int8)
    i8 = {}
    for k in ("data", "pilot", "pilot61"):
        if k in codes:
            key = "_i8_" + k
            if key not in codes:
                codes[key] = np.ascontiguousarray(codes[k], dtype=np.int8)
            i8[k] = codes[key]
    out = np.zeros(18)
    rc, rp = C.c_double(), C.c_double()
    p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    lib().orc_correlate_epoch(_MODE[mode], p(raw), raw.size, p(i8["data"]), p(i8.get("pilot")), p(i8.get("pilot61")),
                              float(remCodePhase), float(codePhaseStep), float(carrFreq), float(remCarrPhase),
                              float(settings.dllCorrelatorSpacing), float(settings.samplingFreq),
                              float(settings.codeLength), p(out), C.byref(rc), C.byref(rp))
    s = {}
    for i, name in enumerate(_NAMES):
        fam = name.split("_")[0]
        if (fam == "p" and "pilot" not in i8) or (fam == "p61" and "pilot61" not in i8):
            continue
        s[name] = float(out[i])
    return s, rc.value, rp.value

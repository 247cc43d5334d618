"""ctypes binding of libbdsgpu.so (include/bdsgpu.h)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, os.environ.get("BDS_LIB_NAME", "libbdsgpu.so"))  # BDS_LIB_NAME: developer A/B builds


class BdsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libbdsgpu error {code}: {msg}")
        self.code = code


class bds_acq_cfg(C.Structure):
    _fields_ = [("samplingFreq", C.c_double), ("IF", C.c_double), ("codeFreqBasis", C.c_double),
                ("codeLength", C.c_int32), ("acqSearchBand", C.c_double), ("acqStep", C.c_double),
                ("acqThreshold", C.c_double), ("acqCohT", C.c_int32), ("pilotACQflag", C.c_int32),
                ("fineNoncoh", C.c_int32), ("fileType", C.c_int32), ("resamplingThreshold", C.c_double),
                ("resamplingflag", C.c_int32), ("tune", C.c_int32)]


class bds_trk_cfg(C.Structure):
    _fields_ = [("samplingFreq", C.c_double), ("codeFreqBasis", C.c_double), ("codeLength", C.c_int32),
                ("dllCorrelatorSpacing", C.c_double), ("intTime", C.c_double), ("pilotTRKflag", C.c_int32),
                ("CNoInterval", C.c_int32), ("tau1code", C.c_double), ("tau2code", C.c_double),
                ("pf3", C.c_double), ("pf2", C.c_double), ("pf1", C.c_double), ("wbFactor", C.c_double),
                ("kernel", C.c_int32), ("reserved", C.c_int32), ("fwPassesPerTask", C.c_int32),
                ("fwPrefetch", C.c_int32), ("debug", C.c_int32), ("traceTickets", C.c_int32),
                ("lockLossPLD", C.c_double), ("lockLossIntervals", C.c_int32), ("fwMaxCtas", C.c_int32),
                ("fileType", C.c_int32), ("b2aClusterSize", C.c_int32)]


class bds_channel(C.Structure):
    _fields_ = [("PRN", C.c_int32), ("status", C.c_int32), ("acquiredFreq", C.c_double),
                ("codePhase", C.c_double), ("codeFreq", C.c_double)]


_PD = C.POINTER(C.c_double)
TRK_PLANES = ["absoluteSample", "codeFreq", "carrFreq", "I_P", "I_E", "I_L", "Q_E", "Q_P", "Q_L",
              "Pilot_I_P", "Pilot_I_E", "Pilot_I_L", "Pilot_Q_E", "Pilot_Q_P", "Pilot_Q_L",
              "dllDiscr", "dllDiscrFilt", "pllDiscr", "pllDiscrFilt", "remCodePhase", "remCarrPhase"]
CNO_PLANES = ["DataCNo", "DataPLD", "PilotCNo", "PilotPLD", "TotalCNo"]


class bds_trk_out(C.Structure):
    _fields_ = ([(n, _PD) for n in TRK_PLANES] + [(n, _PD) for n in CNO_PLANES] +
                [("raw", _PD), ("epochsDone", C.POINTER(C.c_int32)), ("lockLostEpoch", C.POINTER(C.c_int32))])


class bds_sat(C.Structure):
    _fields_ = [("PRN", C.c_int32), ("reserved", C.c_int32), ("doppler", C.c_double), ("codeDelay", C.c_double),
                ("carrPhase", C.c_double), ("amplitude", C.c_double)]


# constants of bdsgpu.h
SIG_B1C, SIG_B2A = 1, 2
TRK_B1C_WB, TRK_B1C_NB, TRK_B2A = 1, 2, 3
CODE_B1C_DATA_PRIMARY, CODE_B1C_PILOT_PRIMARY, CODE_B1C_DATA_BOC11, CODE_B1C_PILOT_BOC11 = 1, 2, 3, 4
CODE_B1C_PILOT_BOC61, CODE_B2A_DATA, CODE_B2A_PILOT = 5, 6, 7
LOC_HOST, LOC_DEVICE = 0, 1
KERNEL_AUTO, KERNEL_GENERAL, KERNEL_FAST = 0, 1, 2
DBG_TIMING, DBG_TRACE = 1, 2
ABI_VERSION = 3
ERR_NO_DEVICE = -2

# every symbol include/bdsgpu.h declares (tests check that the .so exports all of them)
EXPORTS = ["bds_abi_version", "bds_init", "bds_shutdown", "bds_last_error", "bds_launch_count", "bds_device_ok",
           "bds_gen_code", "bds_make_code_table", "bds_acquire", "bds_track_open", "bds_track_open_file",
           "bds_track_feed", "bds_track_run", "bds_track_run_async", "bds_track_run_streamed", "bds_track_run_window", "bds_track_sync", "bds_track_fetch",
           "bds_track_device_block", "bds_track_stats", "bds_track_counters", "bds_track_dump_trace", "bds_track_reset", "bds_track_close",
           "bds_track_correlate_open_loop", "bds_secondary_code", "bds_frame_sync", "bds_synth_if", "bds_dev_alloc", "bds_dev_free",
           "bds_host_alloc_pinned", "bds_host_free_pinned", "bds_memcpy_h2d", "bds_memcpy_d2h", "bds_dev_sync"]

_lib = None


def lib_path() -> str:
    return _LIB_PATH


def lib():
    """The loaded libbdsgpu.so.  Raises (loudly) if it has not been built: no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise BdsError(-100, f"{_LIB_PATH} not built; run `python __graft_entry__.py build` "
                             "(there is no CPU fallback)")
    L = C.CDLL(_LIB_PATH)
    L.bds_last_error.restype = C.c_char_p
    L.bds_launch_count.restype = C.c_longlong
    i8p, vp = C.POINTER(C.c_int8), C.c_void_p
    L.bds_gen_code.argtypes = [C.c_int, C.c_int, vp, C.c_int]
    L.bds_make_code_table.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, vp, C.c_int]
    L.bds_acquire.argtypes = [C.c_int, vp, C.c_size_t, C.c_int, C.POINTER(bds_acq_cfg), vp, C.c_int, C.c_int,
                              C.c_int, vp, vp, vp, C.c_int, vp]
    L.bds_track_open.argtypes = [C.c_int, C.POINTER(bds_trk_cfg), vp, C.c_size_t, C.c_int, C.c_longlong,
                                 C.POINTER(bds_channel), C.c_int, C.POINTER(vp)]
    L.bds_track_open_file.argtypes = [C.c_int, C.POINTER(bds_trk_cfg), C.c_char_p, C.c_longlong, C.c_longlong,
                                      C.POINTER(bds_channel), C.c_int, C.POINTER(vp)]
    L.bds_track_feed.argtypes = [vp, vp, C.c_size_t, C.c_int, C.c_longlong]
    L.bds_track_run.argtypes = [vp, C.c_int, C.POINTER(bds_trk_out), C.c_int]
    L.bds_track_run_async.argtypes = [vp, C.c_int]
    L.bds_track_run_streamed.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_int]
    L.bds_track_run_window.argtypes = [vp, vp, C.c_size_t, C.c_int]
    L.bds_track_sync.argtypes = [vp]
    L.bds_track_fetch.argtypes = [vp, C.POINTER(bds_trk_out), C.c_int]
    L.bds_track_device_block.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(C.c_int),
                                         C.POINTER(C.c_int)]
    L.bds_track_stats.argtypes = [vp, C.POINTER(C.c_longlong), C.POINTER(C.c_int), C.POINTER(C.c_float)]
    L.bds_track_counters.argtypes = [vp, C.POINTER(C.c_longlong)]
    L.bds_track_dump_trace.argtypes = [vp, C.c_char_p]
    L.bds_track_reset.argtypes = [vp]
    L.bds_track_close.argtypes = [vp]
    L.bds_track_close.restype = None
    L.bds_track_correlate_open_loop.argtypes = [C.c_int, C.POINTER(bds_trk_cfg), vp, C.c_size_t, C.c_int, vp,
                                                C.c_int, C.c_int, vp, vp]
    L.bds_secondary_code.argtypes = [C.c_int, vp]
    L.bds_frame_sync.argtypes = [C.c_int, C.c_int, vp, C.c_int, C.c_int, vp, vp, C.c_int, vp]
    L.bds_synth_if.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.POINTER(bds_sat), C.c_int,
                               C.c_double, C.c_uint64, C.c_longlong, C.c_size_t, vp, C.c_int]
    L.bds_dev_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    L.bds_dev_free.argtypes = [vp]
    L.bds_host_alloc_pinned.argtypes = [C.POINTER(vp), C.c_size_t]
    L.bds_host_free_pinned.argtypes = [vp]
    L.bds_memcpy_h2d.argtypes = [vp, vp, C.c_size_t]
    L.bds_memcpy_d2h.argtypes = [vp, vp, C.c_size_t]
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        raise BdsError(rc, lib().bds_last_error().decode(errors="replace"))


def device_ok() -> bool:
    return bool(lib().bds_device_ok())


def launch_count() -> int:
    return int(lib().bds_launch_count())


def init(device: int = 0):
    check(lib().bds_init(device))


def ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def as_int8(x) -> np.ndarray:
    """IF samples as a contiguous int8 vector (the reference reads 'schar' into double; the values are integers by
    construction, postProcessing.m:94).  A complex vector (fileType 2: longSignal = I + 1i*Q, postProcessing.m:96-99)
    becomes the interleaved I, Q byte pairs of the file it was read from; ``is_iq`` tells the two apart."""
    a = np.asarray(x)
    if np.iscomplexobj(a):
        a = np.ascontiguousarray(a.reshape(-1), dtype=np.complex128).view(np.float64)   # I0, Q0, I1, Q1, ...
    if a.dtype != np.int8:
        r = np.rint(a)
        if np.any(r != a) or np.any(np.abs(r) > 127):
            raise BdsError(-1, "IF samples must be integers in [-127,127] (schar file contents)")
        a = r.astype(np.int8)
    return np.ascontiguousarray(a.reshape(-1))


def is_iq(x, settings=None) -> bool:
    """True if the samples are I/Q pairs: a complex array, or a raw byte record / file with settings.fileType == 2."""
    if isinstance(x, np.ndarray) and np.iscomplexobj(x):
        return True
    return settings is not None and int(settings.get("fileType", 1)) == 2

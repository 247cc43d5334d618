"""Host side of acquisition: ``acqResults = acquisition(longSignal, settings)``."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .settings import Struct


def acquire(signal, longSignal, settings, prn_range=None, device_ptr=None, n_samples=None, return_debug=False, iq=None):
    """Replaces BDS-3_B1C/acquisition.m / GPU_acquisition.m and BDS-3_B2a/acquisition.m.

    Returns acqResults with .carrFreq/.codePhase/.peakMetric, each 1 x max(acqSatelliteList),
    indexed by PRN, zero = not found (acquisition.m:161-165)."""
    sats = [int(p) for p in np.atleast_1d(settings.acqSatelliteList)]
    maxprn = max(sats)
    cfg = L.bds_acq_cfg(samplingFreq=settings.samplingFreq, IF=settings.IF, codeFreqBasis=settings.codeFreqBasis,
                        codeLength=int(settings.codeLength), acqSearchBand=settings.acqSearchBand,
                        acqStep=settings.acqStep, acqThreshold=settings.acqThreshold,
                        acqCohT=int(settings.get("acqCohT", 0)), pilotACQflag=int(settings.get("pilotACQflag", 0)),
                        fineNoncoh=int(settings.get("fineNoncoh", 0)),
                        resamplingThreshold=float(settings.get("resamplingThreshold", 0.0)),
                        resamplingflag=int(settings.get("resamplingflag", 0)),
                        tune=int(settings.get("_tune", 0)))   # test hook, see bdsgpu.h
    # a complex longSignal is the reference's fileType-2 record (postProcessing.m:96-99); a device record says so itself
    if iq is None:
        iq = longSignal is not None and np.iscomplexobj(longSignal)
    cfg.fileType = 2 if iq else 1
    prn = np.asarray(sats, dtype=np.int32)
    lo, hi = (0, prn.size) if prn_range is None else prn_range
    carr, cph, pm = np.zeros(maxprn), np.zeros(maxprn), np.zeros(maxprn)
    dbg = np.zeros((maxprn, 4))
    if device_ptr is not None:
        xp, n, loc, keep = C.c_void_p(device_ptr), int(n_samples), L.LOC_DEVICE, None
    else:
        keep = L.as_int8(longSignal)
        xp, n, loc = L.ptr(keep), keep.size // (2 if iq else 1), L.LOC_HOST
    L.check(L.lib().bds_acquire(signal, xp, n, loc, C.byref(cfg), L.ptr(prn), prn.size, int(lo), int(hi), L.ptr(carr),
                                L.ptr(cph), L.ptr(pm), maxprn, L.ptr(dbg)))
    acq = Struct(carrFreq=carr, codePhase=cph, peakMetric=pm)
    return (acq, dbg) if return_debug else acq


def preRun(acqResults, settings, b1c: bool):
    """BDS-3_B1C/include/preRun.m:44-76 / BDS-3_B2a/include/preRun.m:44-76 (host-side glue)."""
    ch = [Struct(PRN=0, acquiredFreq=0.0, codePhase=0, codeFreq=0.0, status="-")
          for _ in range(int(settings.numberOfChannels))]
    order = np.argsort(-np.asarray(acqResults.peakMetric), kind="stable")
    n = min(int(settings.numberOfChannels), int(np.sum(np.asarray(acqResults.carrFreq) != 0)))
    for ii in range(n):
        p = int(order[ii])
        ch[ii].PRN = p + 1
        ch[ii].acquiredFreq = float(acqResults.carrFreq[p])
        ch[ii].codePhase = int(acqResults.codePhase[p])
        if b1c:
            ch[ii].codeFreq = settings.codeFreqBasis - (ch[ii].acquiredFreq - settings.IF) / settings.carrFreqBasis * settings.codeFreqBasis
        else:
            ch[ii].codeFreq = float(settings.codeFreqBasis)
        ch[ii].status = "T"
    return ch

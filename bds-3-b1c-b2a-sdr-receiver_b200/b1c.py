"""BDS-3_B1C call surface: initSettings / acquisition / GPU_acquisition / preRun / WB_tracking /
NB_tracking / postProcessing with the reference's names and argument order."""
from __future__ import annotations

import numpy as np

from . import _acq, _lib as L, _track
from .codes import generateDataBOC11, generatePilotBOC11, generatePilotBOC61, makeDataTable, makePilotTable  # noqa: F401
from .settings import Settings, samples_per_code


def initSettings(**over) -> Settings:
    """BDS-3_B1C/initSettings.m:48-151 (shipped values)."""
    s = Settings(
        fileName="Set_Jan17_2018_13_53_for_Jimi_ch0.bin", dataType="schar", fileType=1,
        IF=1590e6 - 1575.42e6, samplingFreq=53e6, FEBW=27e6, msToProcess=37000,
        acqSatelliteList=[19, 20], pilotACQflag=1, gpuACQflag=1, pilotTRKflag=2, numberOfChannels=10,
        skipNumberOfBytes=0, codeLength=10230, codeFreqBasis=1.023e6, carrFreqBasis=1575.42e6,
        skipAcquisition=0, acqSearchBand=5000, acqCohT=10, acqStep=1000 / 10 / 2, acqThreshold=7.5,
        resamplingThreshold=15e6, resamplingflag=0, dllDampingRatio=0.7, dllNoiseBandwidth=1,
        dllCorrelatorSpacing=0.06, pllDampingRatio=0.7, pllNoiseBandwidth=12, intTime=0.01,
        navSolPeriod=200, elevationMask=5, useTropCorr=1, plotTracking=1, c=299792458, startOffset=68.802,
        CNoInterval=50)
    s.update(over)
    return s


def acquisition(longSignal, settings, **kw):
    """acqResults = acquisition(longSignal, settings)   (BDS-3_B1C/acquisition.m:1)"""
    return _acq.acquire(L.SIG_B1C, longSignal, settings, **kw)


GPU_acquisition = acquisition   # BDS-3_B1C/GPU_acquisition.m:1 — same contract


def preRun(acqResults, settings):
    return _acq.preRun(acqResults, settings, b1c=True)


def WB_tracking(fid, channel, settings, **kw):
    """[trackResults, channel] = WB_tracking(fid, channel, settings)   (BDS-3_B1C/WB_tracking.m:1)"""
    return _track.run_tracking("WB", fid, channel, settings, **kw)


def NB_tracking(fid, channel, settings, **kw):
    """[trackResults, channel] = NB_tracking(fid, channel, settings)   (BDS-3_B1C/NB_tracking.m:1)"""
    return _track.run_tracking("NB", fid, channel, settings, **kw)


def generate2ndCode(PRN):
    """Secondary = generate2ndCode(PRN)   (BDS-3_B1C/include/generate2ndCode.m:1): 1800 chips, +-1"""
    import ctypes as C
    out = np.zeros(1800, dtype=np.int8)
    L.check(L.lib().bds_secondary_code(int(PRN), out.ctypes.data_as(C.c_void_p)))
    return out.astype(np.float64)


def frameSync(trackResult, settings):
    """The frame-synchronisation correlation of BCNAV1decoding.m:66-91 on the device: (XcorrResult for lags >= 0,
    index) with index = find(abs(XcorrResult) >= 1799.5) (1-based epoch numbers where a secondary-code period starts)."""
    bits = trackResult.Pilot_I_P if settings.pilotTRKflag == 2 else trackResult.Pilot_Q_P
    return _track.frame_sync(L.SIG_B1C, bits, trackResult.PRN)


def postProcessing(settings, acqResults=None):
    """Acquisition -> preRun -> tracking part of BDS-3_B1C/postProcessing.m:61-149 (navigation and
    plots stay in MATLAB).  Returns (acqResults, channel, trackResults)."""
    with open(settings.fileName, "rb") as fid:
        if settings.skipAcquisition == 0 or acqResults is None:
            spc = samples_per_code(settings)
            k = 2 if int(settings.get("fileType", 1)) == 2 else 1          # dataAdaptCoeff, postProcessing.m:67-71
            fid.seek(k * int(settings.skipNumberOfBytes))                  # :79
            data = np.frombuffer(fid.read(k * 20 * spc), dtype=np.int8)    # :94; fileType 2: I, Q byte pairs (:96-99)
            acqResults = acquisition(data, settings, iq=(k == 2))
        if not np.any(acqResults.carrFreq):
            return acqResults, None, []
        channel = preRun(acqResults, settings)
        if settings.pilotTRKflag == 2:
            trackResults, channel = WB_tracking(fid, channel, settings)
        else:
            trackResults, channel = NB_tracking(fid, channel, settings)
    return acqResults, channel, trackResults

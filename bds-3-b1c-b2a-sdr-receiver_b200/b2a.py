"""BDS-3_B2a call surface: initSettings / acquisition / preRun / tracking / postProcessing."""
from __future__ import annotations

import numpy as np

from . import _acq, _lib as L, _track
from .codes import generateB2aDataCode, generateB2aPilotCode, makeB2aDataTable, makeB2aPilotTable  # noqa: F401
from .settings import Settings, samples_per_code


def initSettings(**over) -> Settings:
    """BDS-3_B2a/initSettings.m:44-130 (shipped values)."""
    s = Settings(
        msToProcess=49000, numberOfChannels=12, skipNumberOfBytes=0, fileName="Beidou_B2a_IF_signal.bin",
        dataType="schar", fileType=1, IF=13.55e6, samplingFreq=99.375e6, codeLength=10230,
        codeFreqBasis=10.23e6, skipAcquisition=0, acqSatelliteList=[19, 20], acqSearchBand=5000,
        acqThreshold=1.5, acqStep=400, fineNoncoh=15, resamplingThreshold=50e6, resamplingflag=0,
        dllDampingRatio=0.7, dllNoiseBandwidth=2, dllCorrelatorSpacing=0.5, pllDampingRatio=0.7,
        pllNoiseBandwidth=20, intTime=0.001, pilotTRKflag=1, navSolPeriod=500, elevationMask=5,
        useTropCorr=1, plotTracking=1, c=299792458, startOffset=68.802, CNoInterval=200,
        carrFreqBasis=1176.45e6)
    s.update(over)
    return s


def acquisition(longSignal, settings, **kw):
    """acqResults = acquisition(longSignal, settings)   (BDS-3_B2a/acquisition.m:1)"""
    return _acq.acquire(L.SIG_B2A, longSignal, settings, **kw)


def preRun(acqResults, settings):
    return _acq.preRun(acqResults, settings, b1c=False)


def tracking(fid, channel, settings, **kw):
    """[trackResults, channel] = tracking(fid, channel, settings)   (BDS-3_B2a/tracking.m:1)"""
    return _track.run_tracking("B2a", fid, channel, settings, **kw)


def frameSync(I_P_InputBits):
    """The preamble correlation of BCNAV2decoding.m:69-97 on the device: (tlmXcorrResult for lags >= 0, index) with
    index = find(abs(.) > 115)."""
    return _track.frame_sync(L.SIG_B2A, I_P_InputBits)


def postProcessing(settings, acqResults=None):
    """BDS-3_B2a/postProcessing.m:57-129 up to tracking."""
    with open(settings.fileName, "rb") as fid:
        if settings.skipAcquisition == 0 or acqResults is None:
            spc = samples_per_code(settings)
            k = 2 if int(settings.get("fileType", 1)) == 2 else 1          # dataAdaptCoeff, postProcessing.m:63-67
            fid.seek(k * int(settings.skipNumberOfBytes))
            data = np.frombuffer(fid.read(k * spc * (int(settings.fineNoncoh) + 2)), dtype=np.int8)   # :89-90; I, Q pairs (:92-96)
            acqResults = acquisition(data, settings, iq=(k == 2))
        if not np.any(acqResults.carrFreq):
            return acqResults, None, []
        channel = preRun(acqResults, settings)
        trackResults, channel = tracking(fid, channel, settings)
    return acqResults, channel, trackResults

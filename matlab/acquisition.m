function acqResults = acquisition(longSignal, settings)
%ACQUISITION  Drop-in for BDS-3_B1C/acquisition.m, GPU_acquisition.m and BDS-3_B2a/acquisition.m
%   acqResults = acquisition(longSignal, settings)
% Put this directory ahead of the reference's on the MATLAB path.  The search itself runs in
% libbdsgpu (bds_acquire); the struct layout is the reference's (acquisition.m:161-165):
% carrFreq, codePhase, peakMetric, each 1 x max(acqSatelliteList), zero = not found.
isB2a = settings.codeFreqBasis > 5e6;            % B2a: 10.23 Mcps, B1C: 1.023 Mcps
% the resampling pre-conditioner (acquisition.m:56-123) runs inside the library when settings ask for it
cfg = [settings.samplingFreq, settings.IF, settings.codeFreqBasis, settings.codeLength, ...
       settings.acqSearchBand, settings.acqStep, settings.acqThreshold, 0, 0, 0, ...
       settings.resamplingThreshold, settings.resamplingflag];
if isB2a
    cfg(10) = settings.fineNoncoh;  signal = 2;
else
    cfg(8) = settings.acqCohT;  cfg(9) = settings.pilotACQflag;  signal = 1;
end
[acqResults.carrFreq, acqResults.codePhase, acqResults.peakMetric] = ...
    bds_mex('acquire', signal, int8(longSignal), cfg, double(settings.acqSatelliteList));
end

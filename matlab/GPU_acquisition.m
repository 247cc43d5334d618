function acqResults = GPU_acquisition(longSignal, settings)
%GPU_ACQUISITION  Drop-in for BDS-3_B1C/GPU_acquisition.m (the reference's gpuArray variant, selected by
%   settings.gpuACQflag at postProcessing.m:105-111): same contract as acquisition.m, same library call.
acqResults = acquisition(longSignal, settings);
end

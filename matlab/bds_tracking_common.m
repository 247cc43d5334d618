function [trackResults, channel] = bds_tracking_common(mode, fid, channel, settings)
%BDS_TRACKING_COMMON  Shared body of the WB_tracking / NB_tracking / tracking drop-ins.
% mode: 1 = B1C wide-band (WB_tracking.m), 2 = B1C narrow-band (NB_tracking.m), 3 = B2a (tracking.m)
% Builds trackResults with exactly the reference's field creation order, initial values and
% 1 x N double row vectors (WB_tracking.m:53-112, NB_tracking.m:53-100, B2a tracking.m:48-96).
if mode == 3
    N = settings.msToProcess;                                   % B2a/tracking.m:100
else
    N = round(settings.msToProcess/1000/settings.intTime);      % WB_tracking.m:56
end
hasPilot = (mode == 1 && settings.pilotTRKflag == 2) || (mode ~= 1 && settings.pilotTRKflag == 1);
nC = floor(N/settings.CNoInterval);

t.status = '-';
t.absoluteSample = zeros(1, N);
t.codeFreq = inf(1, N);   t.carrFreq = inf(1, N);
t.I_P = zeros(1, N); t.I_E = zeros(1, N); t.I_L = zeros(1, N);
t.Q_E = zeros(1, N); t.Q_P = zeros(1, N); t.Q_L = zeros(1, N);
if hasPilot && mode == 1
    t.Pilot_I_P = zeros(1, N); t.Pilot_I_E = zeros(1, N); t.Pilot_I_L = zeros(1, N);
    t.Pilot_Q_E = zeros(1, N); t.Pilot_Q_P = zeros(1, N); t.Pilot_Q_L = zeros(1, N);
elseif hasPilot
    t.Pilot_I_P = zeros(1, N); t.Pilot_Q_P = zeros(1, N);
end
t.dllDiscr = inf(1, N); t.dllDiscrFilt = inf(1, N); t.pllDiscr = inf(1, N); t.pllDiscrFilt = inf(1, N);
t.remCodePhase = inf(1, N); t.remCarrPhase = inf(1, N);
t.DataCNo = zeros(1, nC); t.DataPLD = zeros(1, nC);
if hasPilot
    t.PilotCNo = zeros(1, nC); t.PilotPLD = zeros(1, nC);
    if mode == 3, t.B2a_CNo = zeros(1, nC); else, t.B1C_CNo = zeros(1, nC); end
end
trackResults = repmat(t, 1, settings.numberOfChannels);

% loop constants exactly as the reference computes them (host side, passed in)
[tau1, tau2] = calcLoopCoef(settings.dllNoiseBandwidth, settings.dllDampingRatio, 1.0);
[pf3, pf2, pf1] = calcLoopCoefCarr(settings);
factor = 0;
if mode == 1, factor = CalcWeighingFactor(settings); end          % WB_tracking.m:138
cfg = [settings.samplingFreq, settings.codeFreqBasis, settings.codeLength, settings.dllCorrelatorSpacing, ...
       settings.intTime, settings.pilotTRKflag, settings.CNoInterval, tau1, tau2, pf3, pf2, pf1, factor, ...
       settings.fileType];                                        % 2 = I/Q pairs (WB_tracking.m:155-159)
cm = zeros(settings.numberOfChannels, 5);
for c = 1:settings.numberOfChannels
    cm(c, :) = [channel(c).PRN, double(channel(c).status), channel(c).acquiredFreq, ...
                channel(c).codePhase, channel(c).codeFreq];
end
fileName = fopen(fid);                        % the library maps the file itself; fid is left untouched
[planes, cno, done] = bds_mex('track', mode, fileName, settings.skipNumberOfBytes, cfg, cm, N);

names = {'absoluteSample','codeFreq','carrFreq','I_P','I_E','I_L','Q_E','Q_P','Q_L', ...
         'Pilot_I_P','Pilot_I_E','Pilot_I_L','Pilot_Q_E','Pilot_Q_P','Pilot_Q_L', ...
         'dllDiscr','dllDiscrFilt','pllDiscr','pllDiscrFilt','remCodePhase','remCarrPhase'};
cnames = {'DataCNo','DataPLD','PilotCNo','PilotPLD','B1C_CNo'};
if mode == 3, cnames{5} = 'B2a_CNo'; end
for c = 1:settings.numberOfChannels
    if channel(c).PRN == 0, continue; end                        % WB_tracking.m:165
    trackResults(c).PRN = channel(c).PRN;                        % WB_tracking.m:167
    for f = 1:numel(names)
        if isfield(trackResults, names{f}), trackResults(c).(names{f}) = planes(:, f, c).'; end
    end
    k = floor(double(done(c))/settings.CNoInterval);
    for f = 1:numel(cnames)
        if isfield(trackResults, cnames{f}), trackResults(c).(cnames{f})(1:k) = cno(1:k, f, c).'; end
    end
    if done(c) < N
        return;                         % short read: the reference returns here (WB_tracking.m:279-283)
    end
    trackResults(c).status = channel(c).status;                  % WB_tracking.m:485-488
end
end

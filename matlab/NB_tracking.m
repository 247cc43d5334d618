function [trackResults, channel] = NB_tracking(fid, channel, settings)
%NB_TRACKING  Drop-in for the reference's NB_tracking.m: same signature, same trackResults layout.
[trackResults, channel] = bds_tracking_common(2, fid, channel, settings);
end

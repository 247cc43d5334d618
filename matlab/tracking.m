function [trackResults, channel] = tracking(fid, channel, settings)
%TRACKING  Drop-in for the reference's tracking.m: same signature, same trackResults layout.
[trackResults, channel] = bds_tracking_common(3, fid, channel, settings);
end

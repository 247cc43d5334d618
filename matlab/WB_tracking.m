function [trackResults, channel] = WB_tracking(fid, channel, settings)
%WB_TRACKING  Drop-in for the reference's WB_tracking.m: same signature, same trackResults layout.
[trackResults, channel] = bds_tracking_common(1, fid, channel, settings);
end

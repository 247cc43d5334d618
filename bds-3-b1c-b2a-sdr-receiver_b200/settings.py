"""The flat ``settings`` struct of the reference (initSettings.m), as a dict with attribute access."""
from __future__ import annotations


class Settings(dict):
    __getattr__ = dict.__getitem__

    def __setattr__(self, k, v):
        self[k] = v

    def copy(self):
        return Settings(dict.copy(self))


class Struct(dict):
    """Generic MATLAB-struct stand-in that keeps field creation order (dict order)."""
    __getattr__ = dict.__getitem__

    def __setattr__(self, k, v):
        self[k] = v

    def fieldnames(self):
        return list(self.keys())


def matlab_round(x: float) -> int:
    import math
    return int(math.floor(abs(x) + 0.5)) * (1 if x >= 0 else -1)


def samples_per_code(settings) -> int:
    """round(samplingFreq / (codeFreqBasis / codeLength))   B1C/acquisition.m:129-130"""
    return matlab_round(settings.samplingFreq / (settings.codeFreqBasis / settings.codeLength))

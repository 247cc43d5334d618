"""CPU oracle (numpy float64) for the BDS-3 B1C/B2a acquisition + tracking correlator path.

TEST INFRASTRUCTURE ONLY.  Nothing in the shipped product path may import this
module: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker /
the timed CPU baseline.

PARITY PINNING: "parity unpinned by reference goldens".  The reference
(lyf8118/BDS-3-B1C-B2a-SDR-receiver, pure MATLAB) ships no tests, fixtures or
golden vectors, and neither MATLAB nor Octave exists in this image, so the
reference cannot be executed here.  This file is a vector-for-vector
restatement of the cited MATLAB lines (paths relative to
``/root/reference/BDS3_B1C_B2a``; ``B1C/`` = ``BDS-3_B1C/``, ``B2a/`` =
``BDS-3_B2a/``), pinned by (1) structural known-answer properties of the codes
(Legendre/Weil construction, balance, autocorrelation), (2) an independent C
restatement (``oracle/c/bds_oracle.c``) that must agree bit-exactly on codes
and to 1e-9 on sums, and (3) closed-loop behaviour on synthetic IF (injected
PRN / Doppler / code phase are recovered, loops lock).  See DESIGN.md.

Float64 everywhere, same IEEE expression order as the MATLAB wherever a value
feeds a ``ceil`` or a loop state.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

# ----------------------------------------------------------------------------
# ICD constants carried by the reference as literal tables
# ----------------------------------------------------------------------------
# B1C Weil-code parameters (phase difference w, truncation point p), PRN 1..63.
# B1C/include/generateDataBOC11.m:43-58 (data), generatePilotBOC11.m:44-59 and
# generatePilotBOC61.m:44-59 (pilot).
B1C_DATA_W = [2678, 4802, 958, 859, 3843, 2232, 124, 4352, 1816, 1126, 1860, 4800, 2267, 424, 4192, 4333, 2656, 4148, 243, 1330, 1593, 1470, 882, 3202, 5095, 2546, 1733, 4795, 4577, 1627, 3638, 2553, 3646, 1087, 1843, 216, 2245, 726, 1966, 670, 4130, 53, 4830, 182, 2181, 2006, 1080, 2288, 2027, 271, 915, 497, 139, 3693, 2054, 4342, 3342, 2592, 1007, 310, 4203, 455, 4318]
B1C_DATA_P = [699, 694, 7318, 2127, 715, 6682, 7850, 5495, 1162, 7682, 6792, 9973, 6596, 2092, 19, 10151, 6297, 5766, 2359, 7136, 1706, 2128, 6827, 693, 9729, 1620, 6805, 534, 712, 1929, 5355, 6139, 6339, 1470, 6867, 7851, 1162, 7659, 1156, 2672, 6043, 2862, 180, 2663, 6940, 1645, 1582, 951, 6878, 7701, 1823, 2391, 2606, 822, 6403, 239, 442, 6769, 2560, 2502, 5072, 7268, 341]
B1C_PILOT_W = [796, 156, 4198, 3941, 1374, 1338, 1833, 2521, 3175, 168, 2715, 4408, 3160, 2796, 459, 3594, 4813, 586, 1428, 2371, 2285, 3377, 4965, 3779, 4547, 1646, 1430, 607, 2118, 4709, 1149, 3283, 2473, 1006, 3670, 1817, 771, 2173, 740, 1433, 2458, 3459, 2155, 1205, 413, 874, 2463, 1106, 1590, 3873, 4026, 4272, 3556, 128, 1200, 130, 4494, 1871, 3073, 4386, 4098, 1923, 1176]
B1C_PILOT_P = [7575, 2369, 5688, 539, 2270, 7306, 6457, 6254, 5644, 7119, 1402, 5557, 5764, 1073, 7001, 5910, 10060, 2710, 1546, 6887, 1883, 5613, 5062, 1038, 10170, 6484, 1718, 2535, 1158, 526, 7331, 5844, 6423, 6968, 1280, 1838, 1989, 6468, 2091, 1581, 1453, 6252, 7122, 7711, 7216, 2113, 1095, 1628, 1713, 6102, 6123, 6070, 1115, 8047, 6795, 2575, 53, 1729, 6388, 682, 5565, 7160, 2277]
# B2a register-2 initial states, PRN 1..63; bit k (0-based) = stage k+1 logic
# level.  B2a/include/generateB2aDataCode.m:39-101, generateB2aPilotCode.m:39-101.
B2A_DATA_G2 = [0x1481, 0x581, 0x16a1, 0x1e51, 0x1551, 0xeb1, 0xef1, 0x1bf1, 0x1299, 0xb79, 0x1585, 0x445, 0x1545, 0x1b45, 0x745, 0x18a5, 0x1de5, 0x1015, 0xf95, 0x1ab5, 0x11b5, 0x194d, 0x8cd, 0x32d, 0xdad, 0x9ed, 0x1fed, 0x91d, 0x79d, 0x10bd, 0x27d, 0x57d, 0x1afd, 0x19fd, 0x1143, 0x523, 0x1da3, 0x1113, 0x1313, 0x1ab3, 0x11b3, 0x973, 0x154b, 0x5cb, 0x1a6b, 0x1d5b, 0x587, 0x1827, 0x1a27, 0x18a7, 0x2a7, 0x1b97, 0x1d37, 0x24f, 0x52f, 0x132f, 0xb6f, 0x3ef, 0x1fef, 0x15bf, 0x804, 0x15fb, 0x978]
B2A_PILOT_G2 = B2A_DATA_G2[:60] + [0xc25, 0x3f4, 0x1558]

WEIL_N = 10243


# ----------------------------------------------------------------------------
# a7  B1C code generation
# ----------------------------------------------------------------------------
def legendre_sequence() -> np.ndarray:
    """L(0)=0, L(k)=1 iff k is a quadratic residue mod 10243, else 0.

    B1C/include/generateDataBOC11.m:61-68 builds this with the recursive
    ``JacobiSymbol`` (JacobiSymbol.m:48-123) and maps -1 -> 0; for prime
    modulus the Jacobi symbol is the Legendre symbol, evaluated here by
    enumerating the squares.
    """
    leg = np.zeros(WEIL_N, dtype=np.int8)
    k = np.arange(1, WEIL_N, dtype=np.int64)
    leg[(k * k) % WEIL_N] = 1
    return leg


_LEG = None


def _weil_primary(w: int, p: int) -> np.ndarray:
    """Bipolar 10230-chip Weil code (generateDataBOC11.m:70-82)."""
    global _LEG
    if _LEG is None:
        _LEG = legendre_sequence()
    ind = np.arange(10230, dtype=np.int64)
    k = (ind + p - 1) % WEIL_N
    chip = _LEG[k] ^ _LEG[(k + w) % WEIL_N]
    return (1 - 2 * chip.astype(np.int64)).astype(np.float64)


def b1c_data_primary(prn: int) -> np.ndarray:
    return _weil_primary(B1C_DATA_W[prn - 1], B1C_DATA_P[prn - 1])


def b1c_pilot_primary(prn: int) -> np.ndarray:
    return _weil_primary(B1C_PILOT_W[prn - 1], B1C_PILOT_P[prn - 1])


def _boc11(primary: np.ndarray) -> np.ndarray:
    """chip c -> [-c, +c]  (generateDataBOC11.m:85-90)."""
    out = np.empty(primary.size * 2)
    out[0::2] = -primary
    out[1::2] = primary
    return out


def generateDataBOC11(settings, PRN: int) -> np.ndarray:
    """B1C/include/generateDataBOC11.m:1-90 -> 1x20460 of +-1."""
    return _boc11(b1c_data_primary(PRN))


def generatePilotBOC11(settings, PRN: int) -> np.ndarray:
    """B1C/include/generatePilotBOC11.m:1-94 -> 1x20460 of +-1."""
    return _boc11(b1c_pilot_primary(PRN))


def generatePilotBOC61(settings, PRN: int) -> np.ndarray:
    """B1C/include/generatePilotBOC61.m:89-96: chip c -> (-1)^ii * c, ii=1..12."""
    prim = b1c_pilot_primary(PRN)
    sub = np.array([(-1.0) ** ii for ii in range(1, 13)])
    return (prim[:, None] * sub[None, :]).reshape(-1)


# ----------------------------------------------------------------------------
# a8  B2a code generation
# ----------------------------------------------------------------------------
def _b2a_code(g2_init: int, taps1, taps2) -> np.ndarray:
    """Two 13-stage registers in +-1 form (B2a/include/generateB2aDataCode.m:104-138).

    out = r1(13)*r2(13); feedback = product of tapped stages; shift towards
    stage 13 and insert the feedback at stage 1; r1 is reset to all -1 after
    chip 8190.
    """
    r1 = [-1] * 13
    r2 = [1 - 2 * ((g2_init >> k) & 1) for k in range(13)]
    out = np.empty(10230)
    for ind in range(1, 10231):
        out[ind - 1] = r1[12] * r2[12]
        f1 = 1
        for t in taps1:
            f1 *= r1[t - 1]
        r1 = [f1] + r1[:12]
        f2 = 1
        for t in taps2:
            f2 *= r2[t - 1]
        r2 = [f2] + r2[:12]
        if ind == 8190:
            r1 = [-1] * 13
    return out


_B2A_CACHE: dict = {}


def generateB2aDataCode(PRN: int, settings=None) -> np.ndarray:
    """B2a/include/generateB2aDataCode.m (taps :108-109)."""
    key = ("d", PRN)
    if key not in _B2A_CACHE:
        _B2A_CACHE[key] = _b2a_code(B2A_DATA_G2[PRN - 1], (1, 5, 11, 13), (3, 5, 9, 11, 12, 13))
    return _B2A_CACHE[key].copy()


def generateB2aPilotCode(PRN: int, settings=None) -> np.ndarray:
    """B2a/include/generateB2aPilotCode.m (taps :108-109)."""
    key = ("p", PRN)
    if key not in _B2A_CACHE:
        _B2A_CACHE[key] = _b2a_code(B2A_PILOT_G2[PRN - 1], (3, 6, 7, 13), (1, 5, 7, 8, 12, 13))
    return _B2A_CACHE[key].copy()


# ----------------------------------------------------------------------------
# settings (a16): flat attribute bag mirroring initSettings()
# ----------------------------------------------------------------------------
class Settings(dict):
    """dict with attribute access; mirrors the MATLAB settings struct."""

    __getattr__ = dict.__getitem__

    def __setattr__(self, k, v):
        self[k] = v

    def copy(self):
        return Settings(dict.copy(self))


def initSettings_B1C(**over) -> Settings:
    """B1C/initSettings.m:48-151 (values as shipped)."""
    s = Settings(
        fileName="Set_Jan17_2018_13_53_for_Jimi_ch0.bin", dataType="schar", fileType=1,
        IF=1590e6 - 1575.42e6, samplingFreq=53e6, FEBW=27e6, msToProcess=37000,
        acqSatelliteList=[19, 20], pilotACQflag=1, gpuACQflag=1, pilotTRKflag=2,
        numberOfChannels=10, skipNumberOfBytes=0, codeLength=10230, codeFreqBasis=1.023e6,
        carrFreqBasis=1575.42e6, skipAcquisition=0, acqSearchBand=5000, acqCohT=10,
        acqStep=1000 / 10 / 2, acqThreshold=7.5, resamplingThreshold=15e6, resamplingflag=0,
        dllDampingRatio=0.7, dllNoiseBandwidth=1, dllCorrelatorSpacing=0.06,
        pllDampingRatio=0.7, pllNoiseBandwidth=12, intTime=0.01, CNoInterval=50,
    )
    s.update(over)
    return s


def initSettings_B2a(**over) -> Settings:
    """B2a/initSettings.m:44-130 (values as shipped)."""
    s = Settings(
        msToProcess=49000, numberOfChannels=12, skipNumberOfBytes=0,
        fileName="Beidou_B2a_IF_signal.bin", dataType="schar", fileType=1, IF=13.55e6,
        samplingFreq=99.375e6, codeLength=10230, codeFreqBasis=10.23e6, skipAcquisition=0,
        acqSatelliteList=[19, 20], acqSearchBand=5000, acqThreshold=1.5, acqStep=400,
        fineNoncoh=15, resamplingThreshold=50e6, resamplingflag=0, dllDampingRatio=0.7,
        dllNoiseBandwidth=2, dllCorrelatorSpacing=0.5, pllDampingRatio=0.7,
        pllNoiseBandwidth=20, intTime=0.001, pilotTRKflag=1, CNoInterval=200,
        carrFreqBasis=1176.45e6,
    )
    s.update(over)
    return s


def matlab_round(x: float) -> int:
    """MATLAB round(): half away from zero."""
    return int(math.floor(abs(x) + 0.5)) * (1 if x >= 0 else -1)


def samples_per_code(settings) -> int:
    """round(fs / (codeFreqBasis/codeLength))  (B1C/acquisition.m:129-130)."""
    return matlab_round(settings.samplingFreq / (settings.codeFreqBasis / settings.codeLength))


# ----------------------------------------------------------------------------
# a6  code sampling
# ----------------------------------------------------------------------------
def _b1c_table(code: np.ndarray, settings) -> np.ndarray:
    """B1C/include/makeDataTable.m:45-68 / makePilotTable.m:45-69."""
    spc = samples_per_code(settings)
    ts = 1 / settings.samplingFreq
    tc = 1 / settings.codeFreqBasis / 2
    idx = np.ceil((ts * np.arange(1, spc + 1, dtype=np.float64)) / tc).astype(np.int64)
    idx[-1] = settings.codeLength * 2
    idx[0] = 1
    return code[idx - 1]


def makeDataTable(settings, PRN):
    return _b1c_table(generateDataBOC11(settings, PRN), settings)


def makePilotTable(settings, PRN):
    return _b1c_table(generatePilotBOC11(settings, PRN), settings)


def _b2a_table(code: np.ndarray, settings) -> np.ndarray:
    """B2a/include/makeB2aDataTable.m:42-67 / makeB2aPilotTable.m:42-68."""
    spc = samples_per_code(settings)
    ts = 1 / settings.samplingFreq
    tc = 1 / settings.codeFreqBasis
    idx = np.ceil((ts * np.arange(1, spc + 1, dtype=np.float64)) / tc).astype(np.int64)
    idx[-1] = settings.codeLength
    return code[idx - 1]


def makeB2aDataTable(PRN, settings):
    return _b2a_table(generateB2aDataCode(PRN, settings), settings)


def makeB2aPilotTable(PRN, settings):
    return _b2a_table(generateB2aPilotCode(PRN, settings), settings)


# ----------------------------------------------------------------------------
# resampling pre-conditioner (B1C/acquisition.m:56-123, B2a/acquisition.m:56-124)
# ----------------------------------------------------------------------------
def fir1_bandpass(order: int, wp) -> np.ndarray:
    """b = fir1(order, [w1 w2]): MATLAB's window-method band-pass (fir1.m): the least-squares fit of the ideal band
    (firls with contiguous bands = the truncated ideal impulse response), times hamming(order+1), scaled to unit
    magnitude at the centre of the pass band (b / abs(exp(-j*2*pi*(0:L-1)*(f0/2))*b.')).  Third-party arithmetic
    (Signal Processing Toolbox), restated from its published algorithm - parity unpinned."""
    w1, w2 = float(wp[0]), float(wp[1])
    L = order + 1
    m = np.arange(L, dtype=np.float64) - order / 2
    ideal = w2 * np.sinc(w2 * m) - w1 * np.sinc(w1 * m)
    win = 0.54 - 0.46 * np.cos(2 * np.pi * np.arange(L) / order)
    b = ideal * win
    f0 = (w1 + w2) / 2
    return b / abs(np.sum(np.exp(-1j * 2 * np.pi * np.arange(L) * (f0 / 2)) * b))


def filtfilt_fir(b: np.ndarray, x: np.ndarray) -> np.ndarray:
    """y = filtfilt(b, 1, x) as filtfilt.m does it: nfact = 3*(numel(b)-1); odd reflection of nfact samples about both
    end points; steady-state initial conditions zi scaled by the first sample of each pass; filter forward, reverse,
    filter again, reverse; strip the reflected margins."""
    from scipy.signal import lfilter
    nb = b.size
    nfact = 3 * (nb - 1)
    if x.size <= nfact:
        raise ValueError("filtfilt: data must have length more than 3 times filter order")
    # zi: filter(b, 1, ones) starts in its steady state.  With a = 1: zi(k) = sum(b(k+1:end))
    zi = np.cumsum(b[::-1])[::-1][1:]
    xt = np.concatenate([2 * x[0] - x[nfact:0:-1], x, 2 * x[-1] - x[-2:-nfact - 2:-1]])
    y, _ = lfilter(b, [1.0], xt, zi=zi * xt[0])
    y = y[::-1]
    y, _ = lfilter(b, [1.0], y, zi=zi * y[0])
    return y[::-1][nfact:-nfact]


def resample_for_acquisition(longSignal, settings, BW):
    """acquisition.m:56-123.  Returns (longSignal, settings', oldFreq, oldIF), or the inputs and None when the branch is off."""
    if not (settings.samplingFreq > settings.resamplingThreshold and settings.get("resamplingflag", 0) == 1):
        return longSignal, settings, None, None
    fs, IF = settings.samplingFreq, settings.IF
    w1, w2 = IF - BW / 2, IF + BW / 2
    wp = [w1 * 2 / fs - 0.002, w2 * 2 / fs + 0.002]
    b = fir1_bandpass(700, wp)
    x = np.asarray(longSignal)
    x = x.astype(np.complex128 if np.iscomplexobj(x) else np.float64)
    x = filtfilt_fir(b, x.real) + 1j * filtfilt_fir(b, x.imag) if np.iscomplexobj(x) else filtfilt_fir(b, x)
    fu = IF + BW / 2
    n = max(1, math.floor(fu / BW))
    lowerFreq = 2 * fu / n
    fl = IF - BW / 2
    upperFreq = 2 * fl / (n - 1) if n > 1 else lowerFreq
    st = Settings(dict(settings))
    st.samplingFreq = float(math.ceil((lowerFreq + upperFreq) / 2))
    signalLen = int(math.floor((x.size - 1) / fs * st.samplingFreq))
    index = np.ceil(np.arange(signalLen, dtype=np.float64) / st.samplingFreq * fs).astype(np.int64)
    index[0] = 1
    x = x[index - 1]
    st.IF = math.fmod(IF, st.samplingFreq)
    return x, st, fs, IF


def _resampling_recovery(acq, PRN, codePhase, st, oldFreq, oldIF):
    """acquisition.m:321-338: results at the original sampling rate."""
    acq.codePhase[PRN - 1] = math.floor((codePhase - 1) / st.samplingFreq * oldFreq) + 1
    if st.IF >= st.samplingFreq / 2:
        doppler = (st.samplingFreq - st.IF) - acq.carrFreq[PRN - 1]
    else:
        doppler = acq.carrFreq[PRN - 1] - st.IF
    acq.carrFreq[PRN - 1] = doppler + oldIF


# ----------------------------------------------------------------------------
# a1/a2  B1C acquisition
# ----------------------------------------------------------------------------
def acquisition_B1C(longSignal: np.ndarray, settings, return_debug: bool = False):
    """B1C/acquisition.m:56-338."""
    longSignal, settings, oldFreq, oldIF = resample_for_acquisition(longSignal, settings, 9e6)   # :56-123
    longSignal = np.asarray(longSignal)
    if not np.iscomplexobj(longSignal):
        longSignal = longSignal.astype(np.float64)
    fs = settings.samplingFreq
    spc = samples_per_code(settings)
    M = matlab_round(spc / 10 * settings.acqCohT)            # samplesXmsLen :132
    N = matlab_round(spc / 10 * (10 + settings.acqCohT))     # len10PlusXms  :135
    sig = longSignal[:N]
    ts = 1 / fs
    phasePoints = np.arange(N, dtype=np.float64) * 2 * np.pi * ts
    nbins = matlab_round(settings.acqSearchBand * 2 / settings.acqStep) + 1
    sigPower = math.sqrt(np.var(sig[:M], ddof=1) * M)        # :150
    maxprn = max(settings.acqSatelliteList)
    acq = Settings(carrFreq=np.zeros(maxprn), codePhase=np.zeros(maxprn), peakMetric=np.zeros(maxprn))
    dbg = {}
    frqBins = settings.IF - settings.acqSearchBand + settings.acqStep * np.arange(nbins)
    # the forward FFT of the mixed signal does not depend on the PRN; the
    # reference recomputes it per PRN (:194-205) with identical results.
    IQ = [None] * nbins
    for PRN in settings.acqSatelliteList:
        DataPriTable = makeDataTable(settings, PRN)
        localData = np.concatenate([DataPriTable[:M], np.zeros(N - M)])
        DataPriFreqDom = np.conj(np.fft.fft(localData))
        if settings.pilotACQflag == 1:
            PilotPriTable = makePilotTable(settings, PRN)
            localPilot = np.concatenate([PilotPriTable[:M], np.zeros(N - M)])
            PilotPriFreqDom = np.conj(np.fft.fft(localPilot))
        rowmax = np.zeros(nbins)
        colmax = np.zeros(N)
        for b in range(nbins):
            if IQ[b] is None:
                sigCarr = np.exp(1j * frqBins[b] * phasePoints)
                IQ[b] = np.fft.fft(sigCarr * sig)
            res = np.abs(np.fft.ifft(IQ[b] * DataPriFreqDom))
            if settings.pilotACQflag == 1:
                res = (res * math.sqrt(11) + np.abs(np.fft.ifft(IQ[b] * PilotPriFreqDom)) * math.sqrt(29)) / math.sqrt(40)
            rowmax[b] = res.max()
            np.maximum(colmax, res, out=colmax)
        frequencyBinIndex = int(np.argmax(rowmax))            # first index wins :229
        codePhase = int(np.argmax(colmax)) + 1                # 1-based :232
        peakSize = float(colmax[codePhase - 1])
        acq.peakMetric[PRN - 1] = peakSize / sigPower
        if codePhase + spc - 1 > longSignal.size:            # :239-241
            codePhase -= spc
        dbg[PRN] = dict(bin=frequencyBinIndex, codePhase=codePhase, peak=peakSize, sigPower=sigPower)
        if peakSize / sigPower > settings.acqThreshold:
            s0 = longSignal[codePhase - 1: codePhase - 1 + spc]
            s0 = s0 - s0.mean()
            xCarrier = s0 * DataPriTable
            if settings.pilotACQflag == 1:
                xCarrierPilot = s0 * PilotPriTable
            fineStep = 25
            nfine = matlab_round(settings.acqStep / 25) * 2 + 1
            finePhasePoints = np.arange(spc, dtype=np.float64) * 2 * np.pi * ts
            FineFrq = np.zeros(nfine)
            FineRes = np.zeros(nfine)
            for j in range(nfine):
                FineFrq[j] = frqBins[frequencyBinIndex] - settings.acqStep + fineStep * j
                c = np.exp(1j * FineFrq[j] * finePhasePoints)
                FineRes[j] = abs(np.sum(xCarrier * c))
                if settings.pilotACQflag == 1:
                    FineRes[j] = (FineRes[j] * 11 + abs(np.sum(xCarrierPilot * c)) * 29) / 40
            mx = int(np.argmax(FineRes))
            acq.carrFreq[PRN - 1] = FineFrq[mx]
            if acq.carrFreq[PRN - 1] == 0:
                acq.carrFreq[PRN - 1] = 1
            acq.codePhase[PRN - 1] = codePhase
            dbg[PRN]["fine"] = FineRes
            if oldFreq is not None:
                _resampling_recovery(acq, PRN, codePhase, settings, oldFreq, oldIF)              # :321-338
    return (acq, dbg) if return_debug else acq


# ----------------------------------------------------------------------------
# a4/a5  B2a acquisition
# ----------------------------------------------------------------------------
def acquisition_B2a(longSignal: np.ndarray, settings, return_debug: bool = False):
    """B2a/acquisition.m:56-365."""
    longSignal, settings, oldFreq, oldIF = resample_for_acquisition(longSignal, settings,
                                                                    settings.codeFreqBasis * 2 + 0.5e6)   # :56-124
    longSignal = np.asarray(longSignal)
    if not np.iscomplexobj(longSignal):
        longSignal = longSignal.astype(np.float64)
    fs = settings.samplingFreq
    spc = samples_per_code(settings)
    len2ms = spc * 2
    s2cc = int(math.ceil(fs / settings.codeFreqBasis)) * 2     # samples2CodeChip :137
    sig = longSignal[:len2ms]
    ts = 1 / fs
    phasePoints = np.arange(len2ms, dtype=np.float64) * 2 * np.pi * ts
    nbins = matlab_round(settings.acqSearchBand * 2 / settings.acqStep) + 1
    maxprn = max(settings.acqSatelliteList)
    acq = Settings(carrFreq=np.zeros(maxprn), codePhase=np.zeros(maxprn), peakMetric=np.zeros(maxprn))
    dbg = {}
    frqBins = settings.IF - settings.acqSearchBand + settings.acqStep * np.arange(nbins)
    IQ = [None] * nbins
    for PRN in settings.acqSatelliteList:
        dT = makeB2aDataTable(PRN, settings)
        pT = makeB2aPilotTable(PRN, settings)
        dF = np.conj(np.fft.fft(np.concatenate([dT, np.zeros(spc)])))
        pF = np.conj(np.fft.fft(np.concatenate([pT, np.zeros(spc)])))
        results = np.zeros((nbins, len2ms))
        for b in range(nbins):
            if IQ[b] is None:
                IQ[b] = np.fft.fft(np.exp(1j * frqBins[b] * phasePoints) * sig)
            results[b] = np.abs(np.fft.ifft(IQ[b] * dF)) + np.abs(np.fft.ifft(IQ[b] * pF))
        frequencyBinIndex = int(np.argmax(results.max(axis=1)))
        colmax = results.max(axis=0)
        codePhase = int(np.argmax(colmax)) + 1
        peakSize = float(colmax[codePhase - 1])
        e1 = codePhase - s2cc
        e2 = codePhase + s2cc
        e3 = codePhase - spc + s2cc
        e4 = codePhase + spc - s2cc
        rng = []
        if e1 >= 1:
            rng.append(np.arange(max(1, e3), e1 + 1))
        if e2 < len2ms:
            rng.append(np.arange(e2, min(e4, len2ms) + 1))
        rng = np.concatenate(rng) if rng else np.zeros(0, dtype=np.int64)
        secondPeakSize = float(results[frequencyBinIndex, rng - 1].max())
        acq.peakMetric[PRN - 1] = peakSize / secondPeakSize
        dbg[PRN] = dict(bin=frequencyBinIndex, codePhase=codePhase, peak=peakSize, second=secondPeakSize)
        if peakSize / secondPeakSize > settings.acqThreshold:
            nfine = matlab_round(settings.acqStep / 25) + 1
            dC = generateB2aDataCode(PRN, settings)
            pC = generateB2aPilotCode(PRN, settings)
            K = settings.fineNoncoh * spc
            cvi = np.floor((ts * np.arange(1, K + 1, dtype=np.float64)) / (1 / settings.codeFreqBasis)).astype(np.int64)
            longD = dC[cvi % settings.codeLength]
            longP = pC[cvi % settings.codeLength]
            finePhasePoints = np.arange(K, dtype=np.float64) * 2 * np.pi * ts
            sigFine = longSignal[codePhase - 1: codePhase - 1 + K]
            FineFrq = np.zeros(nfine)
            FineRes = np.zeros(nfine)
            for j in range(nfine):
                FineFrq[j] = frqBins[frequencyBinIndex] - settings.acqStep / 2 + 25 * j
                c = np.exp(1j * FineFrq[j] * finePhasePoints)
                b1 = (longD * c * sigFine).reshape(settings.fineNoncoh, spc).sum(axis=1)
                b2 = (longP * c * sigFine).reshape(settings.fineNoncoh, spc).sum(axis=1)
                FineRes[j] = np.abs(b1).sum() + np.abs(b2).sum()
            mx = int(np.argmax(FineRes))
            acq.carrFreq[PRN - 1] = FineFrq[mx]
            acq.codePhase[PRN - 1] = codePhase
            if acq.carrFreq[PRN - 1] == 0:
                acq.carrFreq[PRN - 1] = 1
            if oldFreq is not None:
                _resampling_recovery(acq, PRN, codePhase, settings, oldFreq, oldIF)              # :339-356
            dbg[PRN]["fine"] = FineRes
    return (acq, dbg) if return_debug else acq


# ----------------------------------------------------------------------------
# a15  preRun
# ----------------------------------------------------------------------------
def preRun(acqResults, settings, signal: str = "B1C"):
    """B1C/include/preRun.m:44-76, B2a/include/preRun.m:44-76 (B2a: codeFreq = basis)."""
    ch = [Settings(PRN=0, acquiredFreq=0.0, codePhase=0, codeFreq=0.0, status="-")
          for _ in range(settings.numberOfChannels)]
    pm = np.asarray(acqResults.peakMetric)
    order = np.argsort(-pm, kind="stable")                     # sort(...,'descend') is stable
    n = min(settings.numberOfChannels, int(np.sum(np.asarray(acqResults.carrFreq) != 0)))
    for ii in range(n):
        p = int(order[ii])
        ch[ii].PRN = p + 1
        ch[ii].acquiredFreq = float(acqResults.carrFreq[p])
        ch[ii].codePhase = int(acqResults.codePhase[p])
        if signal == "B1C":
            ch[ii].codeFreq = settings.codeFreqBasis - (ch[ii].acquiredFreq - settings.IF) / settings.carrFreqBasis * settings.codeFreqBasis
        else:
            ch[ii].codeFreq = float(settings.codeFreqBasis)
        ch[ii].status = "T"
    return ch


# ----------------------------------------------------------------------------
# a13  loop constants
# ----------------------------------------------------------------------------
def calcLoopCoef(LBW, zeta, k):
    """Common/calcLoopCoef.m:41-45."""
    Wn = LBW * 8 * zeta / (4 * zeta ** 2 + 1)
    tau1 = k / (Wn * Wn)
    tau2 = 2.0 * zeta / Wn
    return tau1, tau2


def calcLoopCoefCarr(settings):
    """Common/calcLoopCoefCarr.m:41-56 -> (pf3, pf2, pf1)."""
    Wn = 1.2 * settings.pllNoiseBandwidth
    T = settings.intTime
    return Wn ** 3 * T ** 2, 2 * Wn ** 2 * T, 2 * Wn


def CalcWeighingFactor(settings):
    """B1C/include/CalcWeighingFactor.m:43-81 (MATLAB ``integral`` -> scipy ``quad``)."""
    from scipy.integrate import quad

    fc = settings.codeFreqBasis
    Tc = 1 / fc
    Br = settings.FEBW

    def g11(f):
        if f == 0:
            return 0.0
        return Tc * (math.sin(math.pi / 2 * f / fc) * math.sin(math.pi * f / fc) / math.cos(math.pi / 2 * f / fc) * fc / f / math.pi) ** 2

    def g61(f):
        if f == 0:
            return 0.0
        return Tc * (math.sin(math.pi / 12 * f / fc) * math.sin(math.pi * f / fc) / math.cos(math.pi / 12 * f / fc) * fc / f / math.pi) ** 2

    def gp(f):
        return 29 / 33 * g11(f) + 4 / 33 * g61(f)

    opts = dict(limit=2000, epsabs=0, epsrel=1e-11)
    # singularities of the integrands (cos = 0) are removable; split at them
    pts11 = [k * fc for k in range(-30, 31, 2) if abs(k * fc) < Br / 2 and k % 4 != 0]
    P11_2 = 2 * quad(lambda f: g11(f) * f * f, 0, Br / 2, points=[p for p in pts11 if p > 0] or None, **opts)[0]
    P11 = 2 * quad(g11, 0, Br / 2, points=[p for p in pts11 if p > 0] or None, **opts)[0]
    pts = sorted(set([p for p in pts11 if p > 0] + [k * 6 * fc for k in (1, 3) if k * 6 * fc < Br / 2]))
    Pp_2 = 2 * quad(lambda f: gp(f) * f * f, 0, Br / 2, points=pts or None, **opts)[0]
    Pp = 2 * quad(gp, 0, Br / 2, points=pts or None, **opts)[0]
    rem11 = (P11_2 / P11) ** 0.5
    remp = (Pp_2 / Pp) ** 0.5
    t1 = 11 * P11 * rem11 ** 2
    t2 = 33 * Pp * remp ** 2
    return t1 / (t1 + t2)


# ----------------------------------------------------------------------------
# a14  C/N0 + PLL lock detector
# ----------------------------------------------------------------------------
def _cno_pld_one(I_P, Q_P, T):
    Z = I_P ** 2 + Q_P ** 2
    Zm = Z.mean()
    Zv = Z.var(ddof=1)
    with np.errstate(invalid="ignore", divide="ignore"):
        Pav = np.sqrt(np.float64(Zm ** 2 - Zv))
        Nv = 0.5 * (Zm - Pav)
        cno = np.abs((1 / T) * Pav / (2 * Nv))
        a = (I_P[I_P > 0].sum() - I_P[I_P < 0].sum()) ** 2
        q = Q_P.sum() ** 2
        pld = (a - q) / (a + q)
    return cno, pld


def Calc_CNo_PLD(tr, settings, loopCnt: int, mode: str):
    """B1C/include/Calc_CNo_PLD.m:45-114 and B2a/include/Calc_CNo_PLD.m:38-100.

    ``mode`` in {"WB","NB","B2a"}; loopCnt is 1-based.  Pilot (I,Q) are used
    as stored for WB (pilotTRKflag==2) and swapped for NB / B2a (flag==1).
    """
    n = settings.CNoInterval
    sl = slice(loopCnt - n, loopCnt)
    CNo = np.zeros(3)
    PLD = np.zeros(2)
    T = settings.intTime
    with np.errstate(invalid="ignore", divide="ignore"):
        d, PLD[0] = _cno_pld_one(tr.I_P[sl], tr.Q_P[sl], T)
        CNo[0] = 10 * np.log10(d)
        p = 0.0
        flag = settings.pilotTRKflag
        has_pilot = (mode == "WB" and flag == 2) or (mode in ("NB", "B2a") and flag == 1)
        if has_pilot:
            if mode == "WB":
                ip, qp = tr.Pilot_I_P[sl], tr.Pilot_Q_P[sl]
            else:
                qp, ip = tr.Pilot_I_P[sl], tr.Pilot_Q_P[sl]
            p, PLD[1] = _cno_pld_one(ip, qp, T)
            CNo[1] = 10 * np.log10(p)
        CNo[2] = 10 * np.log10(d + p)
    return CNo, PLD


# ----------------------------------------------------------------------------
# MATLAB colon operator
# ----------------------------------------------------------------------------
def colon(a: float, d: float, b: float, n_expected: int | None = None) -> np.ndarray:
    """MATLAB ``a:d:b`` for d>0 (documented two-ended construction).

    MATLAB builds the vector from both ends: with n = number of steps and
    c = the (snapped) last element, the first half is ``a + k*d`` and the
    second half ``c - k*d`` (middle element (a+c)/2 when n is even).  The last
    element snaps to ``b`` when a + n*d is within 2*eps*max(|a|,|b|) of it —
    which is always the case for the tracking code vectors, whose stop
    expression is built as ``(blksize-1)*step + start``.  See SURVEY.md §8
    quirk (ii); cannot be verified against MATLAB in this image.
    """
    tol = 2.0 * np.finfo(np.float64).eps * max(abs(a), abs(b))
    n = int(math.floor((b - a) / d + 0.5))
    if abs(a + n * d - b) >= tol:
        n = int(math.floor((b - a) / d))
        c = a + n * d
    else:
        c = b
    if n_expected is not None and n + 1 != n_expected:
        raise AssertionError(f"colon length {n + 1} != expected {n_expected}")
    out = np.empty(n + 1)
    h = n // 2
    k = np.arange(h + 1, dtype=np.float64)
    out[: h + 1] = a + k * d
    out[n - np.arange(h + 1)] = c - k * d
    if n % 2 == 0:
        out[h] = (a + c) / 2
    return out


# ----------------------------------------------------------------------------
# a9-a12  tracking
# ----------------------------------------------------------------------------
_TRK_FIELDS_COMMON = ["absoluteSample", "codeFreq", "carrFreq", "I_P", "I_E", "I_L", "Q_E", "Q_P", "Q_L"]


def _new_track_result(mode: str, settings, N: int) -> Settings:
    """Field creation order/initial values of B1C/WB_tracking.m:53-112,
    NB_tracking.m:53-100, B2a/tracking.m:48-96."""
    tr = Settings()
    tr.status = "-"
    tr.absoluteSample = np.zeros(N)
    tr.codeFreq = np.full(N, np.inf)
    tr.carrFreq = np.full(N, np.inf)
    for f in ("I_P", "I_E", "I_L", "Q_E", "Q_P", "Q_L"):
        tr[f] = np.zeros(N)
    flag = settings.pilotTRKflag
    if mode == "WB" and flag == 2:
        for f in ("Pilot_I_P", "Pilot_I_E", "Pilot_I_L", "Pilot_Q_E", "Pilot_Q_P", "Pilot_Q_L"):
            tr[f] = np.zeros(N)
    elif mode in ("NB", "B2a") and flag == 1:
        tr.Pilot_I_P = np.zeros(N)
        tr.Pilot_Q_P = np.zeros(N)
    for f in ("dllDiscr", "dllDiscrFilt", "pllDiscr", "pllDiscrFilt", "remCodePhase", "remCarrPhase"):
        tr[f] = np.full(N, np.inf)
    nc = N // settings.CNoInterval
    tr.DataCNo = np.zeros(nc)
    tr.DataPLD = np.zeros(nc)
    if (mode == "WB" and flag == 2) or (mode in ("NB", "B2a") and flag == 1):
        tr.PilotCNo = np.zeros(nc)
        tr.PilotPLD = np.zeros(nc)
        tr["B1C_CNo" if mode != "B2a" else "B2a_CNo"] = np.zeros(nc)
    return tr


def num_to_process(mode: str, settings) -> int:
    if mode == "B2a":
        return int(settings.msToProcess)                     # B2a/tracking.m:100
    return matlab_round(settings.msToProcess / 1000 / settings.intTime)   # WB_tracking.m:56


def correlate_epoch(mode, settings, raw, codes, remCodePhase, codePhaseStep, carrFreq, remCarrPhase):
    """One integrate-and-dump epoch.

    WB: B1C/WB_tracking.m:289-380; NB: NB_tracking.m:271-343; B2a: tracking.m:260-331.
    Returns (sums dict, remCodePhase_next, remCarrPhase_next).
    ``raw`` has exactly blksize samples.  ``codes`` are the 1-padded replicas.
    """
    blksize = raw.size
    d = settings.dllCorrelatorSpacing
    L = settings.codeLength
    fs = settings.samplingFreq
    flag = settings.pilotTRKflag
    out = {}
    b1c = mode in ("WB", "NB")
    mul = 2.0 if b1c else 1.0
    reps = {}
    tP_last = None
    for name, off in (("E", -d), ("L", d), ("P", 0.0)):
        if b1c:
            a = (remCodePhase + off) * 2 if off != 0.0 else remCodePhase * 2
            stop = ((blksize - 1) * codePhaseStep + remCodePhase + off) * 2 if off != 0.0 else ((blksize - 1) * codePhaseStep + remCodePhase) * 2
            tcode = colon(a, codePhaseStep * 2, stop, blksize)
        else:
            a = (remCodePhase + off) if off != 0.0 else remCodePhase
            stop = ((blksize - 1) * codePhaseStep + remCodePhase + off) if off != 0.0 else ((blksize - 1) * codePhaseStep + remCodePhase)
            tcode = colon(a, codePhaseStep, stop, blksize)
        tcode2 = np.ceil(tcode).astype(np.int64)             # +1 for MATLAB 1-based == 0-based into padded array
        reps[("d", name)] = codes["data"][tcode2]
        if "pilot" in codes:
            reps[("p", name)] = codes["pilot"][tcode2]
        if "pilot61" in codes:
            reps[("p61", name)] = codes["pilot61"][np.ceil(tcode * 6).astype(np.int64)]
        if name == "P":
            tP_last = tcode[blksize - 1]
    if b1c:
        rem_next = tP_last / 2 + codePhaseStep - L           # WB:327
    else:
        rem_next = (tP_last + codePhaseStep) - L             # B2a:295
    time = np.arange(blksize + 1, dtype=np.float64) / fs
    trigarg = ((carrFreq * 2.0 * np.pi) * time) + remCarrPhase
    remCarr_next = math.fmod(trigarg[blksize], 2 * np.pi)
    if b1c:
        carrsig = np.exp(-1j * trigarg[:blksize])
        m = carrsig * raw
        iB, qB = m.real, m.imag                              # WB:345-346
    else:
        carrsig = np.exp(1j * trigarg[:blksize])
        m = carrsig * raw
        qB, iB = m.real, m.imag                              # B2a:313-314
    for (fam, name), rep in reps.items():
        out[f"{fam}_I_{name}"] = float(np.sum(rep * iB))
        out[f"{fam}_Q_{name}"] = float(np.sum(rep * qB))
    return out, rem_next, remCarr_next


def _pad(code):
    return np.concatenate([code[-1:], code, code[:1]])


def make_track_codes(mode, settings, PRN):
    """Padded replicas: WB_tracking.m:179-192, NB_tracking.m:170-176, B2a/tracking.m:156-165."""
    flag = settings.pilotTRKflag
    codes = {}
    if mode in ("WB", "NB"):
        codes["data"] = _pad(generateDataBOC11(settings, PRN))
        if (mode == "WB" and flag == 2) or (mode == "NB" and flag == 1):
            codes["pilot"] = _pad(generatePilotBOC11(settings, PRN))
        if mode == "WB" and flag == 2:
            codes["pilot61"] = _pad(generatePilotBOC61(settings, PRN))
    else:
        codes["data"] = _pad(generateB2aDataCode(PRN, settings))
        if flag == 1:
            codes["pilot"] = _pad(generateB2aPilotCode(PRN, settings))
    return codes


@dataclass
class LoopState:
    codeFreq: float
    remCodePhase: float = 0.0
    carrFreq: float = 0.0
    carrFreqBasis: float = 0.0
    remCarrPhase: float = 0.0
    oldCodeNco: float = 0.0
    oldCodeError: float = 0.0
    d2CarrError: float = 0.0
    dCarrError: float = 0.0
    pos: int = 0          # 0-based absolute sample index of the next block start


def _atan_div(q, i):
    """MATLAB atan(Q/I) with IEEE division semantics (x/0 -> +-Inf, 0/0 -> NaN)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        return float(np.arctan(np.float64(q) / np.float64(i)))


def _dll(ie, qe, il, ql):
    with np.errstate(divide="ignore", invalid="ignore"):
        e = np.sqrt(np.float64(ie * ie + qe * qe))
        l = np.sqrt(np.float64(il * il + ql * ql))
        return float((e - l) / (e + l))


def close_loops(mode, settings, s, st: LoopState, coef, chCodeFreq):
    """Discriminators + loop filters.  WB_tracking.m:375-430, NB_tracking.m:349-395,
    B2a/tracking.m:337-389.  ``s`` = sums dict from correlate_epoch.
    Returns the per-epoch stored values; mutates ``st`` (carrFreq, codeFreq, filter memory)."""
    tau1, tau2, pf3, pf2, pf1, factor = coef
    d = settings.dllCorrelatorSpacing
    flag = settings.pilotTRKflag
    PDI = settings.intTime
    o = {}
    I_E, Q_E, I_P, Q_P, I_L, Q_L = (s["d_I_E"], s["d_Q_E"], s["d_I_P"], s["d_Q_P"], s["d_I_L"], s["d_Q_L"])
    carrError = _atan_div(Q_P, I_P) / (2.0 * np.pi)
    if mode == "WB" and flag == 2:
        a, b = math.sqrt(4 / 33), math.sqrt(29 / 33)
        p = {}
        for n in ("E", "P", "L"):
            p["I_" + n] = -a * s[f"p61_I_{n}"] + b * s[f"p_Q_{n}"]     # WB:375-380
            p["Q_" + n] = -a * s[f"p61_Q_{n}"] - b * s[f"p_I_{n}"]
        pe = _atan_div(p["Q_P"], p["I_P"]) / (2.0 * np.pi)
        carrError = (carrError * 1 + pe * 3) / 4
        for n in ("E", "P", "L"):
            o["Pilot_I_" + n] = p["I_" + n]
            o["Pilot_Q_" + n] = p["Q_" + n]
    elif mode == "NB" and flag == 1:
        pe = _atan_div(-s["p_I_P"], s["p_Q_P"]) / (2.0 * np.pi)       # NB:357
        carrError = (carrError * 11 + pe * 29) / 40
        o["Pilot_I_P"] = s["p_I_P"]
        o["Pilot_Q_P"] = s["p_Q_P"]
    elif mode == "B2a" and flag == 1:
        QI = (s["p_I_P"] + 1j * s["p_Q_P"]) * np.exp(-1j * np.pi / 2)  # B2a:345
        pe = _atan_div(QI.imag, QI.real) / (2.0 * np.pi)
        carrError = (carrError + pe) / 2
        o["Pilot_I_P"] = s["p_I_P"]
        o["Pilot_Q_P"] = s["p_Q_P"]
    st.d2CarrError = st.d2CarrError + carrError * pf3
    st.dCarrError = st.d2CarrError + carrError * pf2 + st.dCarrError
    carrNco = st.dCarrError + carrError * pf1
    o["carrFreq"] = st.carrFreq
    st.carrFreq = st.carrFreqBasis + carrNco

    codeError = _dll(I_E, Q_E, I_L, Q_L)
    if mode in ("WB", "NB"):
        codeError = codeError * (1 - d)
    if mode == "WB" and flag == 2:
        pc = _dll(o["Pilot_I_E"], o["Pilot_Q_E"], o["Pilot_I_L"], o["Pilot_Q_L"]) * (1 - d)
        codeError = codeError * factor + pc * (1 - factor)
    elif mode == "NB" and flag == 1:
        pc = _dll(s["p_I_E"], s["p_Q_E"], s["p_I_L"], s["p_Q_L"]) * (1 - d)
        codeError = (codeError * 11 + pc * 29) / 40
    elif mode == "B2a" and flag == 1:
        pc = _dll(s["p_I_E"], s["p_Q_E"], s["p_I_L"], s["p_Q_L"])
        codeError = (codeError + pc) / 2
    codeNco = st.oldCodeNco + (tau2 / tau1) * (codeError - st.oldCodeError) + codeError * (PDI / tau1)
    st.oldCodeNco = codeNco
    st.oldCodeError = codeError
    o["codeFreq"] = st.codeFreq
    st.codeFreq = chCodeFreq - codeNco
    o.update(dllDiscr=codeError, dllDiscrFilt=codeNco, pllDiscr=carrError, pllDiscrFilt=carrNco,
             I_E=I_E, I_P=I_P, I_L=I_L, Q_E=Q_E, Q_P=Q_P, Q_L=Q_L)
    return o


def loop_coefficients(mode, settings):
    tau1, tau2 = calcLoopCoef(settings.dllNoiseBandwidth, settings.dllDampingRatio, 1.0)
    pf3, pf2, pf1 = calcLoopCoefCarr(settings)
    factor = CalcWeighingFactor(settings) if (mode == "WB") else 0.0
    return (tau1, tau2, pf3, pf2, pf1, factor)


def tracking(mode: str, data: np.ndarray, channel, settings, n_epochs: int | None = None,
             record_nco: bool = False, correlator=None):
    """[trackResults, channel] = {WB_,NB_,}tracking(fid, channel, settings).

    ``data`` is the whole IF record (the file contents after byte 0, real int8
    samples, or complex for fileType 2); the per-channel ``fseek`` to
    ``skipNumberOfBytes + codePhase - 1`` (WB_tracking.m:174-176) becomes an
    index.  A short read stops that channel *and the whole function* like the
    reference's bare ``return`` (WB_tracking.m:279-283): the remaining channels
    keep their template values.
    ``correlator`` lets tests substitute the C oracle for ``correlate_epoch``.
    """
    N = num_to_process(mode, settings) if n_epochs is None else n_epochs
    results = [_new_track_result(mode, settings, N) for _ in range(settings.numberOfChannels)]
    coef = loop_coefficients(mode, settings)
    corr = correlator or correlate_epoch
    L = settings.codeLength
    for chn in range(settings.numberOfChannels):
        ch = channel[chn]
        if ch.PRN == 0:
            continue
        tr = results[chn]
        tr.PRN = ch.PRN
        codes = make_track_codes(mode, settings, ch.PRN)
        st = LoopState(codeFreq=ch.codeFreq, carrFreq=ch.acquiredFreq, carrFreqBasis=ch.acquiredFreq,
                       pos=int(settings.skipNumberOfBytes + ch.codePhase - 1))
        CNoValue = np.zeros(3)
        tempCNo = np.zeros(3)
        low_lock, lost = 0, False
        if record_nco:
            tr.nco = np.zeros((N, 6))
        for k in range(N):
            tr.absoluteSample[k] = st.pos
            step = st.codeFreq / settings.samplingFreq
            blksize = int(math.ceil((L - st.remCodePhase) / step))
            if st.pos + blksize > data.size:
                return results, channel                       # short read: bare return
            raw = data[st.pos: st.pos + blksize]
            if record_nco:
                tr.nco[k] = (st.pos, blksize, st.remCodePhase, step, st.carrFreq, st.remCarrPhase)
            tr.remCodePhase[k] = st.remCodePhase
            tr.remCarrPhase[k] = st.remCarrPhase
            s, st.remCodePhase, st.remCarrPhase = corr(mode, settings, raw, codes, st.remCodePhase, step,
                                                       st.carrFreq, st.remCarrPhase)
            st.pos += blksize
            o = close_loops(mode, settings, s, st, coef, ch.codeFreq)
            for f, v in o.items():
                tr[f][k] = v
            if (k + 1) % settings.CNoInterval == 0:
                CNoValue, PLD = Calc_CNo_PLD(tr, settings, k + 1, mode)
                c = (k + 1) // settings.CNoInterval - 1
                tr.DataCNo[c] = CNoValue[0] * 0.5 + tempCNo[0] * 0.5
                tr.DataPLD[c] = PLD[0]
                if "PilotCNo" in tr:
                    tr.PilotCNo[c] = CNoValue[1] * 0.5 + tempCNo[1] * 0.5
                    tr["B1C_CNo" if mode != "B2a" else "B2a_CNo"][c] = CNoValue[2] * 0.5 + tempCNo[2] * 0.5
                    tr.PilotPLD[c] = PLD[1]
                # EXTENSION (not in the reference, which copies channel.status unconditionally, WB_tracking.m:485-488):
                # settings.lockLossPLD > 0 drops a channel whose lock detector stays below it for lockLossIntervals
                # consecutive C/N0 intervals; the channel keeps status '-' and the following channels are still tracked
                if settings.get("lockLossPLD", 0) > 0:
                    pld = PLD[1] if "PilotCNo" in tr else PLD[0]
                    low_lock = low_lock + 1 if pld < settings.lockLossPLD else 0
                    if low_lock >= max(1, int(settings.get("lockLossIntervals", 1))):
                        tr.lockLostEpoch = k + 1
                        lost = True
            tempCNo = CNoValue
            if lost:
                break
        if not lost:
            tr.status = ch.status
    return results, channel


def WB_tracking(data, channel, settings, **kw):
    return tracking("WB", data, channel, settings, **kw)


def NB_tracking(data, channel, settings, **kw):
    return tracking("NB", data, channel, settings, **kw)


def B2a_tracking(data, channel, settings, **kw):
    return tracking("B2a", data, channel, settings, **kw)


# ----------------------------------------------------------------------------
# 8(f) rank 3  frame synchronisation (first consumer of the tracking output)
# ----------------------------------------------------------------------------
# generate2ndCode.m:44-56: pilot secondary code (w, p) per PRN (ICD constants)
B1C_2ND_WP = [(269, 1889), (1448, 1268), (1028, 1593), (1324, 1186), (822, 1239), (5, 1930), (155, 176), (458, 1696),
              (310, 26), (959, 1344), (1238, 1271), (1180, 1182), (1288, 1381), (334, 1604), (885, 1333), (1362, 1185),
              (181, 31), (1648, 704), (838, 1190), (313, 1646), (750, 1385), (225, 113), (1477, 860), (309, 1656),
              (108, 1921), (1457, 1173), (149, 1928), (322, 57), (271, 150), (576, 1214), (1103, 1148), (450, 1458),
              (399, 1519), (241, 1635), (1045, 1257), (164, 1687), (513, 1382), (687, 1514), (422, 1), (303, 1583),
              (324, 1806), (495, 1664), (725, 1338), (780, 1111), (367, 1706), (882, 1543), (631, 1813), (37, 228),
              (647, 2871), (1043, 2884), (24, 1823), (120, 75), (134, 11), (136, 63), (158, 1937), (214, 22), (335, 1768),
              (340, 1526), (661, 1402), (889, 1445), (929, 1680), (1002, 1290), (1149, 1245)]


def generate2ndCode(PRN: int) -> np.ndarray:
    """B1C pilot secondary code, 1800 chips, +-1 (B1C/include/generate2ndCode.m:58-84): Weil code of the Legendre
    sequence of length 3607, chip ind = L(k) xor L((k + w) mod N), k = (ind + p - 1) mod N, bipolar 1 - 2*chip."""
    N = 3607
    leg = np.zeros(N, dtype=np.int64)
    for ind in range(1, N):
        leg[ind] = 1 if pow(ind, (N - 1) // 2, N) == 1 else 0      # JacobiSymbol(ind, N) == 1 for prime N; -1 -> 0
    w, p = B1C_2ND_WP[PRN - 1]
    ind = np.arange(1800)
    k = (ind + p - 1) % N
    return (1 - 2 * (leg[k] ^ leg[(k + w) % N])).astype(np.float64)


def _xcorr_nonneg_lags(bits: np.ndarray, pattern: np.ndarray) -> np.ndarray:
    """MATLAB xcorr(bits, pattern) for lags 0 .. numel(bits)-1 (the shorter input is zero padded):
    X(lag + 1) = sum_n bits(n + lag) * pattern(n)."""
    n, K = bits.size, pattern.size
    padded = np.concatenate([bits, np.zeros(K)])
    return np.array([float(np.dot(padded[lag:lag + K], pattern)) for lag in range(n)])


def frame_sync_B1C(trackResult, settings):
    """BCNAV1decoding.m:66-91 -> (XcorrResult for lags >= 0, index (1-based))."""
    bits = np.array(trackResult.Pilot_I_P if settings.pilotTRKflag == 2 else trackResult.Pilot_Q_P, dtype=np.float64)
    bits = np.where(bits > 0, 1.0, -1.0)
    X = _xcorr_nonneg_lags(bits, generate2ndCode(trackResult.PRN))
    return X, np.flatnonzero(np.abs(X) >= 1799.5) + 1


def frame_sync_B2a(I_P_InputBits):
    """BCNAV2decoding.m:69-97 -> (tlmXcorrResult for lags >= 0, index (1-based))."""
    secondCode = np.array([1, 1, 1, -1, 1], dtype=np.float64)
    preamble_bits = np.array([-1, -1, -1, 1, 1, 1, -1, 1, 1, -1, 1, 1, -1, -1, 1, -1, -1, -1, -1, 1, -1, 1, 1, 1], dtype=np.float64)
    preamble_ms = np.kron(preamble_bits, secondCode)
    bits = np.where(np.asarray(I_P_InputBits, dtype=np.float64) > 0, 1.0, -1.0)
    X = _xcorr_nonneg_lags(bits, preamble_ms)
    return X, np.flatnonzero(np.abs(X) > 115) + 1

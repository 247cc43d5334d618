// TEST INFRASTRUCTURE - not part of the product (only tests/ may build or call this).
//
// Second, independently written CPU model of the reference's closed-loop trackers and code generators: one scalar
// loop per sample, written from the MATLAB text of
//   BDS-3_B1C/WB_tracking.m:162-488, BDS-3_B1C/NB_tracking.m:146-448, BDS-3_B2a/tracking.m:134-441,
//   BDS-3_B1C/include/generateDataBOC11.m:60-90, generatePilotBOC11.m:61-94, generatePilotBOC61.m:61-96,
//   BDS-3_B2a/include/generateB2aDataCode.m:104-138, generateB2aPilotCode.m:104-138,
//   BDS-3_B1C/include/Calc_CNo_PLD.m:45-114, BDS-3_B2a/include/Calc_CNo_PLD.m:38-100
// (paths under /root/reference/BDS3_B1C_B2a) - NOT from oracle/bds_oracle.py, which vectorises with numpy, nor from
// oracle/c/bds_oracle.c.  tests/test_ref_model.py requires the two restatements to agree to 1e-12 over >= 100 closed-loop
// epochs per tracker and bit for bit on every code; a slip in either reading of the MATLAB shows up as a disagreement.
// The per-PRN constant tables (Weil (w, p), register-2 initial states) are passed in by the caller, which takes them
// from the MATLAB files themselves when /root/reference is present.
//
// Deliberate differences in method (same mathematics): Legendre symbol by Euler's criterion (modular exponentiation)
// instead of the recursive JacobiSymbol.m; sums accumulated sequentially in long double; replica values looked up
// per sample instead of through index vectors.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

const double kTwoPi = 2.0 * 3.14159265358979323846;   // MATLAB's (2 * pi)

// ---- codes ---------------------------------------------------------------------------------------
long long powmod(long long b, long long e, long long m) {
    long long r = 1;
    b %= m;
    while (e > 0) {
        if (e & 1) r = r * b % m;
        b = b * b % m;
        e >>= 1;
    }
    return r;
}

// generateDataBOC11.m:60-82 (the pilot files differ only in the (w, p) table): legendre(1) = 0,
// legendre(ind+1) = JacobiSymbol(ind, N) with -1 mapped to 0; chip ind (0-based) = L(k) xor L((k+w) mod N),
// k = (ind + p - 1) mod N; bipolar 1 - 2*chip.
void weil_primary(int w, int p, std::vector<int>& out) {
    const int N = 10243;
    std::vector<int> leg(N, 0);
    for (int ind = 1; ind <= N - 1; ++ind) leg[ind] = powmod(ind, (N - 1) / 2, N) == 1 ? 1 : 0;   // Euler: a^((N-1)/2) = (a|N)
    out.assign(10230, 0);
    for (int ind = 0; ind <= 10229; ++ind) {
        const int k = (ind + p - 1) % N;
        const int chip = leg[k] ^ leg[(k + w) % N];
        out[ind] = 1 - 2 * chip;
    }
}

// generateB2aDataCode.m:104-138: registers as +-1 row vectors (-1 = logic 1), output = register1(end)*register2(end),
// feedback = product of the tapped stages, circshift by one and insert the feedback at stage 1; register 1 is reset
// to all -1 after chip 8190.
void b2a_code(const int* taps1, int n1, const int* taps2, int n2, const int* reg2ini /*13 logic values*/, std::vector<int>& out) {
    int r1[14], r2[14];   // 1-based stages
    for (int i = 1; i <= 13; ++i) {
        r1[i] = -1;
        r2[i] = 1 - 2 * reg2ini[i - 1];
    }
    out.assign(10230, 0);
    for (int ind = 1; ind <= 10230; ++ind) {
        out[ind - 1] = r1[13] * r2[13];
        int f1 = 1, f2 = 1;
        for (int i = 0; i < n1; ++i) f1 *= r1[taps1[i]];
        for (int i = 0; i < n2; ++i) f2 *= r2[taps2[i]];
        for (int i = 13; i >= 2; --i) {
            r1[i] = r1[i - 1];
            r2[i] = r2[i - 1];
        }
        r1[1] = f1;
        r2[1] = f2;
        if (ind == 8190)
            for (int i = 1; i <= 13; ++i) r1[i] = -1;
    }
}

// ---- MATLAB colon a:d:b, element k (0-based) of n+1 elements --------------------------------------
// MATLAB builds the vector from both ends towards the middle (first half a + k*d, second half b - (n-k)*d, the
// middle element of an odd-length vector as the mean of its neighbours' construction (a + b)/2), with b the stop
// value snapped to the last element.  n = number of steps.
inline double colon_elem(double a, double d, double b, long n, long k) {
    const long h = n / 2;
    if (n % 2 == 0 && k == h) return (a + b) * 0.5;
    if (k <= h) return a + (double)k * d;
    return b - (double)(n - k) * d;
}

struct Settings {
    double samplingFreq, codeFreqBasis, dllCorrelatorSpacing, intTime;
    double tau1code, tau2code, pf3, pf2, pf1, factor;
    int codeLength, pilotTRKflag, CNoInterval, fileType;
    long long skipNumberOfBytes;
};

// Calc_CNo_PLD.m: the C/N0 of one component (variance summing) and its lock metric over n prompts
void cno_one(const double* I, const double* Q, int n, double T, double& cnoLin, double& pld) {
    long double zm = 0;
    for (int i = 0; i < n; ++i) zm += (long double)I[i] * I[i] + (long double)Q[i] * Q[i];
    const double Zm = (double)(zm / n);
    long double zv = 0;
    for (int i = 0; i < n; ++i) {
        const double z = I[i] * I[i] + Q[i] * Q[i];
        zv += (long double)(z - Zm) * (z - Zm);
    }
    const double Zv = (double)(zv / (n - 1));      // var() normalises by N-1
    const double Pav = std::sqrt(Zm * Zm - Zv);
    const double Nv = 0.5 * (Zm - Pav);
    cnoLin = std::fabs((1 / T) * Pav / (2 * Nv));
    long double sp = 0, sn = 0, sq = 0;
    for (int i = 0; i < n; ++i) {
        if (I[i] > 0) sp += I[i];
        if (I[i] < 0) sn += I[i];
        sq += Q[i];
    }
    const double a = (double)(sp - sn), b = (double)sq;
    pld = (a * a - b * b) / (a * a + b * b);
}

}  // namespace

extern "C" {

// component: 0 B1C data primary (w,p) / 1 B1C pilot primary (w,p): arg = {w, p}
int ref_weil_primary(int w, int p, int8_t* out10230) {
    std::vector<int> c;
    weil_primary(w, p, c);
    for (int i = 0; i < 10230; ++i) out10230[i] = (int8_t)c[i];
    return 0;
}
// pilot = 0: data taps (generateB2aDataCode.m:108-109), 1: pilot taps (generateB2aPilotCode.m:108-109)
int ref_b2a_code(int pilot, const int* reg2ini13, int8_t* out10230) {
    static const int d1[] = {1, 5, 11, 13}, d2[] = {3, 5, 9, 11, 12, 13};
    static const int p1[] = {3, 6, 7, 13}, p2[] = {1, 5, 7, 8, 12, 13};
    std::vector<int> c;
    if (pilot) b2a_code(p1, 4, p2, 6, reg2ini13, c);
    else b2a_code(d1, 4, d2, 6, reg2ini13, c);
    for (int i = 0; i < 10230; ++i) out10230[i] = (int8_t)c[i];
    return 0;
}

// One channel, closed loop.  mode: 1 = WB_tracking.m (pilotTRKflag 2 or 0), 2 = NB_tracking.m (flag 1 or 0),
// 3 = B2a tracking.m (flag 1 or 0).  x: the file contents as schar (fileType 1: n real samples; fileType 2: n
// interleaved I/Q pairs = 2n bytes).  dataPrim / pilotPrim: +-1 primary codes of the PRN (10230 chips).
// out: [22][nEpochs] doubles in the order absoluteSample, codeFreq, carrFreq, I_P, I_E, I_L, Q_E, Q_P, Q_L,
// Pilot_I_P, Pilot_I_E, Pilot_I_L, Pilot_Q_E, Pilot_Q_P, Pilot_Q_L, dllDiscr, dllDiscrFilt, pllDiscr, pllDiscrFilt,
// remCodePhase, remCarrPhase, (row 21 unused); cno: [5][nEpochs / CNoInterval] DataCNo, DataPLD, PilotCNo, PilotPLD,
// total.  Returns the number of epochs completed (short read stops like the reference's bare return).
int ref_track(int mode, const int8_t* x, long long n, const double* S /*settings, see below*/, const int8_t* dataPrim,
              const int8_t* pilotPrim, double chCodeFreq, double acquiredFreq, double codePhase, int nEpochs, double* out,
              double* cno) {
    Settings st;
    st.samplingFreq = S[0];
    st.codeFreqBasis = S[1];
    st.codeLength = (int)S[2];
    st.dllCorrelatorSpacing = S[3];
    st.intTime = S[4];
    st.pilotTRKflag = (int)S[5];
    st.CNoInterval = (int)S[6];
    st.tau1code = S[7];
    st.tau2code = S[8];
    st.pf3 = S[9];
    st.pf2 = S[10];
    st.pf1 = S[11];
    st.factor = S[12];
    st.fileType = (int)S[13];
    st.skipNumberOfBytes = (long long)S[14];
    const int L = st.codeLength;
    const bool b1c = mode != 3;
    const bool pilot = (mode == 1 && st.pilotTRKflag == 2) || (mode != 1 && st.pilotTRKflag == 1);
    const bool p61 = mode == 1 && st.pilotTRKflag == 2;
    const double earlyLateSpc = st.dllCorrelatorSpacing;
    const double PDIcode = st.intTime;
    // padded replicas: [code(end) code code(1)]
    std::vector<int8_t> dRep, pRep, p61Rep;
    if (b1c) {
        // BOC(1,1): chip c -> [-c, +c] (generateDataBOC11.m:85-90)
        std::vector<int8_t> d(2 * L), p(2 * L);
        for (int j = 0; j < L; ++j) {
            d[2 * j] = (int8_t)-dataPrim[j];
            d[2 * j + 1] = dataPrim[j];
            p[2 * j] = (int8_t)-pilotPrim[j];
            p[2 * j + 1] = pilotPrim[j];
        }
        dRep.push_back(d[2 * L - 1]);
        dRep.insert(dRep.end(), d.begin(), d.end());
        dRep.push_back(d[0]);
        pRep.push_back(p[2 * L - 1]);
        pRep.insert(pRep.end(), p.begin(), p.end());
        pRep.push_back(p[0]);
        // BOC(6,1): chip c -> (-1)^ii * c, ii = 1..12 (generatePilotBOC61.m:89-96)
        std::vector<int8_t> q(12 * L);
        for (int j = 0; j < L; ++j)
            for (int ii = 1; ii <= 12; ++ii) q[12 * j + ii - 1] = (int8_t)((ii % 2 ? -1 : 1) * pilotPrim[j]);
        p61Rep.push_back(q[12 * L - 1]);
        p61Rep.insert(p61Rep.end(), q.begin(), q.end());
        p61Rep.push_back(q[0]);
    } else {
        dRep.push_back(dataPrim[L - 1]);
        dRep.insert(dRep.end(), dataPrim, dataPrim + L);
        dRep.push_back(dataPrim[0]);
        pRep.push_back(pilotPrim[L - 1]);
        pRep.insert(pRep.end(), pilotPrim, pilotPrim + L);
        pRep.push_back(pilotPrim[0]);
    }
    const int dataAdaptCoeff = st.fileType == 1 ? 1 : 2;
    long long filePos = dataAdaptCoeff * (st.skipNumberOfBytes + (long long)codePhase - 1);   // fseek(..., 'bof')
    const long long fileBytes = n * dataAdaptCoeff;
    double codeFreq = chCodeFreq, remCodePhase = 0.0, carrFreq = acquiredFreq, carrFreqBasis = acquiredFreq, remCarrPhase = 0.0;
    double oldCodeNco = 0.0, oldCodeError = 0.0, d2CarrError = 0.0, dCarrError = 0.0;
    double CNoValue[3] = {0, 0, 0}, tempCNoValue[3] = {0, 0, 0};
    auto O = [&](int f, int k) -> double& { return out[(size_t)f * nEpochs + k]; };
    const int nC = st.CNoInterval > 0 ? nEpochs / st.CNoInterval : 0;
    const double mul = b1c ? 2.0 : 1.0;
    for (int loopCnt = 1; loopCnt <= nEpochs; ++loopCnt) {
        const int k = loopCnt - 1;
        O(0, k) = (double)filePos / dataAdaptCoeff;
        const double codePhaseStep = codeFreq / st.samplingFreq;
        const long blksize = (long)std::ceil((L - remCodePhase) / codePhaseStep);
        if (filePos + dataAdaptCoeff * blksize > fileBytes) return loopCnt - 1;   // samplesRead ~= blksize: return
        const int8_t* raw = x + filePos;
        filePos += dataAdaptCoeff * blksize;
        O(19, k) = remCodePhase;
        // the three colon vectors (start, step, stop) exactly as written in the MATLAB
        double a[3], b[3], dstep;
        const long nst = blksize - 1;
        if (b1c) {
            a[0] = (remCodePhase - earlyLateSpc) * 2;
            b[0] = ((blksize - 1) * codePhaseStep + remCodePhase - earlyLateSpc) * 2;
            a[1] = remCodePhase * 2;
            b[1] = ((blksize - 1) * codePhaseStep + remCodePhase) * 2;
            a[2] = (remCodePhase + earlyLateSpc) * 2;
            b[2] = ((blksize - 1) * codePhaseStep + remCodePhase + earlyLateSpc) * 2;
            dstep = codePhaseStep * 2;
        } else {
            a[0] = (remCodePhase - earlyLateSpc);
            b[0] = ((blksize - 1) * codePhaseStep + remCodePhase - earlyLateSpc);
            a[1] = remCodePhase;
            b[1] = ((blksize - 1) * codePhaseStep + remCodePhase);
            a[2] = (remCodePhase + earlyLateSpc);
            b[2] = ((blksize - 1) * codePhaseStep + remCodePhase + earlyLateSpc);
            dstep = codePhaseStep;
        }
        O(20, k) = remCarrPhase;
        long double acc[3][3][2];   // [family d,p,p61][E,P,L][I,Q]
        std::memset(acc, 0, sizeof acc);
        for (long s = 0; s < blksize; ++s) {
            const double time = (double)s / st.samplingFreq;
            const double trigarg = ((carrFreq * 2.0 * 3.14159265358979323846) * time) + remCarrPhase;
            const double c = std::cos(trigarg), sn = std::sin(trigarg);
            double xr, xi = 0.0;
            if (dataAdaptCoeff == 1) xr = raw[s];
            else {
                xr = raw[2 * s];
                xi = raw[2 * s + 1];
            }
            double iB, qB;
            if (b1c) {   // carrsig = exp(-1i*trigarg); i = real(carrsig .* raw), q = imag(...)
                iB = c * xr + sn * xi;
                qB = c * xi - sn * xr;
            } else {     // carrsig = exp(+1i*trigarg); q = real(...), i = imag(...)
                qB = c * xr - sn * xi;
                iB = c * xi + sn * xr;
            }
            for (int o = 0; o < 3; ++o) {
                const double t = colon_elem(a[o], dstep, b[o], nst, s);
                const long i2 = (long)std::ceil(t) + 1;                 // MATLAB 1-based index into the padded replica
                const double dv = dRep[i2 - 1];
                acc[0][o][0] += dv * iB;
                acc[0][o][1] += dv * qB;
                if (pilot) {
                    const double pv = pRep[i2 - 1];
                    acc[1][o][0] += pv * iB;
                    acc[1][o][1] += pv * qB;
                }
                if (p61) {
                    const long i12 = (long)std::ceil(t * 6) + 1;
                    const double qv = p61Rep[i12 - 1];
                    acc[2][o][0] += qv * iB;
                    acc[2][o][1] += qv * qB;
                }
            }
        }
        const double tP_last = colon_elem(a[1], dstep, b[1], nst, nst);
        if (b1c) remCodePhase = tP_last / 2 + codePhaseStep - L;
        else remCodePhase = (tP_last + codePhaseStep) - L;
        {
            const double time = (double)blksize / st.samplingFreq;
            const double trigarg = ((carrFreq * 2.0 * 3.14159265358979323846) * time) + remCarrPhase;
            remCarrPhase = std::fmod(trigarg, kTwoPi);
        }
        const double I_E = (double)acc[0][0][0], Q_E = (double)acc[0][0][1], I_P = (double)acc[0][1][0],
                     Q_P = (double)acc[0][1][1], I_L = (double)acc[0][2][0], Q_L = (double)acc[0][2][1];
        double pI[3] = {0, 0, 0}, pQ[3] = {0, 0, 0};   // stored pilot E, P, L
        double carrError = std::atan(Q_P / I_P) / (2.0 * 3.14159265358979323846);
        double codeError;
        if (b1c)
            codeError = (std::sqrt(I_E * I_E + Q_E * Q_E) - std::sqrt(I_L * I_L + Q_L * Q_L)) /
                        (std::sqrt(I_E * I_E + Q_E * Q_E) + std::sqrt(I_L * I_L + Q_L * Q_L)) * (1 - earlyLateSpc);
        else
            codeError = (std::sqrt(I_E * I_E + Q_E * Q_E) - std::sqrt(I_L * I_L + Q_L * Q_L)) /
                        (std::sqrt(I_E * I_E + Q_E * Q_E) + std::sqrt(I_L * I_L + Q_L * Q_L));
        if (p61) {
            for (int o = 0; o < 3; ++o) {
                pI[o] = -std::sqrt(4.0 / 33) * (double)acc[2][o][0] + std::sqrt(29.0 / 33) * (double)acc[1][o][1];
                pQ[o] = -std::sqrt(4.0 / 33) * (double)acc[2][o][1] - std::sqrt(29.0 / 33) * (double)acc[1][o][0];
            }
            const double p_carrError = std::atan(pQ[1] / pI[1]) / (2.0 * 3.14159265358979323846);
            carrError = (carrError * 1 + p_carrError * 3) / 4;
            const double p_codeError = (std::sqrt(pI[0] * pI[0] + pQ[0] * pQ[0]) - std::sqrt(pI[2] * pI[2] + pQ[2] * pQ[2])) /
                                       (std::sqrt(pI[0] * pI[0] + pQ[0] * pQ[0]) + std::sqrt(pI[2] * pI[2] + pQ[2] * pQ[2])) *
                                       (1 - earlyLateSpc);
            codeError = codeError * st.factor + p_codeError * (1 - st.factor);
        } else if (pilot && mode == 2) {
            const double p11_I_P = (double)acc[1][1][0], p11_Q_P = (double)acc[1][1][1];
            const double p11_I_E = (double)acc[1][0][0], p11_Q_E = (double)acc[1][0][1], p11_I_L = (double)acc[1][2][0],
                         p11_Q_L = (double)acc[1][2][1];
            const double p11_carrError = std::atan(-p11_I_P / p11_Q_P) / (2.0 * 3.14159265358979323846);
            carrError = (carrError * 11 + p11_carrError * 29) / 40;
            const double p11_codeError =
                (std::sqrt(p11_I_E * p11_I_E + p11_Q_E * p11_Q_E) - std::sqrt(p11_I_L * p11_I_L + p11_Q_L * p11_Q_L)) /
                (std::sqrt(p11_I_E * p11_I_E + p11_Q_E * p11_Q_E) + std::sqrt(p11_I_L * p11_I_L + p11_Q_L * p11_Q_L)) *
                (1 - earlyLateSpc);
            codeError = (codeError * 11 + p11_codeError * 29) / 40;
            pI[1] = p11_I_P;
            pQ[1] = p11_Q_P;
        } else if (pilot && mode == 3) {
            const double pilot_I_P = (double)acc[1][1][0], pilot_Q_P = (double)acc[1][1][1];
            const double pilot_I_E = (double)acc[1][0][0], pilot_Q_E = (double)acc[1][0][1], pilot_I_L = (double)acc[1][2][0],
                         pilot_Q_L = (double)acc[1][2][1];
            // QI = (pilot_I_P + 1i*pilot_Q_P) * exp(-1i*pi/2), with exp evaluated in double: cos(pi/2) = 6.1e-17
            const double er = std::cos(-3.14159265358979323846 / 2), ei = std::sin(-3.14159265358979323846 / 2);
            const double QIr = pilot_I_P * er - pilot_Q_P * ei, QIi = pilot_I_P * ei + pilot_Q_P * er;
            const double carrErrorQ = std::atan(QIi / QIr) / (2.0 * 3.14159265358979323846);
            carrError = (carrError + carrErrorQ) / 2;
            const double codeErrorQ =
                (std::sqrt(pilot_I_E * pilot_I_E + pilot_Q_E * pilot_Q_E) - std::sqrt(pilot_I_L * pilot_I_L + pilot_Q_L * pilot_Q_L)) /
                (std::sqrt(pilot_I_E * pilot_I_E + pilot_Q_E * pilot_Q_E) + std::sqrt(pilot_I_L * pilot_I_L + pilot_Q_L * pilot_Q_L));
            codeError = (codeError + codeErrorQ) / 2;
            pI[1] = pilot_I_P;
            pQ[1] = pilot_Q_P;
        }
        d2CarrError = d2CarrError + carrError * st.pf3;
        dCarrError = d2CarrError + carrError * st.pf2 + dCarrError;
        const double carrNco = dCarrError + carrError * st.pf1;
        O(2, k) = carrFreq;
        carrFreq = carrFreqBasis + carrNco;
        const double codeNco = oldCodeNco + (st.tau2code / st.tau1code) * (codeError - oldCodeError) + codeError * (PDIcode / st.tau1code);
        oldCodeNco = codeNco;
        oldCodeError = codeError;
        O(1, k) = codeFreq;
        codeFreq = chCodeFreq - codeNco;
        O(15, k) = codeError;
        O(16, k) = codeNco;
        O(17, k) = carrError;
        O(18, k) = carrNco;
        O(3, k) = I_P;
        O(4, k) = I_E;
        O(5, k) = I_L;
        O(6, k) = Q_E;
        O(7, k) = Q_P;
        O(8, k) = Q_L;
        if (pilot) {
            O(9, k) = pI[1];
            O(13, k) = pQ[1];
            if (p61) {
                O(10, k) = pI[0];
                O(11, k) = pI[2];
                O(12, k) = pQ[0];
                O(14, k) = pQ[2];
            }
        }
        if (st.CNoInterval > 0 && loopCnt % st.CNoInterval == 0) {
            const int n0 = loopCnt - st.CNoInterval;   // 0-based start of the window
            double dl, dp, pl = 0, pp = 0;
            cno_one(&O(3, n0), &O(7, n0), st.CNoInterval, st.intTime, dl, dp);
            CNoValue[0] = 10 * std::log10(dl);
            if (pilot) {
                if (p61) cno_one(&O(9, n0), &O(13, n0), st.CNoInterval, st.intTime, pl, pp);   // flag 2: I = Pilot_I_P
                else cno_one(&O(13, n0), &O(9, n0), st.CNoInterval, st.intTime, pl, pp);       // flag 1: I = Pilot_Q_P
                CNoValue[1] = 10 * std::log10(pl);
            }
            CNoValue[2] = 10 * std::log10(dl + pl);
            const int cc = loopCnt / st.CNoInterval - 1;
            cno[0 * nC + cc] = CNoValue[0] * 0.5 + tempCNoValue[0] * 0.5;
            cno[1 * nC + cc] = dp;
            if (pilot) {
                cno[2 * nC + cc] = CNoValue[1] * 0.5 + tempCNoValue[1] * 0.5;
                cno[3 * nC + cc] = pp;
                cno[4 * nC + cc] = CNoValue[2] * 0.5 + tempCNoValue[2] * 0.5;
            }
        }
        for (int i = 0; i < 3; ++i) tempCNoValue[i] = CNoValue[i];
    }
    return nEpochs;
}

}  // extern "C"

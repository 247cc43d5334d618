// Tracking correlator + closed-loop persistent kernel (sm_100a).
//
// Replaces the per-channel, per-epoch loop of the reference
//   BDS-3_B1C/WB_tracking.m:223-483, BDS-3_B1C/NB_tracking.m:205-448, BDS-3_B2a/tracking.m:195-441
// Design: one persistent cooperative grid; work items are
// (channel, epoch, slice).  A slice correlates a contiguous part of the epoch's
// sample block; the last slice to arrive (atomic counter) reduces the partials in a
// fixed order, closes the PLL/DLL in fp64 exactly as the reference does and
// publishes the next epoch's NCO parameters with a release store, which makes
// that channel's next-epoch slices runnable.  CTAs interleave the channels of their
// group so the loop-closure latency of one channel is hidden behind the others.
#include <algorithm>
#include <atomic>
#include <climits>
#include <thread>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <sys/mman.h>
#include <sys/stat.h>
#include <fcntl.h>
#include <unistd.h>

#include "bds_codes.h"
#include "bds_track.cuh"
#define FAST_GEOM_MULTI   // the chip-synchronous B1C kernel is instantiated once per geometry (g99, g53) below
#include "bds_track_fast.cuh"

namespace bds {

// ======================================================================================
// General (exact, any configuration) slice correlator
// ======================================================================================
// Follows the reference sample by sample, including MATLAB's two-ended colon
// construction of the code-phase vectors (SURVEY §8 quirk ii) so that chip lookups
// are decided by the same IEEE expressions as the float64 oracle.
struct SliceCtx {
    const int8_t* x;        // window base
    long long winFirst, winLen;
    EpochParams p;
    int mode, hasPilot, hasP61;
    double L, d, fs;
    int iq;                 // 1: interleaved I/Q int8 pairs (settings.fileType == 2, WB_tracking.m:155-159,270-274)
};

template <typename T>
__device__ __forceinline__ T load_cg(const T* p) {
    static_assert(sizeof(T) % 16 == 0, "16-byte multiple");
    T v;
    const uint4* s = reinterpret_cast<const uint4*>(p);
    uint4* d = reinterpret_cast<uint4*>(&v);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); ++i) d[i] = __ldcg(s + i);
    return v;
}
template <typename T>
__device__ __forceinline__ void store_cg(T* p, const T& v) {
    uint4* d = reinterpret_cast<uint4*>(p);
    const uint4* s = reinterpret_cast<const uint4*>(&v);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); ++i) __stcg(d + i, s[i]);
}

__device__ __forceinline__ int code_bit(const uint32_t* w, int chip) { return (w[chip >> 5] >> (chip & 31)) & 1; }

// two-ended colon element k of a:dd:stop with n steps (numel = n+1)
__device__ __forceinline__ double colon_elem(double a, double dd, double stop, int n, int k) {
    int h = n >> 1;
    if (!(n & 1) && k == h) return __dmul_rn(__dadd_rn(a, stop), 0.5);
    if (k <= h) return __dadd_rn(a, __dmul_rn((double)k, dd));
    return __dadd_rn(stop, -__dmul_rn((double)(n - k), dd));
}

// sign (+1/-1 as float) of the BOC(1,1)-expanded code at padded index idx (0-based
// into [code(end) code code(1)]): WB_tracking.m:181,292-294
__device__ __forceinline__ float boc11_sign(const uint32_t* w, int idx, int L2) {
    int h = idx - 1;
    if (h < 0) h = L2 - 1;
    if (h >= L2) h = 0;
    int neg = code_bit(w, h >> 1) ^ ((h & 1) == 0);
    return neg ? -1.f : 1.f;
}
__device__ __forceinline__ float boc61_sign(const uint32_t* w, int idx, int L12) {
    int s = idx - 1;
    if (s < 0) s = L12 - 1;
    if (s >= L12) s = 0;
    int chip = s / 12;
    int i = s - chip * 12;
    int neg = code_bit(w, chip) ^ ((i & 1) == 0);
    return neg ? -1.f : 1.f;
}
__device__ __forceinline__ float plain_sign(const uint32_t* w, int idx, int L) {
    int c = idx - 1;
    if (c < 0) c = L - 1;
    if (c >= L) c = 0;
    return code_bit(w, c) ? -1.f : 1.f;
}

// Accumulates this thread's share of block-relative samples in the 16-byte chunk
// range [q0, q1) of the window into acc[18].
__device__ void correlate_general(const SliceCtx& c, const uint32_t* bitsData, const uint32_t* bitsPilot,
                                  long long q0, long long q1, float* acc) {
    const EpochParams& p = c.p;
    const long long B0 = p.pos - c.winFirst;  // sample offset of the block in the window
    const int iq = c.iq;                      // bytes per sample = 1 << iq
    const long long winBytes = c.winLen << iq;
    const int n = p.blksize - 1;              // colon steps
    const bool b1c = c.mode != BDS_TRK_B2A;
    const double mul = b1c ? 2.0 : 1.0;
    const double dd = b1c ? __dmul_rn(p.step, 2.0) : p.step;
    const double base = __dadd_rn(__dmul_rn((double)n, p.step), p.rem);  // (blksize-1)*step + rem
    double a[3], stop[3];
    // order E, P, L: offsets -d, 0, +d   (WB_tracking.m:289-317)
    a[0] = __dmul_rn(__dadd_rn(p.rem, -c.d), mul);
    a[1] = __dmul_rn(p.rem, mul);
    a[2] = __dmul_rn(__dadd_rn(p.rem, c.d), mul);
    stop[0] = __dmul_rn(__dadd_rn(base, -c.d), mul);
    stop[1] = __dmul_rn(base, mul);
    stop[2] = __dmul_rn(__dadd_rn(base, c.d), mul);
    // carrier NCO in 2^-64 turns
    double r = p.carrFreq / c.fs;
    r -= floor(r);
    const unsigned long long dphi = __double2ull_rn(r * 18446744073709551616.0);
    double r0 = p.remCarr / 6.283185307179586476925286766559;
    r0 -= floor(r0);
    const unsigned long long phi0 = __double2ull_rn(r0 * 18446744073709551616.0);
    const int Li = (int)c.L;

    for (long long q = q0 + threadIdx.x; q < q1; q += blockDim.x) {
        const int8_t* src = c.x + q * 16;
        uint4 v = make_uint4(0, 0, 0, 0);
        if ((q + 1) * 16 <= winBytes) {
            v = ldg_nc_v4(src);
        } else {  // last, partial chunk of a caller-owned buffer: byte loads only
            int8_t* vb = reinterpret_cast<int8_t*>(&v);
            for (int j = 0; j < 16 && q * 16 + j < winBytes; ++j) vb[j] = src[j];
        }
        const int8_t* b = reinterpret_cast<const int8_t*>(&v);
        const int per = 16 >> iq;   // samples in the chunk
#pragma unroll 1
        for (int j = 0; j < per; ++j) {
            long long k = ((q * 16) >> iq) + j - B0;
            if (k < 0 || k > n) continue;
            const float xs = (float)b[j << iq];
            const float xi = iq ? (float)b[(j << 1) + 1] : 0.f;   // rawSignal = I + 1i*Q, WB_tracking.m:270-274
            unsigned long long ph = phi0 + (unsigned long long)k * dphi;
            float sn, cs;
            sincospif((float)(int)(ph >> 32) * 4.656612873077392578125e-10f, &sn, &cs);  // 2^-31 -> angle/pi
            float iB, qB;
            if (b1c) {  // carrsig = exp(-i*theta), i = real(carrsig .* x), q = imag(.): WB_tracking.m:341-346
                iB = xs * cs + xi * sn;
                qB = xi * cs - xs * sn;
            } else {    // exp(+i*theta), I = imag, Q = real: B2a tracking.m:309-314
                qB = xs * cs - xi * sn;
                iB = xs * sn + xi * cs;
            }
#pragma unroll
            for (int o = 0; o < 3; ++o) {
                double t = colon_elem(a[o], dd, stop[o], n, (int)k);
                int idx = (int)ceil(t);
                float sd, sp = 0.f, s6 = 0.f;
                if (b1c) {
                    sd = boc11_sign(bitsData, idx, 2 * Li);
                    if (c.hasPilot) sp = boc11_sign(bitsPilot, idx, 2 * Li);
                    if (c.hasP61) s6 = boc61_sign(bitsPilot, (int)ceil(__dmul_rn(t, 6.0)), 12 * Li);
                } else {
                    sd = plain_sign(bitsData, idx, Li);
                    if (c.hasPilot) sp = plain_sign(bitsPilot, idx, Li);
                }
                acc[sum_idx(0, o, 0)] += sd * iB;
                acc[sum_idx(0, o, 1)] += sd * qB;
                acc[sum_idx(1, o, 0)] += sp * iB;
                acc[sum_idx(1, o, 1)] += sp * qB;
                acc[sum_idx(2, o, 0)] += s6 * iB;
                acc[sum_idx(2, o, 1)] += s6 * qB;
            }
        }
    }
}

// chunk range of slice sl of S for the epoch block
__device__ __forceinline__ void slice_chunks(const EpochParams& p, long long winFirst, int iq, int sl, int S, long long& q0,
                                             long long& q1) {
    long long B0 = (p.pos - winFirst) << iq;            // byte range of the block in the window
    long long B1 = B0 + ((long long)p.blksize << iq);
    long long qa = B0 >> 4, qb = (B1 + 15) >> 4;
    long long nq = qb - qa;
    q0 = qa + nq * sl / S;
    q1 = qa + nq * (sl + 1) / S;
}

// block-wide reduction of acc[18] (fp32 per thread) -> double, result valid in
// threads 0..17 of warp 0 ... returned through smem red[18]
__device__ void block_reduce18(float* acc, double* red /*[8][18] smem*/, double* out18 /*smem [18]*/) {
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < kNSum; ++i) {
        double v = warp_sum((double)acc[i]);
        if (lane == 0) red[wid * kNSum + i] = v;
    }
    __syncthreads();
    if (threadIdx.x < kNSum) {
        double s = 0;
        int nw = blockDim.x >> 5;
        for (int w = 0; w < nw; ++w) s += red[w * kNSum + threadIdx.x];
        out18[threadIdx.x] = s;
    }
    __syncthreads();
}

// ======================================================================================
// Loop closure (fp64, one thread)  — WB_tracking.m:375-481, NB_tracking.m:349-445,
// B2a/tracking.m:337-434, Calc_CNo_PLD.m
// ======================================================================================
__device__ __forceinline__ double dll_disc(double ie, double qe, double il, double ql) {
    double e = sqrt(ie * ie + qe * qe), l = sqrt(il * il + ql * ql);
    return (e - l) / (e + l);
}

__device__ void cno_pld(const double* ip, const double* qp, int n, double T, double& cno, double& pld) {
    // Calc_CNo_PLD.m:52-73 ; var() normalises by N-1
    double zm = 0;
    for (int i = 0; i < n; ++i) {
        double a = __ldcg(ip + i), b = __ldcg(qp + i);
        zm += a * a + b * b;
    }
    zm /= n;
    double zv = 0, pos = 0, neg = 0, sq = 0;
    for (int i = 0; i < n; ++i) {
        double a = __ldcg(ip + i), b = __ldcg(qp + i);
        double z = a * a + b * b - zm;
        zv += z * z;
        if (a > 0) pos += a;
        if (a < 0) neg += a;
        sq += b;
    }
    zv /= (n - 1);
    double pav = sqrt(zm * zm - zv);
    double nv = 0.5 * (zm - pav);
    cno = fabs((1.0 / T) * pav / (2.0 * nv));
    double aa = (pos - neg) * (pos - neg), qq = sq * sq;
    pld = (aa - qq) / (aa + qq);
}

// Computes the params of the epoch that follows state `st`; returns false if the
// block is not fully inside the resident window (short read, WB_tracking.m:279-283).
__device__ bool next_params(const TrkDev& g, const ChanState& st, EpochParams& np) {
    np.pos = st.pos;
    np.rem = st.remCodePhase;
    np.step = st.codeFreq / g.fs;                                 // WB:258
    np.blksize = (int)ceil((g.L - st.remCodePhase) / np.step);    // WB:261
    np.carrFreq = st.carrFreq;
    np.remCarr = st.remCarrPhase;
    np.pad = 0;
    long long off = np.pos - g.winFirst;
    return off >= 0 && off + np.blksize <= g.winLen && np.blksize > 0 && st.lockLost == 0;
}

// Loop-closure arithmetic of one epoch (no memory traffic), in two phases so that a latency-critical caller can publish
// the next epoch's NCO before it formats this epoch's outputs:
//   close_nco  discriminators, loop filters, state update (everything the next epoch depends on);
//   close_out  the value of every trackResults plane for this epoch (plus the raw sums).
//   s[18]  correlator sums;  p  the epoch's NCO parameters;  st  channel state (updated in place);
//   pre (optional; modes with a pilot) = {atan(Q_P/I_P)/2pi, atan(pQ_P/pI_P)/2pi, dll(data), dll(pilot), fmod(trig,2pi)}
//   already evaluated with the expressions below by other lanes of a closing warp.
struct CloseAux {
    double carrError, codeError, carrNco, codeNco, carrFreqOld, codeFreqOld;
    double pI[3], pQ[3];   // stored pilot values (E, P, L): WB composite, NB / B2a the pilot prompt itself in [EPL_P]
};

__device__ void close_nco(const TrkDev& g, const double* s, const EpochParams& p, double chCodeFreq, ChanState& st,
                          CloseAux& a, const double* pre = nullptr) {
    const double twopi = 6.283185307179586476925286766559;
    const bool b1c = g.mode != BDS_TRK_B2A;
    // remCodePhase / remCarrPhase updates (WB:327,335-337; B2a:295,303-305)
    double base = (double)(p.blksize - 1) * p.step + p.rem;
    st.remCodePhase = base + p.step - g.L;
    if (pre) st.remCarrPhase = pre[4];
    else {
        double trig = ((p.carrFreq * 2.0 * 3.14159265358979323846) * ((double)p.blksize / g.fs)) + p.remCarr;
        st.remCarrPhase = fmod(trig, twopi);  // rem(trigarg(blksize+1), 2*pi), WB:337
    }
    st.pos = p.pos + p.blksize;
    st.samples += p.blksize;

    double I_E = s[sum_idx(0, EPL_E, 0)], Q_E = s[sum_idx(0, EPL_E, 1)];
    double I_P = s[sum_idx(0, EPL_P, 0)], Q_P = s[sum_idx(0, EPL_P, 1)];
    double I_L = s[sum_idx(0, EPL_L, 0)], Q_L = s[sum_idx(0, EPL_L, 1)];
    double carrError = pre ? pre[0] : atan(Q_P / I_P) / twopi;
    double codeError = pre ? pre[2] : dll_disc(I_E, Q_E, I_L, Q_L);
    if (b1c) codeError = codeError * (1.0 - g.d);
#pragma unroll
    for (int o = 0; o < 3; ++o) a.pI[o] = a.pQ[o] = 0.0;
    if (g.mode == BDS_TRK_B1C_WB && g.hasPilot) {
        const double ka = sqrt(4.0 / 33.0), kb = sqrt(29.0 / 33.0);
#pragma unroll
        for (int o = 0; o < 3; ++o) {  // WB:375-380
            a.pI[o] = -ka * s[sum_idx(2, o, 0)] + kb * s[sum_idx(1, o, 1)];
            a.pQ[o] = -ka * s[sum_idx(2, o, 1)] - kb * s[sum_idx(1, o, 0)];
        }
        double pe = pre ? pre[1] : atan(a.pQ[EPL_P] / a.pI[EPL_P]) / twopi;
        carrError = (carrError * 1 + pe * 3) / 4;
        double pc = (pre ? pre[3] : dll_disc(a.pI[EPL_E], a.pQ[EPL_E], a.pI[EPL_L], a.pQ[EPL_L])) * (1.0 - g.d);
        codeError = codeError * g.factor + pc * (1.0 - g.factor);
    } else if (g.hasPilot) {
        double pIP = s[sum_idx(1, EPL_P, 0)], pQP = s[sum_idx(1, EPL_P, 1)];
        double pc = dll_disc(s[sum_idx(1, EPL_E, 0)], s[sum_idx(1, EPL_E, 1)], s[sum_idx(1, EPL_L, 0)],
                             s[sum_idx(1, EPL_L, 1)]);
        if (pre) pc = pre[3];
        if (g.mode == BDS_TRK_B1C_NB) {
            double pe = pre ? pre[1] : atan(-pIP / pQP) / twopi;   // NB:357
            carrError = (carrError * 11 + pe * 29) / 40;
            codeError = (codeError * 11 + pc * (1.0 - g.d) * 29) / 40;
        } else {
            const double cr = 6.123233995736766e-17;  // cos(pi/2) in double, B2a:345
            double re = pIP * cr + pQP, im = pQP * cr - pIP;
            double pe = pre ? pre[1] : atan(im / re) / twopi;
            carrError = (carrError + pe) / 2;
            codeError = (codeError + pc) / 2;
        }
        a.pI[EPL_P] = pIP;
        a.pQ[EPL_P] = pQP;
    }
    // PLL filter WB:399-406
    st.d2CarrError = st.d2CarrError + carrError * g.pf3;
    st.dCarrError = st.d2CarrError + carrError * g.pf2 + st.dCarrError;
    double carrNco = st.dCarrError + carrError * g.pf1;
    a.carrFreqOld = st.carrFreq;
    st.carrFreq = st.carrFreqBasis + carrNco;
    // DLL filter WB:422-430
    double codeNco = st.oldCodeNco + g.tau2over1 * (codeError - st.oldCodeError) + codeError * g.PDIoverTau1;  // (tau2/tau1), (PDI/tau1)
    st.oldCodeNco = codeNco;
    st.oldCodeError = codeError;
    a.codeFreqOld = st.codeFreq;
    st.codeFreq = chCodeFreq - codeNco;
    a.carrError = carrError;
    a.codeError = codeError;
    a.carrNco = carrNco;
    a.codeNco = codeNco;
}

__device__ void close_out(const TrkDev& g, const double* s, const EpochParams& p, const CloseAux& a, double* outv) {
#pragma unroll
    for (int i = 0; i < kNFields; ++i) outv[i] = 0.0;
    outv[F_ABS] = (double)p.pos;           // WB:254
    outv[F_REMCODE] = p.rem;               // WB:287
    outv[F_REMCARR] = p.remCarr;           // WB:332
#pragma unroll
    for (int i = 0; i < kNSum; ++i) outv[F_RAW0 + i] = s[i];
    outv[F_PI_P] = a.pI[EPL_P];
    outv[F_PI_E] = a.pI[EPL_E];
    outv[F_PI_L] = a.pI[EPL_L];
    outv[F_PQ_P] = a.pQ[EPL_P];
    outv[F_PQ_E] = a.pQ[EPL_E];
    outv[F_PQ_L] = a.pQ[EPL_L];
    outv[F_CARRFREQ] = a.carrFreqOld;
    outv[F_CODEFREQ] = a.codeFreqOld;
    outv[F_DLL] = a.codeError;
    outv[F_DLLF] = a.codeNco;
    outv[F_PLL] = a.carrError;
    outv[F_PLLF] = a.carrNco;
    outv[F_I_E] = s[sum_idx(0, EPL_E, 0)];
    outv[F_I_P] = s[sum_idx(0, EPL_P, 0)];
    outv[F_I_L] = s[sum_idx(0, EPL_L, 0)];
    outv[F_Q_E] = s[sum_idx(0, EPL_E, 1)];
    outv[F_Q_P] = s[sum_idx(0, EPL_P, 1)];
    outv[F_Q_L] = s[sum_idx(0, EPL_L, 1)];
}

__device__ void close_core(const TrkDev& g, const double* s, const EpochParams& p, double chCodeFreq, ChanState& st,
                           double* outv, const double* pre = nullptr) {
    CloseAux a;
    close_nco(g, s, p, chCodeFreq, st, a, pre);
    close_out(g, s, p, a, outv);
}

// planes a mode does not produce keep their preallocation values (NB / B2a have no E/L pilot planes, data-only
// modes no pilot planes at all)
__device__ __forceinline__ bool field_written(const TrkDev& g, int f) {
    if (f >= F_PI_P && f <= F_PQ_L) {
        if (!g.hasPilot) return false;
        if (g.mode != BDS_TRK_B1C_WB) return f == F_PI_P || f == F_PQ_P;
    }
    return true;
}

// Lock-loss status and early channel drop (an extension: the reference copies channel.status
// unconditionally, WB_tracking.m:485-488).  Off unless cfg.lockLossPLD > 0.  Evaluated at the end of every C/N0
// interval on the lock detector of Calc_CNo_PLD.m:70-73 (the pilot's when the mode tracks a pilot, else the data
// component's): lockLossIntervals consecutive values below lockLossPLD drop the channel - it runs no further epoch.
__device__ __forceinline__ void lock_update(const TrkDev& g, ChanState& st, double pldData, double pldPilot, int e) {
    if (!(g.lockPLD > 0.0)) return;
    const double pld = g.hasPilot ? pldPilot : pldData;
    st.lowLock = pld < g.lockPLD ? st.lowLock + 1 : 0;    // NaN (0/0: no signal at all) counts as locked, like a comparison in MATLAB
    if (st.lowLock >= g.lockIntervals && st.lockLost == 0) st.lockLost = e + 1;
}

// C/N0 + lock detector every CNoInterval epochs (WB:459-481), single thread
__device__ void close_cno(const TrkDev& g, int c, int e, ChanState& st) {
    if (!(g.cnoInterval > 0 && (e + 1) % g.cnoInterval == 0)) return;
    const int ci = (e + 1) / g.cnoInterval - 1;
    if (ci >= g.cnoCap) return;
    const int cap = g.capacity;
    const double* out = g.out + (size_t)c * kNFields * cap;
    const int n = g.cnoInterval, e0 = e + 1 - n;
    double* cn = g.cno + (size_t)c * kNCno * g.cnoCap;
    double d, dp, pv = 0, pp = 0;
    cno_pld(out + F_I_P * cap + e0, out + F_Q_P * cap + e0, n, g.PDI, d, dp);
    double c0 = 10.0 * log10(d), c1 = 0;
    if (g.hasPilot) {
        if (g.mode == BDS_TRK_B1C_WB)
            cno_pld(out + F_PI_P * cap + e0, out + F_PQ_P * cap + e0, n, g.PDI, pv, pp);
        else  // NB / B2a swap pilot I and Q (Calc_CNo_PLD.m:80-88)
            cno_pld(out + F_PQ_P * cap + e0, out + F_PI_P * cap + e0, n, g.PDI, pv, pp);
        c1 = 10.0 * log10(pv);
    }
    double c2 = 10.0 * log10(d + pv);
    cn[0 * g.cnoCap + ci] = c0 * 0.5 + st.cnoPrev[0] * 0.5;
    cn[1 * g.cnoCap + ci] = dp;
    if (g.hasPilot) {
        cn[2 * g.cnoCap + ci] = c1 * 0.5 + st.cnoPrev[1] * 0.5;
        cn[3 * g.cnoCap + ci] = pp;
        cn[4 * g.cnoCap + ci] = c2 * 0.5 + st.cnoPrev[2] * 0.5;
    }
    st.cnoPrev[0] = c0;
    st.cnoPrev[1] = c1;
    st.cnoPrev[2] = c2;
    lock_update(g, st, dp, pp, e);
}

// Single-thread closure used by the general kernel: loads state, closes the loops, stores outputs + state and
// stages the next epoch's params (published by the CTA, publish_next).
__device__ void close_epoch(const TrkDev& g, int c, int e, const double* s /*18 sums*/, EpochParams& npOut, int& npOk) {
    ChanState st = load_cg(g.st + c);
    const EpochParams p = load_cg(g.params + c * 2 + (e & 1));
    double* out = g.out + (size_t)c * kNFields * g.capacity;
    const int cap = g.capacity;
    double outv[kNFields];
    close_core(g, s, p, g.cc[c].chCodeFreq, st, outv);
#pragma unroll
    for (int f = 0; f < kNFields; ++f)
        if (field_written(g, f)) out[(size_t)f * cap + e] = outv[f];
    __threadfence();
    close_cno(g, c, e, st);
    st.epoch = e + 1;
    store_cg(g.st + c, st);

    // stage the next epoch's params; the CTA publishes them (publish_next)
    EpochParams np;
    bool ok = next_params(g, st, np) && e + 1 < g.epochLimit;
    if (!ok && e + 1 < g.epochLimit && st.lockLost == 0) out[F_ABS * cap + e + 1] = (double)st.pos;  // WB_tracking.m:254 precedes the failed read
    npOut = np;
    npOk = ok;
}

// ======================================================================================
// Persistent closed-loop kernel
// ======================================================================================
struct __align__(16) TrkSmem {
    uint32_t bits[2][kPackedWordsDev];
    double red[8 * kNSum];
    double sums[kNSum];
    double part[kNSum * 14];
    EpochParams p;
    EpochParams np;      // next epoch's params, staged between loop closure and publication
    unsigned scratch[128];
    int go;
    int last;
    int curChan;
    int npOk;
};

__device__ void load_code_bits(const TrkDev& g, TrkSmem& sm, int c) {
    if (sm.curChan == c) return;
    __syncthreads();
    const uint32_t* src = g.codeBits + (size_t)c * 2 * kPackedWordsDev;
    for (int i = threadIdx.x; i < 2 * kPackedWordsDev; i += blockDim.x) (&sm.bits[0][0])[i] = __ldg(src + i);
    __syncthreads();
    if (threadIdx.x == 0) sm.curChan = c;
    __syncthreads();
}

// Sum the S slice partials of channel c in a fixed order (deterministic).
__device__ void reduce_partials(const TrkDev& g, TrkSmem& sm, int c) {
    const double* part = g.partial + (size_t)c * g.S * kNSum;
    int t = threadIdx.x;
    if (t < kNSum * 14) {
        int l = t / 14, j = t % 14;
        double a = 0;
        for (int s = j; s < g.S; s += 14) a += __ldcg(part + (size_t)s * kNSum + l);
        sm.part[l * 14 + j] = a;
    }
    __syncthreads();
    if (t < kNSum) {
        double a = 0;
#pragma unroll
        for (int j = 0; j < 14; ++j) a += sm.part[t * 14 + j];
        sm.sums[t] = a;
    }
    __syncthreads();
}

__device__ __forceinline__ void correlate_slice(const TrkDev& g, TrkSmem& sm, const EpochParams& p, int sl, int S,
                                                float* acc) {
    if (g.counters && threadIdx.x == 0) atomicAdd(g.counters + 2, 1ull);
    SliceCtx ctx{g.x, g.winFirst, g.winLen, p, g.mode, g.hasPilot, g.hasP61, g.L, g.d, g.fs, g.iq};
    long long q0, q1;
    slice_chunks(p, g.winFirst, g.iq, sl, S, q0, q1);
    correlate_general(ctx, sm.bits[0], sm.bits[1], q0, q1, acc);
}

// Publishes the params staged in sm.np or stops the channel.  Whole CTA.
__device__ void publish_next(const TrkDev& g, TrkSmem& sm, int c, int eNext) {
    __syncthreads();
    const int ok = sm.npOk;
    if (ok && threadIdx.x == 0) store_cg(g.params + c * 2 + (eNext & 1), sm.np);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (ok) st_release(g.ready + c, eNext);
        else st_release(g.stop + c, eNext);
    }
}

// Persistent closed-loop kernel (general correlator).  Work items t = blockIdx.x, blockIdx.x +
// gridDim.x, ... in the global order (round, channel, slice); the item's channel-epoch must have
// been published.
__global__ void __launch_bounds__(kTrkThreads) trk_persistent_kernel(TrkDev g) {
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    TrkSmem& sm = *reinterpret_cast<TrkSmem*>(dyn_smem);
    if (threadIdx.x == 0) sm.curChan = -1;
    __syncthreads();
    const long long perRound = (long long)g.nAct * g.S;
    const long long total = perRound * g.maxEpochs;
    for (long long t = blockIdx.x; t < total; t += gridDim.x) {
        const int i = (int)(t / perRound);
        const int idx = (int)(t - (long long)i * perRound);
        const int c = g.act[idx / g.S], sl = idx % g.S;
        const int e = g.cc[c].pad + i;   // channels advance from their own completed-epoch count
        if (threadIdx.x == 0) {
            int go = 0;
            while (true) {
                if (ld_acquire(g.stop + c) <= e) break;
                if (ld_acquire(g.ready + c) >= e) {
                    go = 1;
                    break;
                }
                __nanosleep(32);
            }
            sm.go = go;
            if (go) sm.p = load_cg(g.params + c * 2 + (e & 1));
        }
        __syncthreads();
        if (!sm.go) {
            __syncthreads();
            continue;
        }
        load_code_bits(g, sm, c);
        const EpochParams p = sm.p;
        float acc[kNSum];
#pragma unroll
        for (int k = 0; k < kNSum; ++k) acc[k] = 0.f;
        correlate_slice(g, sm, p, sl, g.S, acc);
        block_reduce18(acc, sm.red, sm.sums);
        if (threadIdx.x < kNSum) {
            g.partial[((size_t)c * g.S + sl) * kNSum + threadIdx.x] = sm.sums[threadIdx.x];
            __threadfence();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int old = atomicAdd(g.count + c, 1);
            sm.last = (old == g.S - 1);
            if (sm.last) g.count[c] = 0;
        }
        __syncthreads();
        if (sm.last) {
            __threadfence();
            reduce_partials(g, sm, c);
            if (threadIdx.x == 0) close_epoch(g, c, e, sm.sums, sm.np, sm.npOk);
            publish_next(g, sm, c, e + 1);
        }
        __syncthreads();
    }
}

// Computes the first params of every channel for the current window (run start): one CTA per channel.
__global__ void __launch_bounds__(kTrkThreads) trk_prepare_kernel(TrkDev g) {
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    TrkSmem& sm = *reinterpret_cast<TrkSmem*>(dyn_smem);
    const int c = blockIdx.x;
    if (!g.cc[c].active) {
        if (threadIdx.x == 0) {
            g.count[c] = 0;
            g.stop[c] = 0;
            g.ready[c] = -1;
        }
        return;
    }
    if (threadIdx.x == 0) {
        g.count[c] = 0;
        ChanState st = g.st[c];
        g.cc[c].pad = st.epoch;
        sm.npOk = next_params(g, st, sm.np) && st.epoch < g.epochLimit;
        if (!sm.npOk && st.epoch < g.epochLimit && st.lockLost == 0)
            g.out[((size_t)c * kNFields + F_ABS) * g.capacity + st.epoch] = (double)st.pos;
        g.ready[c] = st.epoch - 1;
        g.stop[c] = INT_MAX;
        sm.go = st.epoch;
    }
    __syncthreads();
    publish_next(g, sm, c, sm.go);
}

}  // namespace bds
#include "bds_track_fw.cuh"       // geometry g99: fs = 99.375 MHz (BASELINE)
#undef FAST_GEOM_NS
#undef FAST_GEN_INC
#define FAST_GEOM_NS g53          // geometry g53: fs = 53 MHz, the reference's shipped B1C setting (B1C/initSettings.m:57)
#define FAST_GEN_INC "bds_track_fast_gen_53.inc"
#include "bds_track_fast.cuh"
#include "bds_track_fw.cuh"
#undef FAST_GEOM_NS
#undef FAST_GEN_INC
#define FAST_GEOM_NS g99n         // narrow-band bodies (NB_tracking.m: data + BOC(1,1) pilot, six segments per chip)
#define FAST_GEN_INC "bds_track_fast_gen_nb.inc"
#include "bds_track_fast.cuh"
#include "bds_track_fw.cuh"
#undef FAST_GEOM_NS
#undef FAST_GEN_INC
#define FAST_GEOM_NS g53n
#define FAST_GEN_INC "bds_track_fast_gen_53_nb.inc"
#include "bds_track_fast.cuh"
#include "bds_track_fw.cuh"
#include "bds_track_b2a.cuh"
namespace bds {
// the instantiations of the chip-synchronous B1C kernel
struct FwGeom {
    const char* name;
    bool (*supported)(int mode, int hasPilot, int hasP61, double fs, double fc, int codeLength, double d);
    void (*kernel)(TrkDev);
    size_t smemBytes;
};
static const FwGeom kFwGeoms[] = {   // first match wins: the narrow-band bodies come before the wide-band ones, which also serve NB
    {"fs = 99.375 MHz, narrow band", g99n::fast_wb_supported, g99n::trk_fw_kernel, sizeof(g99n::FwSmem)},
    {"fs = 53 MHz, narrow band", g53n::fast_wb_supported, g53n::trk_fw_kernel, sizeof(g53n::FwSmem)},
    {"fs = 99.375 MHz", g99::fast_wb_supported, g99::trk_fw_kernel, sizeof(g99::FwSmem)},
    {"fs = 53 MHz", g53::fast_wb_supported, g53::trk_fw_kernel, sizeof(g53::FwSmem)},
};
constexpr int kFwGeomFirstWide = 2;   // cfg.reserved bit 1 (test hook): narrow band on the wide-band body, as before the NB bodies existed
constexpr int kFwGeomCount = (int)(sizeof(kFwGeoms) / sizeof(kFwGeoms[0]));
}  // namespace bds
namespace bds {

// ======================================================================================
// Open-loop ("teacher forced") correlator: grid = (S, n_ch * n_epochs)
// ======================================================================================
__global__ void __launch_bounds__(kTrkThreads) trk_open_loop_kernel(TrkDev g, const EpochParams* params, int nEpochs,
                                                                   double* partial /*[ce][S][18]*/) {
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    TrkSmem& sm = *reinterpret_cast<TrkSmem*>(dyn_smem);
    const int ce = blockIdx.y, sl = blockIdx.x, S = gridDim.x;
    const int c = ce / nEpochs;
    if (threadIdx.x == 0) sm.curChan = -1;
    __syncthreads();
    load_code_bits(g, sm, c);
    const EpochParams p = params[ce];
    float acc[kNSum];
#pragma unroll
    for (int k = 0; k < kNSum; ++k) acc[k] = 0.f;
    correlate_slice(g, sm, p, sl, S, acc);
    block_reduce18(acc, sm.red, sm.sums);
    if (threadIdx.x < kNSum) partial[((size_t)ce * S + sl) * kNSum + threadIdx.x] = sm.sums[threadIdx.x];
}

__global__ void trk_open_loop_reduce_kernel(const double* partial, int S, int nce, double* sums) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nce * kNSum) return;
    int ce = i / kNSum, l = i % kNSum;
    double a = 0;
    for (int s = 0; s < S; ++s) a += partial[((size_t)ce * S + s) * kNSum + l];
    sums[i] = a;
}

}  // namespace bds

// ======================================================================================
// Host side: session object + C ABI
// ======================================================================================
using namespace bds;

struct bds_trk {
    int mode = 0;
    bds_trk_cfg cfg{};
    int nCh = 0;
    std::vector<bds_channel> ch;
    long long skip = 0;
    // IF window
    int8_t* dX = nullptr;
    bool ownX = false;
    size_t xCap = 0;
    long long winFirst = 0, winLen = 0;
    // device tables
    uint32_t* dBits = nullptr;
    ChanConst* dCC = nullptr;
    ChanState* dSt = nullptr;
    EpochParams* dParams = nullptr;
    int *dReady = nullptr, *dStop = nullptr, *dCount = nullptr;
    double* dPartial = nullptr;
    double* dAcc = nullptr;
    double* dOut = nullptr;
    double* dCno = nullptr;
    int capacity = 0, cnoCap = 0;
    int S = 0, nAct = 0, gridBlocks = 0, nCompute = 0;
    int* dAct = nullptr;
    unsigned long long* dCounters = nullptr;
    unsigned long long* dQueue = nullptr;
    unsigned* dQctl = nullptr;
    unsigned qSize = 0;
    unsigned long long* dTrace = nullptr;
    unsigned traceCap = 0;
    bool fast = false;
    int geom = 0;           // index into kFwGeoms when fast
    bool b2aUnit = false;   // B2a on the per-channel chip-synchronous kernel
    int b2aCluster = 1;     // ... with this many CTAs (a thread-block cluster) per channel
    int iq = 0;             // 1: the record holds interleaved I/Q int8 pairs (cfg.fileType == 2); window quantities are samples
    size_t smemBytes = 0;
    int epochsRun = 0;  // max over channels, as seen by the host
    cudaStream_t stream = nullptr, copyStream = nullptr;
    std::vector<cudaEvent_t> chunkEv;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float lastMs = 0.f;
    bool pending = false;
    std::vector<ChanState> hSt;
    // mmap'd file
    void* map = nullptr;
    size_t mapLen = 0;
    int fd = -1;                    // the same file, kept open: its bytes reach the device through pinned staging buffers
    // pinned staging ring for pageable sources (files): filled by parallel pread()s, drained by the copy engine
    std::vector<int8_t*> stage;
    std::vector<cudaEvent_t> stageEv;
    size_t stageNext = 0;
};

namespace {

unsigned long long g_open_loop_counters[4] = {0, 0, 0, 0};
unsigned long long g_open_loop_us = 0;   // device time of the last open-loop chip-synchronous kernel launch

bool mode_flags(int mode, int flag, int& hasPilot, int& hasP61) {
    hasPilot = (mode == BDS_TRK_B1C_WB && flag == 2) || (mode != BDS_TRK_B1C_WB && flag == 1);
    hasP61 = (mode == BDS_TRK_B1C_WB && flag == 2);
    return mode == BDS_TRK_B1C_WB || mode == BDS_TRK_B1C_NB || mode == BDS_TRK_B2A;
}

int build_code_bits(int mode, const int32_t* prn, int nCh, std::vector<uint32_t>& bits) {
    bits.assign((size_t)nCh * 2 * kPackedWords, 0);
    std::vector<uint8_t> chips;
    for (int c = 0; c < nCh; ++c) {
        if (prn[c] == 0) continue;
        if (prn[c] < 1 || prn[c] > 63) return set_error(BDS_ERR_ARG, "channel %d: PRN %d out of range", c, prn[c]);
        int cd = mode == BDS_TRK_B2A ? BDS_CODE_B2A_DATA : BDS_CODE_B1C_DATA_PRIMARY;
        int cp = mode == BDS_TRK_B2A ? BDS_CODE_B2A_PILOT : BDS_CODE_B1C_PILOT_PRIMARY;
        primary_bits(cd, prn[c], chips);
        pack_bits(chips, &bits[((size_t)c * 2 + 0) * kPackedWords]);
        primary_bits(cp, prn[c], chips);
        pack_bits(chips, &bits[((size_t)c * 2 + 1) * kPackedWords]);
    }
    return BDS_OK;
}

void fill_dev(const bds_trk* h, TrkDev& g, int maxEpochs) {
    std::memset(&g, 0, sizeof(g));
    g.x = h->dX;
    g.winFirst = h->winFirst;
    g.winLen = h->winLen;
    g.winStage = h->ownX ? ((h->winLen + 16) & ~15LL) : (h->winLen & ~15LL);
    g.mode = h->mode;
    mode_flags(h->mode, h->cfg.pilotTRKflag, g.hasPilot, g.hasP61);
    g.nCh = h->nCh;
    g.S = h->S;
    g.nAct = h->nAct;
    g.act = h->dAct;
    g.counters = h->dCounters;
    g.queue = h->dQueue;
    g.qctl = h->dQctl;
    g.qMask = h->qSize ? h->qSize - 1 : 0;
    g.pubTime = (h->cfg.debug & BDS_DBG_TIMING) ? (unsigned long long*)(h->dCounters + 32) : nullptr;
    g.trace = h->dTrace;
    g.traceCap = h->traceCap;
    g.nCompute = h->nCompute;
    g.stages = kFwStages;
    g.tune = 0;
    // passes a producer may have in flight before it takes its next task: deep prefetch when the queue has a backlog
    // (many channels), none when tasks are scarce (a ready task must not wait behind a busy CTA)
    const long long tasksPerRound = (long long)h->nAct * h->S;
    g.ahead = tasksPerRound >= 3LL * h->gridBlocks ? 2 : (2 * tasksPerRound >= 3LL * h->gridBlocks ? 1 : 0);
    if (h->cfg.fwPrefetch > 0) g.ahead = h->cfg.fwPrefetch - 1;   // caller's tuning (bds_trk_cfg)
    g.ahead = std::max(0, std::min(g.ahead, g.stages - 1));   // more passes in flight than stages would deadlock the producer
    g.maxEpochs = maxEpochs;
    g.capacity = h->capacity;
    g.epochLimit = h->capacity;
    g.cnoCap = h->cnoCap;
    g.cnoInterval = h->cfg.CNoInterval;
    g.kernelKind = h->fast ? BDS_KERNEL_FAST : BDS_KERNEL_GENERAL;
    g.pad = h->cfg.reserved & 1;  // test hook: wide guard band in the fast kernel
    g.fs = h->cfg.samplingFreq;
    g.L = (double)h->cfg.codeLength;
    g.d = h->cfg.dllCorrelatorSpacing;
    g.PDI = h->cfg.intTime;
    g.tau1 = h->cfg.tau1code;
    g.tau2 = h->cfg.tau2code;
    g.tau2over1 = h->cfg.tau2code / h->cfg.tau1code;   // same IEEE divisions as WB_tracking.m:422-424, hoisted
    g.PDIoverTau1 = h->cfg.intTime / h->cfg.tau1code;
    g.lockPLD = h->cfg.lockLossPLD;
    g.lockIntervals = std::max(1, (int)h->cfg.lockLossIntervals);
    g.iq = h->iq;
    g.pf1 = h->cfg.pf1;
    g.pf2 = h->cfg.pf2;
    g.pf3 = h->cfg.pf3;
    g.factor = h->cfg.wbFactor;
    g.codeBits = h->dBits;
    g.cc = h->dCC;
    g.st = h->dSt;
    g.params = h->dParams;
    g.ready = h->dReady;
    g.stop = h->dStop;
    g.count = h->dCount;
    g.partial = h->dPartial;
    g.acc = h->dAcc;
    g.out = h->dOut;
    g.cno = h->dCno;
}

int choose_fast(int mode, const bds_trk_cfg* cfg, bool& fast, int& geom) {
    int hasPilot, hasP61;
    mode_flags(mode, cfg->pilotTRKflag, hasPilot, hasP61);
    if (cfg->fileType != 0 && cfg->fileType != 1 && cfg->fileType != 2)
        return set_error(BDS_ERR_ARG, "fileType must be 1 (real) or 2 (I/Q), got %d", cfg->fileType);
    const bool iq = cfg->fileType == 2;   // the chip-synchronous bodies pack real samples: I/Q records take the general kernel
    bool can = false;
    geom = 0;
    for (int k = (cfg->reserved & 2) ? kFwGeomFirstWide : 0; k < kFwGeomCount && !iq && !can; ++k)
        if (kFwGeoms[k].supported(mode, hasPilot, hasP61, cfg->samplingFreq, cfg->codeFreqBasis, cfg->codeLength,
                                  cfg->dllCorrelatorSpacing)) {
            can = true;
            geom = k;
        }
    if (cfg->kernel == BDS_KERNEL_FAST && !can &&
        (iq || !fastb_supported(mode, cfg->samplingFreq, cfg->codeFreqBasis, cfg->codeLength, cfg->dllCorrelatorSpacing)))
        return set_error(BDS_ERR_UNSUPPORTED, "fast tracking kernel does not support this configuration");
    fast = can && cfg->kernel != BDS_KERNEL_GENERAL;   // B1C chip-synchronous kernel; B2a: see b2a_unit_enabled
    return BDS_OK;
}

// B2a at the supported configuration runs on the per-channel chip-synchronous kernel (validated on hardware against the
// oracle: tests/test_gpu_b2a_unit.py); BDS_KERNEL_GENERAL selects the exact general kernel.
bool b2a_unit_enabled(int mode, const bds_trk_cfg* cfg) {
    return cfg->kernel != BDS_KERNEL_GENERAL && cfg->fileType != 2 &&
           fastb_supported(mode, cfg->samplingFreq, cfg->codeFreqBasis, cfg->codeLength, cfg->dllCorrelatorSpacing);
}

size_t smem_bytes(bool fast, int geom) {
    return fast ? std::max(kFwGeoms[geom].smemBytes, kFwCloseBase + sizeof(FwCloseScratch) * (kFwThreads / 32)) : sizeof(TrkSmem);
}

int init_state(bds_trk* h) {
    std::vector<ChanConst> cc(h->nCh);
    h->hSt.assign(h->nCh, ChanState{});
    for (int c = 0; c < h->nCh; ++c) {
        const bds_channel& ch = h->ch[c];
        cc[c].prn = ch.PRN;
        cc[c].active = ch.PRN != 0;
        cc[c].status = ch.status;
        cc[c].pad = 0;
        cc[c].chCodeFreq = ch.codeFreq;
        cc[c].acquiredFreq = ch.acquiredFreq;
        cc[c].startPos = h->skip + (long long)ch.codePhase - 1;  // WB_tracking.m:174-176
        ChanState& s = h->hSt[c];
        s.codeFreq = ch.codeFreq;           // WB:195-206
        s.carrFreq = ch.acquiredFreq;
        s.carrFreqBasis = ch.acquiredFreq;
        s.pos = cc[c].startPos;
    }
    BDS_CUDA(cudaMemcpy(h->dCC, cc.data(), sizeof(ChanConst) * h->nCh, cudaMemcpyHostToDevice));
    BDS_CUDA(cudaMemcpy(h->dSt, h->hSt.data(), sizeof(ChanState) * h->nCh, cudaMemcpyHostToDevice));
    h->epochsRun = 0;
    return BDS_OK;
}

// initial values exactly as the reference preallocates them (WB_tracking.m:53-112): Inf for the
// frequency / discriminator / remainder planes, 0 elsewhere.  Epochs [e0, cap) of every channel.
__global__ void init_out_kernel(double* out, int nCh, int cap, int e0) {
    const size_t per = (size_t)(cap - e0);
    const size_t total = (size_t)nCh * kNFields * per;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t row = i / per;
        int f = (int)(row % kNFields);
        bool isInf = f == F_CODEFREQ || f == F_CARRFREQ || f == F_DLL || f == F_DLLF || f == F_PLL || f == F_PLLF ||
                     f == F_REMCODE || f == F_REMCARR;
        out[row * cap + e0 + (i - row * per)] = isInf ? INFINITY : 0.0;
    }
}

int init_out_block(bds_trk* h, double* out, int cap, int e0) {
    if (cap <= e0) return BDS_OK;
    init_out_kernel<<<g_num_sms * 4, 256, 0, h->stream>>>(out, h->nCh, cap, e0);
    count_launch();
    BDS_CUDA(cudaGetLastError());
    return BDS_OK;
}

int ensure_capacity(bds_trk* h, int need) {
    if (need <= h->capacity) return BDS_OK;
    int ncap = std::max(need, h->capacity * 2);
    ncap = (ncap + 7) & ~7;
    double* nout = nullptr;
    BDS_CUDA(cudaMalloc(&nout, sizeof(double) * (size_t)h->nCh * kNFields * ncap));
    int ci = std::max(1, h->cfg.CNoInterval);
    int ncno = std::max(1, ncap / ci);
    double* ncn = nullptr;
    BDS_CUDA(cudaMalloc(&ncn, sizeof(double) * (size_t)h->nCh * kNCno * ncno));
    BDS_CUDA(cudaMemsetAsync(ncn, 0, sizeof(double) * (size_t)h->nCh * kNCno * ncno, h->stream));
    int old = h->capacity;
    if (old > 0) {
        BDS_CUDA(cudaMemcpy2DAsync(nout, sizeof(double) * ncap, h->dOut, sizeof(double) * old, sizeof(double) * old,
                                   (size_t)h->nCh * kNFields, cudaMemcpyDeviceToDevice, h->stream));
        BDS_CUDA(cudaMemcpy2DAsync(ncn, sizeof(double) * ncno, h->dCno, sizeof(double) * h->cnoCap,
                                   sizeof(double) * h->cnoCap, (size_t)h->nCh * kNCno, cudaMemcpyDeviceToDevice,
                                   h->stream));
    }
    int rc = init_out_block(h, nout, ncap, old);
    if (rc) return rc;
    BDS_CUDA(cudaStreamSynchronize(h->stream));
    if (h->dOut) cudaFree(h->dOut);
    if (h->dCno) cudaFree(h->dCno);
    h->dOut = nout;
    h->dCno = ncn;
    h->capacity = ncap;
    h->cnoCap = ncno;
    return BDS_OK;
}

// grid geometry: a cooperative grid of co-resident CTAs; S slices per channel-epoch.
int plan_grid(bds_trk* h) {
    int occ = 0;
    h->smemBytes = smem_bytes(h->fast, h->geom);
    int nAct = 0;
    for (auto& c : h->ch) nAct += c.PRN != 0;
    h->nAct = std::max(nAct, 1);
    if (h->b2aUnit) {
        h->smemBytes = sizeof(B2aSmem);
        BDS_CUDA(cudaFuncSetAttribute(trk_b2a_unit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smemBytes));
        BDS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, trk_b2a_unit_kernel, kB2aThreadsAll, h->smemBytes));
        if (occ < 1) return set_error(BDS_ERR_CUDA, "B2a tracking kernel does not fit on an SM");
        // CTAs per channel (one thread-block cluster): the loop closure is a serial chain of 1 000 steps per second of
        // signal, so with few channels per GPU the idle SMs shorten each step - the largest cluster for which every
        // channel's cluster is resident at once (a cluster of 8 needs 8 free SMs in one GPC)
        int cs = h->cfg.b2aClusterSize;
        if (cs != 0 && cs != 1 && cs != 2 && cs != 4 && cs != 8)
            return set_error(BDS_ERR_ARG, "b2aClusterSize must be 0 (auto), 1, 2, 4 or 8, got %d", cs);
        const bool autoCs = cs == 0;
        if (autoCs) cs = kB2aMaxCluster;
        for (; cs > 1; cs >>= 1) {
            if ((long long)h->nAct * cs > g_num_sms) continue;
            cudaLaunchConfig_t lc{};
            lc.gridDim = dim3(h->nAct * cs);
            lc.blockDim = dim3(kB2aThreadsAll);
            lc.dynamicSmemBytes = h->smemBytes;
            cudaLaunchAttribute at{};
            at.id = cudaLaunchAttributeClusterDimension;
            at.val.clusterDim.x = cs;
            at.val.clusterDim.y = at.val.clusterDim.z = 1;
            lc.attrs = &at;
            lc.numAttrs = 1;
            int nClusters = 0;
            if (cudaOccupancyMaxActiveClusters(&nClusters, trk_b2a_unit_kernel, &lc) == cudaSuccess && nClusters >= h->nAct) break;
            cudaGetLastError();
            if (!autoCs) return set_error(BDS_ERR_UNSUPPORTED, "b2aClusterSize %d: only %d clusters fit at once, %d channels", cs, nClusters, h->nAct);
        }
        h->b2aCluster = std::max(1, cs);
        h->gridBlocks = h->nAct * h->b2aCluster;   // one cluster per channel, all resident at once
        h->S = 1;
    } else if (h->fast) {
        BDS_CUDA(cudaFuncSetAttribute(kFwGeoms[h->geom].kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smemBytes));
        BDS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kFwGeoms[h->geom].kernel, kFwThreads, h->smemBytes));
        if (occ < 1) return set_error(BDS_ERR_CUDA, "tracking kernel does not fit on an SM");
        h->gridBlocks = g_num_sms;
        if (h->cfg.fwMaxCtas > 0) h->gridBlocks = std::max(2, std::min(g_num_sms, (int)h->cfg.fwMaxCtas));
        // a few CTAs only close loops: one warp per channel (16 warps per CTA)
        int nCloser = (h->nAct + (kFwThreads / 32) - 1) / (kFwThreads / 32);
        nCloser = std::min(nCloser, std::max(1, h->gridBlocks / 16));   // many channels: several channels per closer warp
        if ((long long)nCloser * (kFwThreads / 32) * 8 < h->nAct)
            return set_error(BDS_ERR_UNSUPPORTED, "chip-synchronous kernel: too many channels for the closer warps");
        h->nCompute = std::max(1, h->gridBlocks - nCloser);
        // slices of kFwChips*k chips (one chip per compute thread and pass); >= ~4 work items per CTA and round
        int k = 4;
        while (k > 1 && (long long)h->nAct * ((10230 + kFwChips * k - 1) / (kFwChips * k)) < 4LL * h->gridBlocks) --k;
        if (h->cfg.fwPassesPerTask > 0) k = std::max(1, std::min(8, (int)h->cfg.fwPassesPerTask));  // caller's tuning
        h->S = (10230 + kFwChips * k - 1) / (kFwChips * k);
        // queue entry: 10 bits of channel, 6 bits of slice, 19 bits of epoch (fw_put); a closer warp owns up to 8 channels
        if (h->nCh > kFwMaxChannels || h->S > 63)
            return set_error(BDS_ERR_UNSUPPORTED, "chip-synchronous kernel: at most %d channels per session (got %d); "
                                                  "open several sessions or use BDS_KERNEL_GENERAL", kFwMaxChannels, h->nCh);
    } else {
        BDS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, trk_persistent_kernel, kTrkThreads, h->smemBytes));
        if (occ < 1) return set_error(BDS_ERR_CUDA, "tracking kernel does not fit on an SM");
        occ = std::min(occ, 4);
        h->gridBlocks = g_num_sms * occ;
        int S = (3 * h->gridBlocks + h->nAct - 1) / h->nAct;
        h->S = std::max(4, std::min(S, 512));
    }
    return BDS_OK;
}

int open_common(int mode, const bds_trk_cfg* cfg, long long skip, const bds_channel* ch, int n_ch, bds_trk** out) {
    if (!cfg || !ch || !out || n_ch <= 0) return set_error(BDS_ERR_ARG, "bds_track_open: null/empty argument");
    int hp, h6;
    if (!mode_flags(mode, cfg->pilotTRKflag, hp, h6)) return set_error(BDS_ERR_ARG, "unknown tracking mode %d", mode);
    if (cfg->codeLength != kCodeLen) return set_error(BDS_ERR_UNSUPPORTED, "codeLength must be 10230");
    int rc = require_device();
    if (rc) return rc;
    bds_trk* h = new bds_trk();
    h->mode = mode;
    h->cfg = *cfg;
    h->nCh = n_ch;
    h->ch.assign(ch, ch + n_ch);
    h->skip = skip;
    rc = choose_fast(mode, cfg, h->fast, h->geom);
    if (rc) {
        delete h;
        return rc;
    }
    h->iq = cfg->fileType == 2;
    if (h->fast && cfg->kernel == BDS_KERNEL_AUTO && n_ch > kFwMaxChannels) h->fast = false;   // queue entries hold 10 bits of channel
    h->b2aUnit = !h->fast && b2a_unit_enabled(mode, cfg);
    auto fail = [&](int code) {
        bds_track_close(h);
        return code;
    };
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess)
        return fail(set_error(BDS_ERR_CUDA, "stream/event creation failed"));
    std::vector<int32_t> prn(n_ch);
    for (int c = 0; c < n_ch; ++c) prn[c] = ch[c].PRN;
    std::vector<uint32_t> bits;
    rc = build_code_bits(mode, prn.data(), n_ch, bits);
    if (rc) return fail(rc);
    rc = plan_grid(h);
    if (rc) return fail(rc);
#define TRY(x)                                                                                       \
    if ((x) != cudaSuccess) return fail(set_error(BDS_ERR_NOMEM, "device allocation failed: %s", #x))
    TRY(cudaMalloc(&h->dBits, bits.size() * 4));
    TRY(cudaMemcpy(h->dBits, bits.data(), bits.size() * 4, cudaMemcpyHostToDevice));
    TRY(cudaMalloc(&h->dCC, sizeof(ChanConst) * n_ch));
    TRY(cudaMalloc(&h->dSt, sizeof(ChanState) * n_ch));
    TRY(cudaMalloc(&h->dParams, sizeof(EpochParams) * n_ch * 2));
    TRY(cudaMalloc(&h->dReady, sizeof(int) * n_ch));
    TRY(cudaMalloc(&h->dStop, sizeof(int) * n_ch));
    TRY(cudaMalloc(&h->dCount, sizeof(int) * n_ch));
    TRY(cudaMalloc(&h->dPartial, sizeof(double) * (size_t)n_ch * h->S * kNSum));
    TRY(cudaMalloc(&h->dAcc, sizeof(double) * (size_t)n_ch * kNSum));
    TRY(cudaMemset(h->dAcc, 0, sizeof(double) * (size_t)n_ch * kNSum));
    TRY(cudaMalloc(&h->dAct, sizeof(int) * n_ch));
    TRY(cudaMalloc(&h->dCounters, 256 + 8 * 128));
    TRY(cudaMemset(h->dCounters, 0, 256 + 8 * 128));
    {
        std::vector<int> act;
        for (int c = 0; c < n_ch; ++c)
            if (ch[c].PRN != 0) act.push_back(c);
        act.resize(n_ch, 0);
        TRY(cudaMemcpy(h->dAct, act.data(), sizeof(int) * n_ch, cudaMemcpyHostToDevice));
    }
    if (h->fast) {
        unsigned need = (unsigned)(2 * n_ch * h->S + 2 * h->gridBlocks + 64);
        h->qSize = 1024;
        while (h->qSize < need) h->qSize <<= 1;
        TRY(cudaMalloc(&h->dQueue, sizeof(unsigned long long) * kFwQueueWords * h->qSize));   // 64-byte entries
        TRY(cudaMalloc(&h->dQctl, sizeof(unsigned) * kQWords));
        if ((cfg->debug & BDS_DBG_TRACE) && cfg->traceTickets > 0) {
            h->traceCap = (unsigned)cfg->traceTickets;
            TRY(cudaMalloc(&h->dTrace, sizeof(unsigned long long) * 8 * h->traceCap));
            TRY(cudaMemset(h->dTrace, 0, sizeof(unsigned long long) * 8 * h->traceCap));
        }
    }
#undef TRY
    rc = init_state(h);
    if (rc) return fail(rc);
    *out = h;
    return BDS_OK;
}

int set_window(bds_trk* h, const int8_t* x, size_t n, int x_loc, long long first) {
    if (!x && n) return set_error(BDS_ERR_ARG, "null IF buffer");
    if (x_loc == BDS_LOC_DEVICE) {
        if (((uintptr_t)x & 15) != 0) return set_error(BDS_ERR_ARG, "device IF buffer must be 16-byte aligned");
        if (h->ownX && h->dX) cudaFree(h->dX);
        h->dX = const_cast<int8_t*>(x);
        h->ownX = false;
        h->xCap = n << h->iq;   // used in place, all n samples: nothing is read past the last one (TrkDev::winStage)
    } else {
        const size_t nb = n << h->iq;   // bytes
        if (!h->ownX || h->xCap < nb + 64) {
            if (h->ownX && h->dX) cudaFree(h->dX);
            h->dX = nullptr;
            h->ownX = true;
            h->xCap = nb + 64;
            BDS_CUDA(cudaMalloc(&h->dX, h->xCap));
        }
        BDS_CUDA(cudaMemcpyAsync(h->dX, x, nb, cudaMemcpyHostToDevice, h->stream));
        BDS_CUDA(cudaMemsetAsync(h->dX + nb, 0, 64, h->stream));
    }
    h->winFirst = first;
    h->winLen = (long long)n;
    return BDS_OK;
}

}  // namespace

extern "C" {

int bds_track_open(int mode, const bds_trk_cfg* cfg, const int8_t* x, size_t n, int x_loc, long long skip,
                   const bds_channel* ch, int n_ch, bds_trk** out) {
    bds_trk* h = nullptr;
    int rc = open_common(mode, cfg, skip, ch, n_ch, &h);
    if (rc) return rc;
    rc = set_window(h, x, n, x_loc, 0);
    if (rc) {
        bds_track_close(h);
        return rc;
    }
    *out = h;
    return BDS_OK;
}

int bds_track_open_file(int mode, const bds_trk_cfg* cfg, const char* path, long long skip, long long max_samples,
                        const bds_channel* ch, int n_ch, bds_trk** out) {
    if (!path) return set_error(BDS_ERR_ARG, "null path");
    int fd = ::open(path, O_RDONLY);
    if (fd < 0) return set_error(BDS_ERR_IO, "cannot open %s", path);
    struct stat sb;
    if (fstat(fd, &sb) != 0 || sb.st_size <= 0) {
        ::close(fd);
        return set_error(BDS_ERR_IO, "cannot stat %s", path);
    }
    size_t len = (size_t)sb.st_size;
    const int iq = cfg && cfg->fileType == 2;   // fileType 2: two bytes per sample (postProcessing.m:67-71)
    if (max_samples > 0 && ((size_t)max_samples << iq) < len) len = (size_t)max_samples << iq;
    len &= ~(size_t)iq;
    void* m = mmap(nullptr, len, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) {
        ::close(fd);
        return set_error(BDS_ERR_IO, "mmap of %s failed", path);
    }
    madvise(m, len, MADV_SEQUENTIAL);
    bds_trk* h = nullptr;
    int rc = open_common(mode, cfg, skip, ch, n_ch, &h);
    if (rc == BDS_OK) rc = set_window(h, nullptr, 0, BDS_LOC_HOST, 0);
    if (rc) {
        munmap(m, len);
        ::close(fd);
        if (h) bds_track_close(h);
        return rc;
    }
    // the file is uploaded by the runs, streamed chunk by chunk under the tracking kernel
    h->map = m;
    h->mapLen = len;
    h->fd = fd;
    *out = h;
    return BDS_OK;
}

int bds_track_feed(bds_trk* h, const int8_t* x, size_t n, int x_loc, long long first_sample) {
    if (!h) return set_error(BDS_ERR_ARG, "null handle");
    return set_window(h, x, n, x_loc, first_sample);
}

int bds_track_run_streamed(bds_trk* h, const int8_t* x, size_t n, size_t chunk_bytes, int n_epochs);
static int run_streamed_from(bds_trk* h, const int8_t* x, size_t n, size_t chunk_bytes, int n_epochs, long long first,
                             long long fileOff = -1);

// One launch of the persistent kernel on the session's stream: up to maxEpochs more epochs per channel,
// no epoch index >= epochLimit, over the currently resident window.
static int launch_run(bds_trk* h, int maxEpochs, int epochLimit) {
    TrkDev g;
    fill_dev(h, g, maxEpochs);
    g.epochLimit = std::min(epochLimit, h->capacity);
    if (h->b2aUnit) {   // self-contained: every cluster starts from its channel's device-side state
        cudaLaunchConfig_t lc{};
        lc.gridDim = dim3(h->nAct * h->b2aCluster);
        lc.blockDim = dim3(kB2aThreadsAll);
        lc.dynamicSmemBytes = h->smemBytes;
        lc.stream = h->stream;
        cudaLaunchAttribute at{};
        at.id = cudaLaunchAttributeClusterDimension;
        at.val.clusterDim.x = h->b2aCluster;
        at.val.clusterDim.y = at.val.clusterDim.z = 1;
        lc.attrs = &at;
        lc.numAttrs = 1;
        BDS_CUDA(cudaLaunchKernelEx(&lc, trk_b2a_unit_kernel, g));
        count_launch();
        return BDS_OK;
    }
    if (h->fast) {
        BDS_CUDA(cudaMemsetAsync(h->dQueue, 0, sizeof(unsigned long long) * kFwQueueWords * h->qSize, h->stream));
        fw_prepare_kernel<<<1, 1024, 0, h->stream>>>(g, h->nCompute);
    } else trk_prepare_kernel<<<h->nCh, kTrkThreads, sizeof(TrkSmem), h->stream>>>(g);
    count_launch();
    void* args[] = {&g};
    if (h->fast)
        BDS_CUDA(cudaLaunchCooperativeKernel((const void*)kFwGeoms[h->geom].kernel, dim3(h->gridBlocks), dim3(kFwThreads), args,
                                             h->smemBytes, h->stream));
    else
        BDS_CUDA(cudaLaunchCooperativeKernel((const void*)trk_persistent_kernel, dim3(h->gridBlocks), dim3(kTrkThreads),
                                             args, h->smemBytes, h->stream));
    count_launch();
    return BDS_OK;
}

int bds_track_run_async(bds_trk* h, int n_epochs) {
    if (!h || n_epochs <= 0) return set_error(BDS_ERR_ARG, "bds_track_run: bad arguments");
    if (h->map) {
        // file-backed session (the reference's fid): like the reference's one fread per epoch, only the part of the file
        // the requested epochs can touch is brought in - from the earliest channel position to the latest one plus
        // n_epochs + 1 code periods at the slowest plausible code rate - streamed under the tracking kernel
        int rc = bds_track_sync(h);
        if (rc) return rc;
        long long lo = LLONG_MAX, hi = 0;
        const double spc = h->cfg.samplingFreq / (h->cfg.codeFreqBasis / (double)h->cfg.codeLength);
        const long long span = (long long)std::ceil(spc * (1.0 + 1e-4) * (n_epochs + 1)) + 64;
        for (int c = 0; c < h->nCh; ++c) {
            if (h->ch[c].PRN == 0) continue;
            lo = std::min(lo, h->hSt[c].pos);
            hi = std::max(hi, h->hSt[c].pos + span);
        }
        const long long mapSamples = (long long)(h->mapLen >> h->iq);
        if (lo == LLONG_MAX) lo = hi = 0;
        if (lo >= mapSamples) lo = mapSamples - 1;   // past the end of the file: the run reports the short read
        lo = std::max(0LL, lo) & ~4095LL;                       // page aligned in the mapping, 16-byte aligned on the device
        hi = std::min(hi, mapSamples);
        if (hi <= lo) hi = std::min(mapSamples, lo + 1);   // nothing left: the run reports the short read
        return run_streamed_from(h, (const int8_t*)h->map + (lo << h->iq), (size_t)(hi - lo), 0, n_epochs, lo, lo << h->iq);
    }
    int rc = BDS_OK;
    if (h->pending) {  // the output block may be re-allocated below: finish the previous run first
        rc = bds_track_sync(h);
        if (rc) return rc;
    }
    if (h->fast && h->epochsRun + n_epochs > (1 << 19) - 1)
        return set_error(BDS_ERR_UNSUPPORTED, "chip-synchronous kernel: at most 524287 epochs per session");
    rc = ensure_capacity(h, h->epochsRun + n_epochs);
    if (rc) return rc;
    BDS_CUDA(cudaEventRecord(h->ev0, h->stream));
    rc = launch_run(h, n_epochs, h->capacity);
    if (rc) return rc;
    BDS_CUDA(cudaEventRecord(h->ev1, h->stream));
    h->pending = true;
    return BDS_OK;
}

// Streams a HOST record into HBM in chunks on a copy stream while the persistent kernel tracks the
// part that has already arrived: launch i runs every channel as far as chunks 0..i allow (a channel that
// runs out of samples stops exactly like a short read and is resumed from its device-side state by the
// next launch).  Loop state never leaves the device between launches.
int bds_track_run_streamed(bds_trk* h, const int8_t* x, size_t n, size_t chunk_bytes, int n_epochs) {
    return run_streamed_from(h, x, n, chunk_bytes, n_epochs, 0);
}

// Pageable source (a file): bytes [off, off + len) of h->fd -> dst on the device, through a ring of pinned staging
// buffers filled by parallel pread()s (the page cache delivers several times what one thread copies) and drained by
// the copy engine at the pinned rate; cudaMemcpyAsync straight from the pageable mapping runs at a fraction of it.
static int stage_file_range(bds_trk* h, long long off, size_t len, int8_t* dst) {
    constexpr size_t kPiece = (size_t)32 << 20;
    constexpr int kRing = 4;
    if (h->stage.empty()) {
        for (int i = 0; i < kRing; ++i) {
            int8_t* p = nullptr;
            cudaEvent_t e;
            BDS_CUDA(cudaHostAlloc((void**)&p, kPiece, cudaHostAllocDefault));
            h->stage.push_back(p);
            BDS_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            h->stageEv.push_back(e);
        }
    }
    const unsigned nThreads = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
    for (size_t o = 0; o < len; o += kPiece) {
        const size_t piece = std::min(kPiece, len - o);
        const size_t slot = h->stageNext++ % h->stage.size();
        BDS_CUDA(cudaEventSynchronize(h->stageEv[slot]));   // the copy that last used this buffer has drained it
        int8_t* buf = h->stage[slot];
        std::atomic<int> bad{0};
        auto work = [&](unsigned t) {
            const size_t per = ((piece + nThreads - 1) / nThreads + 4095) & ~(size_t)4095;
            size_t a = std::min(piece, (size_t)t * per), b = std::min(piece, a + per);
            while (a < b) {
                const ssize_t r = pread(h->fd, buf + a, b - a, (off_t)(off + (long long)(o + a)));
                if (r <= 0) {
                    bad = 1;
                    return;
                }
                a += (size_t)r;
            }
        };
        std::vector<std::thread> th;
        for (unsigned t = 1; t < nThreads; ++t) th.emplace_back(work, t);
        work(0);
        for (auto& t : th) t.join();
        if (bad) return set_error(BDS_ERR_IO, "read of the IF file failed at byte %lld", off + (long long)o);
        BDS_CUDA(cudaMemcpyAsync(dst + o, buf, piece, cudaMemcpyHostToDevice, h->copyStream));
        BDS_CUDA(cudaEventRecord(h->stageEv[slot], h->copyStream));
    }
    return BDS_OK;
}

// x[n] holds samples [first, first + n) of the record (fileOff >= 0: the same bytes start at this offset of h->fd)
static int run_streamed_from(bds_trk* h, const int8_t* x, size_t n, size_t chunk_bytes, int n_epochs, long long first,
                             long long fileOff) {
    if (!h || !x || n == 0 || n_epochs <= 0) return set_error(BDS_ERR_ARG, "bds_track_run_streamed: bad arguments");
    int rc = BDS_OK;
    if (h->pending) {
        rc = bds_track_sync(h);
        if (rc) return rc;
    }
    if (h->fast && h->epochsRun + n_epochs > (1 << 19) - 1)
        return set_error(BDS_ERR_UNSUPPORTED, "chip-synchronous kernel: at most 524287 epochs per session");
    rc = ensure_capacity(h, h->epochsRun + n_epochs);
    if (rc) return rc;
    if (chunk_bytes == 0) chunk_bytes = (size_t)128 << 20;
    chunk_bytes = (chunk_bytes + 4095) & ~(size_t)4095;
    n <<= h->iq;   // from here on n, the chunk ends and the copies are in BYTES; the window the kernels see is in samples
    if (!h->ownX || h->xCap < n + 64) {
        if (h->ownX && h->dX) cudaFree(h->dX);
        h->dX = nullptr;
        h->ownX = true;
        h->xCap = n + 64;
        BDS_CUDA(cudaMalloc(&h->dX, h->xCap));
    }
    if (!h->copyStream) BDS_CUDA(cudaStreamCreateWithFlags(&h->copyStream, cudaStreamNonBlocking));
    // chunk boundaries: the first chunks are small (16 MiB, doubling) so that tracking starts almost at once
    std::vector<size_t> ends;
    for (size_t o = 0, c = std::min(chunk_bytes, (size_t)16 << 20); o < n; c = std::min(chunk_bytes, c * 2)) {
        o = std::min(n, o + c);
        ends.push_back(o);
    }
    const size_t nChunks = ends.size();
    while (h->chunkEv.size() < nChunks) {
        cudaEvent_t e;
        BDS_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->chunkEv.push_back(e);
    }
    const int limit = h->epochsRun + n_epochs;
    auto copy_chunk = [&](size_t i) -> int {
        const size_t o = i ? ends[i - 1] : 0, len = ends[i] - o;
        if (fileOff >= 0 && h->fd >= 0) {
            int rcf = stage_file_range(h, fileOff + (long long)o, len, h->dX + o);
            if (rcf) return rcf;
        } else {
            BDS_CUDA(cudaMemcpyAsync(h->dX + o, x + o, len, cudaMemcpyHostToDevice, h->copyStream));
        }
        if (i + 1 == nChunks) BDS_CUDA(cudaMemsetAsync(h->dX + n, 0, 64, h->copyStream));
        BDS_CUDA(cudaEventRecord(h->chunkEv[i], h->copyStream));
        return BDS_OK;
    };
    h->winFirst = first;
    BDS_CUDA(cudaEventRecord(h->ev0, h->stream));
    rc = copy_chunk(0);
    if (rc) return rc;
    for (size_t i = 0; i < nChunks; ++i) {
        // the launch over chunks 0..i is queued before chunk i+1 is brought in: staging a file chunk blocks the host
        BDS_CUDA(cudaStreamWaitEvent(h->stream, h->chunkEv[i], 0));
        h->winLen = (long long)(ends[i] >> h->iq);
        rc = launch_run(h, n_epochs, limit);
        if (rc) return rc;
        if (i + 1 < nChunks) {
            rc = copy_chunk(i + 1);
            if (rc) return rc;
        }
    }
    BDS_CUDA(cudaEventRecord(h->ev1, h->stream));
    h->pending = true;
    return BDS_OK;
}

// One launch over a caller-filled DEVICE record of which the first n_avail samples are valid (and device-synchronised):
// every channel runs as far as the data allows, never past epoch index epoch_limit.  Asynchronous; loop state stays on
// the device, so successive calls with a growing n_avail track a record that is still arriving (e.g. over NVLink).
int bds_track_run_window(bds_trk* h, const int8_t* x_dev, size_t n_avail, int epoch_limit) {
    if (!h || !x_dev || epoch_limit <= 0) return set_error(BDS_ERR_ARG, "bds_track_run_window: bad arguments");
    if (((uintptr_t)x_dev & 15) != 0) return set_error(BDS_ERR_ARG, "device IF buffer must be 16-byte aligned");
    int rc = BDS_OK;
    if (epoch_limit > h->capacity) {   // the output block is re-allocated: finish what is in flight first
        if (h->pending) {
            rc = bds_track_sync(h);
            if (rc) return rc;
        }
        rc = ensure_capacity(h, epoch_limit);
        if (rc) return rc;
    }
    if (h->fast && epoch_limit > (1 << 19) - 1)
        return set_error(BDS_ERR_UNSUPPORTED, "chip-synchronous kernel: at most 524287 epochs per session");
    if (h->ownX && h->dX) {   // a record owned by the session (earlier host / streamed run) is replaced
        if (h->copyStream) BDS_CUDA(cudaStreamSynchronize(h->copyStream));
        BDS_CUDA(cudaStreamSynchronize(h->stream));
        cudaFree(h->dX);
    }
    h->dX = const_cast<int8_t*>(x_dev);
    h->ownX = false;
    h->xCap = n_avail << h->iq;
    h->winFirst = 0;
    h->winLen = (long long)n_avail;   // nothing is read past x_dev[n_avail-1] (TrkDev::winStage)
    if (!h->pending) BDS_CUDA(cudaEventRecord(h->ev0, h->stream));
    rc = launch_run(h, epoch_limit, epoch_limit);
    if (rc) return rc;
    BDS_CUDA(cudaEventRecord(h->ev1, h->stream));
    h->pending = true;
    return BDS_OK;
}

int bds_track_sync(bds_trk* h) {
    if (!h) return set_error(BDS_ERR_ARG, "null handle");
    BDS_CUDA(cudaStreamSynchronize(h->stream));
    if (h->pending) {
        cudaEventElapsedTime(&h->lastMs, h->ev0, h->ev1);
        h->pending = false;
    }
    BDS_CUDA(cudaMemcpy(h->hSt.data(), h->dSt, sizeof(ChanState) * h->nCh, cudaMemcpyDeviceToHost));
    int me = 0;
    for (auto& s : h->hSt) me = std::max(me, s.epoch);
    h->epochsRun = me;
    return BDS_OK;
}

int bds_track_fetch(bds_trk* h, const bds_trk_out* o, int stride) {
    if (!h || !o) return set_error(BDS_ERR_ARG, "null argument");
    int rc = bds_track_sync(h);
    if (rc) return rc;
    if (stride < h->epochsRun) return set_error(BDS_ERR_ARG, "out_stride %d < epochs run %d", stride, h->epochsRun);
    // entries past the last completed epoch carry the reference's preallocation values
    const int nE = std::min(stride, h->capacity);
    if (nE == 0) return BDS_OK;
    double* planes[kNFieldsLoop] = {o->absoluteSample, o->codeFreq, o->carrFreq, o->I_P, o->I_E, o->I_L, o->Q_E,
                                    o->Q_P, o->Q_L, o->Pilot_I_P, o->Pilot_I_E, o->Pilot_I_L, o->Pilot_Q_E,
                                    o->Pilot_Q_P, o->Pilot_Q_L, o->dllDiscr, o->dllDiscrFilt, o->pllDiscr,
                                    o->pllDiscrFilt, o->remCodePhase, o->remCarrPhase};
    // device block [c][field][capacity] -> caller plane [c][stride]: one strided copy per requested plane, straight
    // into the caller's memory (pinned destinations run at full PCIe rate)
    const size_t srcPitch = sizeof(double) * (size_t)kNFields * h->capacity;
    for (int f = 0; f < kNFieldsLoop; ++f)
        if (planes[f])
            BDS_CUDA(cudaMemcpy2DAsync(planes[f], sizeof(double) * (size_t)stride, h->dOut + (size_t)f * h->capacity,
                                       srcPitch, sizeof(double) * (size_t)nE, (size_t)h->nCh, cudaMemcpyDeviceToHost,
                                       h->stream));
    std::vector<double> raw;
    if (o->raw) {
        raw.resize((size_t)h->nCh * kNSum * nE);
        for (int c = 0; c < h->nCh; ++c)
            for (int k = 0; k < kNSum; ++k)
                BDS_CUDA(cudaMemcpyAsync(raw.data() + ((size_t)c * kNSum + k) * nE,
                                         h->dOut + ((size_t)c * kNFields + F_RAW0 + k) * h->capacity,
                                         sizeof(double) * (size_t)nE, cudaMemcpyDeviceToHost, h->stream));
    }
    const int ci = std::max(1, h->cfg.CNoInterval);
    const int nC = stride / ci;
    std::vector<double> cn;
    if (nC > 0 && (o->DataCNo || o->DataPLD || o->PilotCNo || o->PilotPLD || o->TotalCNo)) {
        cn.resize((size_t)h->nCh * kNCno * h->cnoCap);
        BDS_CUDA(cudaMemcpyAsync(cn.data(), h->dCno, cn.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    }
    BDS_CUDA(cudaStreamSynchronize(h->stream));
    if (o->raw)
        for (int c = 0; c < h->nCh; ++c)
            for (int e = 0; e < nE; ++e)
                for (int k = 0; k < kNSum; ++k)
                    o->raw[((size_t)c * stride + e) * kNSum + k] = raw[((size_t)c * kNSum + k) * nE + e];
    if (o->epochsDone)
        for (int c = 0; c < h->nCh; ++c) o->epochsDone[c] = h->hSt[c].epoch;
    if (o->lockLostEpoch)
        for (int c = 0; c < h->nCh; ++c) o->lockLostEpoch[c] = (int32_t)h->hSt[c].lockLost;
    if (!cn.empty()) {
        double* cp[kNCno] = {o->DataCNo, o->DataPLD, o->PilotCNo, o->PilotPLD, o->TotalCNo};
        const int n = std::min(nC, h->cnoCap);
        for (int c = 0; c < h->nCh; ++c)
            for (int f = 0; f < kNCno; ++f)
                if (cp[f]) {
                    std::memset(cp[f] + (size_t)c * nC, 0, sizeof(double) * nC);
                    std::memcpy(cp[f] + (size_t)c * nC, cn.data() + ((size_t)c * kNCno + f) * h->cnoCap, sizeof(double) * n);
                }
    }
    return BDS_OK;
}

int bds_track_run(bds_trk* h, int n_epochs, const bds_trk_out* out, int out_stride) {
    int rc = bds_track_run_async(h, n_epochs);
    if (rc) return rc;
    return bds_track_fetch(h, out, out_stride);
}

int bds_track_device_block(bds_trk* h, void** dev_ptr, size_t* bytes, int* n_fields, int* capacity) {
    if (!h) return set_error(BDS_ERR_ARG, "null handle");
    if (dev_ptr) *dev_ptr = h->dOut;
    if (bytes) *bytes = sizeof(double) * (size_t)h->nCh * kNFields * h->capacity;
    if (n_fields) *n_fields = kNFields;
    if (capacity) *capacity = h->capacity;
    return BDS_OK;
}

int bds_track_stats(bds_trk* h, long long* channel_samples, int* epochs_run, float* last_kernel_ms) {
    if (!h) return set_error(BDS_ERR_ARG, "null handle");
    int rc = bds_track_sync(h);
    if (rc) return rc;
    long long tot = 0;
    int me = 0;
    for (auto& s : h->hSt) {
        tot += s.samples;
        me = std::max(me, s.epoch);
    }
    if (channel_samples) *channel_samples = tot;
    if (epochs_run) *epochs_run = me;
    if (last_kernel_ms) *last_kernel_ms = h->lastMs;
    return BDS_OK;
}

int bds_track_counters(bds_trk* h, long long* out4) {
    if (!out4) return set_error(BDS_ERR_ARG, "null argument");
    if (!h) {  // counters of the last bds_track_correlate_open_loop call
        for (int i = 0; i < 4; ++i) out4[i] = (long long)g_open_loop_counters[i];
        return BDS_OK;
    }
    BDS_CUDA(cudaStreamSynchronize(h->stream));
    unsigned long long v[24];
    BDS_CUDA(cudaMemcpy(v, h->dCounters, 192, cudaMemcpyDeviceToHost));
    for (int i = 0; i < 4; ++i) out4[i] = (long long)v[i];
    if ((h->cfg.debug & BDS_DBG_TIMING) && h->b2aUnit && v[8])   // developer build (-DBDS_FW_DEV): cycles per epoch of thread 0
        fprintf(stderr, "[bds timing] b2a per epoch (cycles): tile wait %.0f, correlate %.0f, loop closure %.0f (exchange %.0f, sums + discriminators %.0f, "
                        "filters %.0f, next NCO || table + barrier %.0f) [%.0f]\n",
                (double)v[4] / v[8], (double)v[5] / v[8], (double)(v[6] + v[9] + v[10] + v[11]) / v[8], (double)v[9] / v[8], (double)v[10] / v[8],
                (double)v[11] / v[8], (double)v[6] / v[8], (double)v[7] / v[8]);
    if ((h->cfg.debug & BDS_DBG_TIMING) && !h->b2aUnit)  // developer breakdown (SM cycles summed over CTAs)
        fprintf(stderr, "[bds timing] producer: queue %llu empty %llu total %llu | compute(w2): full-wait %llu res-wait %llu | closer: closure %llu epilogue %llu closures %llu\n",
                v[4], v[5], v[6], v[7], v[8], v[9], v[10], v[11]);
    if (h->cfg.debug & BDS_DBG_TIMING)
        fprintf(stderr, "[bds timing] producer detail: ticket atomic %llu, proxy fence %llu, pass issue %llu\n", v[18], v[19], v[20]);
    if ((h->cfg.debug & BDS_DBG_TIMING) && v[11])
        fprintf(stderr, "[bds timing] per closure: publish->slices done %.2f us, closure %.2f us\n",
                (double)v[12] / (double)v[11] * 1e-3, (double)v[13] / (double)v[11] * 1e-3);
    if ((h->cfg.debug & BDS_DBG_TIMING) && v[11])
        fprintf(stderr, "[bds timing] closure cycles: reduce %.0f, close_epoch %.0f, build_tab %.0f, fence+publish %.0f\n",
                (double)v[14] / v[11], (double)v[15] / v[11], (double)v[16] / v[11], (double)v[17] / v[11]);
    return BDS_OK;
}

int bds_track_dump_trace(bds_trk* h, const char* path) {
    if (!h || !h->dTrace || !path) return set_error(BDS_ERR_ARG, "no trace");
    BDS_CUDA(cudaStreamSynchronize(h->stream));
    std::vector<unsigned long long> t((size_t)h->traceCap * 8);
    BDS_CUDA(cudaMemcpy(t.data(), h->dTrace, t.size() * 8, cudaMemcpyDeviceToHost));
    FILE* f = fopen(path, "wb");
    if (!f) return set_error(BDS_ERR_IO, "cannot write %s", path);
    fwrite(t.data(), 8, t.size(), f);
    fclose(f);
    return BDS_OK;
}

int bds_track_reset(bds_trk* h) {
    if (!h) return set_error(BDS_ERR_ARG, "null handle");
    BDS_CUDA(cudaStreamSynchronize(h->stream));
    int rc = init_state(h);
    if (rc) return rc;
    if (h->capacity > 0) {
        rc = init_out_block(h, h->dOut, h->capacity, 0);
        if (rc) return rc;
        BDS_CUDA(cudaMemsetAsync(h->dCno, 0, sizeof(double) * (size_t)h->nCh * kNCno * h->cnoCap, h->stream));
    }
    return BDS_OK;
}

void bds_track_close(bds_trk* h) {
    if (!h) return;
    if (h->copyStream) cudaStreamSynchronize(h->copyStream);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->map) munmap(h->map, h->mapLen);
    if (h->fd >= 0) ::close(h->fd);
    for (auto p : h->stage) cudaFreeHost(p);
    for (auto e : h->stageEv) cudaEventDestroy(e);
    if (h->ownX && h->dX) cudaFree(h->dX);
    cudaFree(h->dBits);
    cudaFree(h->dCC);
    cudaFree(h->dSt);
    cudaFree(h->dParams);
    cudaFree(h->dReady);
    cudaFree(h->dStop);
    cudaFree(h->dCount);
    cudaFree(h->dPartial);
    cudaFree(h->dAcc);
    cudaFree(h->dOut);
    cudaFree(h->dCno);
    cudaFree(h->dAct);
    cudaFree(h->dCounters);
    cudaFree(h->dQueue);
    cudaFree(h->dQctl);
    cudaFree(h->dTrace);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    for (auto e : h->chunkEv) cudaEventDestroy(e);
    if (h->copyStream) cudaStreamDestroy(h->copyStream);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int bds_track_correlate_open_loop(int mode, const bds_trk_cfg* cfg, const int8_t* x, size_t n, int x_loc,
                                  const int32_t* prn, int n_ch, int n_epochs, const double* nco, double* sums) {
    if (!cfg || !x || !prn || !nco || !sums || n_ch <= 0 || n_epochs <= 0)
        return set_error(BDS_ERR_ARG, "bds_track_correlate_open_loop: bad arguments");
    int hp, h6;
    if (!mode_flags(mode, cfg->pilotTRKflag, hp, h6)) return set_error(BDS_ERR_ARG, "unknown tracking mode %d", mode);
    int rc = require_device();
    if (rc) return rc;
    bool fast = false;
    int geom = 0;
    rc = choose_fast(mode, cfg, fast, geom);
    if (rc) return rc;
    const int nce = n_ch * n_epochs;
    std::vector<EpochParams> hp_(nce);
    for (int i = 0; i < nce; ++i) {
        const double* q = nco + (size_t)i * 6;
        EpochParams& p = hp_[i];
        p.pos = (long long)q[0];
        p.blksize = (int)q[1];
        p.pad = 0;
        p.rem = q[2];
        p.step = q[3];
        p.carrFreq = q[4];
        p.remCarr = q[5];
        if (p.pos < 0 || p.blksize <= 0 || (size_t)(p.pos + p.blksize) > n)
            return set_error(BDS_ERR_ARG, "open loop: block %d outside the record", i);
    }
    const int iq = cfg->fileType == 2;
    const size_t nb = n << iq;   // bytes of the record
    std::vector<uint32_t> bits;
    rc = build_code_bits(mode, prn, n_ch, bits);
    if (rc) return rc;
    int8_t* dX = nullptr;
    uint32_t* dBits = nullptr;
    EpochParams* dP = nullptr;
    double *dPart = nullptr, *dSums = nullptr;
    unsigned long long* dCnt = nullptr;
    const int S = fast ? (10230 + kFwChips * 3 - 1) / (kFwChips * 3) : 32;
    auto cleanup = [&]() {
        if (x_loc == BDS_LOC_HOST) cudaFree(dX);
        cudaFree(dBits);
        cudaFree(dP);
        cudaFree(dPart);
        cudaFree(dSums);
        cudaFree(dCnt);
    };
#define TRYC(x_)                                                       \
    if ((x_) != cudaSuccess) {                                         \
        cleanup();                                                     \
        return set_error(BDS_ERR_CUDA, "open loop: %s failed", #x_);   \
    }
    if (x_loc == BDS_LOC_HOST) {
        TRYC(cudaMalloc(&dX, nb + 64));
        TRYC(cudaMemcpy(dX, x, nb, cudaMemcpyHostToDevice));
        TRYC(cudaMemset(dX + nb, 0, 64));
    } else {
        dX = const_cast<int8_t*>(x);
    }
    TRYC(cudaMalloc(&dBits, bits.size() * 4));
    TRYC(cudaMemcpy(dBits, bits.data(), bits.size() * 4, cudaMemcpyHostToDevice));
    TRYC(cudaMalloc(&dP, sizeof(EpochParams) * nce));
    TRYC(cudaMemcpy(dP, hp_.data(), sizeof(EpochParams) * nce, cudaMemcpyHostToDevice));
    TRYC(cudaMalloc(&dPart, sizeof(double) * (size_t)nce * S * kNSum));
    TRYC(cudaMalloc(&dSums, sizeof(double) * (size_t)nce * kNSum));
    TrkDev g;
    std::memset(&g, 0, sizeof(g));
    g.x = dX;
    g.winFirst = 0;
    // staged tiles may extend 16 bytes past winLen only where the buffer has slack: a host record was copied with 64
    // bytes of it, a caller-owned device buffer has none
    g.winLen = (long long)n;
    g.winStage = x_loc == BDS_LOC_HOST ? (((long long)n + 16) & ~15LL) : ((long long)n & ~15LL);
    g.mode = mode;
    g.hasPilot = hp;
    g.hasP61 = h6;
    g.nCh = n_ch;
    g.fs = cfg->samplingFreq;
    g.L = (double)cfg->codeLength;
    g.d = cfg->dllCorrelatorSpacing;
    g.codeBits = dBits;
    g.S = S;
    g.pad = cfg->reserved & 1;
    g.iq = iq;
    TRYC(cudaMalloc(&dCnt, 128));
    TRYC(cudaMemset(dCnt, 0, 128));
    g.counters = dCnt;
    size_t smem = smem_bytes(fast, geom);
    const bool b2aUnit = !fast && b2a_unit_enabled(mode, cfg);
    if (b2aUnit) {
        smem = sizeof(B2aSmem);
        TRYC(cudaFuncSetAttribute(trk_b2a_unit_open_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        trk_b2a_unit_open_kernel<<<nce, kB2aThreads, smem>>>(g, dP, n_epochs, dSums);
    } else if (fast) {
        TRYC(cudaFuncSetAttribute(kFwGeoms[geom].kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        g.partial = dPart;
        g.olParams = dP;
        g.olEpochs = n_epochs;
        g.olCount = nce;
        g.nCompute = g_num_sms;
        g.stages = kFwStages;
        g.tune = 0;
        g.ahead = -1;
        cudaEvent_t e0, e1;
        TRYC(cudaEventCreate(&e0));
        TRYC(cudaEventCreate(&e1));
        TRYC(cudaEventRecord(e0, 0));
        kFwGeoms[geom].kernel<<<g_num_sms, kFwThreads, smem>>>(g);
        TRYC(cudaEventRecord(e1, 0));
        TRYC(cudaEventSynchronize(e1));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        g_open_loop_us = (unsigned long long)(ms * 1000.f);
    } else {
        trk_open_loop_kernel<<<dim3(S, nce), kTrkThreads, smem>>>(g, dP, n_epochs, dPart);
    }
    count_launch();
    if (!b2aUnit) {
        trk_open_loop_reduce_kernel<<<(nce * kNSum + 127) / 128, 128>>>(dPart, S, nce, dSums);
        count_launch();
    }
    TRYC(cudaGetLastError());
    TRYC(cudaMemcpy(sums, dSums, sizeof(double) * (size_t)nce * kNSum, cudaMemcpyDeviceToHost));
    TRYC(cudaMemcpy(g_open_loop_counters, dCnt, 32, cudaMemcpyDeviceToHost));
    g_open_loop_counters[3] = g_open_loop_us;
#undef TRYC
    cleanup();
    return BDS_OK;
}

}  // extern "C"

// FFT parallel code-phase acquisition (sm_100a) — replaces
//   BDS-3_B1C/acquisition.m:129-338 and BDS-3_B2a/acquisition.m:130-365.
//
// The reference correlates over the exact circular length N (1 987 500 for B1C at
// 99.375 MHz, not a power of two) with MATLAB's fft/ifft.  Here the same N-circular
// correlation is obtained exactly from a radix-2 FFT of length P = 2^p >= N+M-1: the
// local code is non-zero only on its first M samples, so the carrier-mixed signal is
// extended periodically by M-1 samples and zero padded, and lags 0..N-1 of the P-point
// linear correlation equal the N-circular correlation (SURVEY §7 hard part 4).
//
// FFT: four-step P = P1 x P2, both passes in shared memory, hand-rolled radix-2
// decimation-in-frequency forward (natural in, bit-reversed out) and decimation-in-time
// inverse (bit-reversed in, natural out), so no reordering pass is ever needed: signal and
// code spectra live in the same scrambled layout and are only multiplied pointwise.
//   forward  : column pass (int8 load + carrier mix fused, twiddle fused) -> row pass
//   inverse  : row pass (spectrum x conj(code spectrum) fused, twiddle fused)
//              -> column pass (|.|, data/pilot combine, max/arg-max fused; the
//              `results` matrix of acquisition.m:154 is never materialised)
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "bds_codes.h"
#include "bds_common.cuh"

namespace bds {

constexpr int kTwN = 4096;           // butterfly twiddle table size (covers sub-FFTs <= 4096)
constexpr int kColTile = 8;          // columns per CTA in the column passes
constexpr int kAcqThreads = 512;

struct AcqPlan {
    int log2P, log2P1, log2P2;   // P = P1*P2 ; P1 = column (strided) length, P2 = row length
    int P, P1, P2;
    int N, M, Next;              // circular length, code length, N+M-1
    const float2* tw;            // exp(-2 pi i k / kTwN)
    const float2* twHi;          // exp(-2 pi i (k*2048) / P), k < P/2048
    const float2* twLo;          // exp(-2 pi i k / P), k < 2048
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {  // a * conj(b)
    return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
// W_P^m, 0 <= m < P
__device__ __forceinline__ float2 twiddleP(const AcqPlan& pl, unsigned m) {
    float2 hi = __ldg(pl.twHi + (m >> 11)), lo = __ldg(pl.twLo + (m & 2047));
    return cmul(hi, lo);
}

// In-place radix-2 FFT over `len` = 2^lg elements for `batch` independent sequences held in
// shared memory: element (i, b) at buf[i*si + b*sb].  kInv=false: DIF forward (natural ->
// bit reversed).  kInv=true: DIT inverse (bit reversed -> natural), unnormalised.
template <bool kInv>
__device__ void smem_fft(float2* buf, int lg, int batch, int si, int sb, const float2* tw) {
    const int len = 1 << lg, half = len >> 1;
    const int total = half * batch;
    for (int s = 0; s < lg; ++s) {
        const int lh = kInv ? s : (lg - 1 - s);  // log2 of the half span
        const int h = 1 << lh;
        const int twStride = kTwN >> (lh + 1);
        for (int u = threadIdx.x; u < total; u += blockDim.x) {
            int b = u % batch, v = u / batch;
            int j = v & (h - 1);
            int i0 = ((v >> lh) << (lh + 1)) + j;
            float2* pa = buf + i0 * si + b * sb;
            float2* pb = pa + h * si;
            float2 a = *pa, c = *pb;
            float2 w = __ldg(tw + j * twStride);
            if (kInv) {
                float2 t = cmulc(c, w);
                *pa = make_float2(a.x + t.x, a.y + t.y);
                *pb = make_float2(a.x - t.x, a.y - t.y);
            } else {
                float2 d = make_float2(a.x - c.x, a.y - c.y);
                *pa = make_float2(a.x + c.x, a.y + c.y);
                *pb = cmul(d, w);
            }
        }
        __syncthreads();
    }
}

// ---- forward column pass -------------------------------------------------------------
// kind 0: signal  y[n] = x[n mod N] * exp(+i 2 pi f n'/fs), n' = n mod N, n < N+M-1, else 0
//                 (acquisition.m:194-205; periodic extension see file header)
// kind 1: code    y[n] = table[n] for n < M else 0   (acquisition.m:176-180)
// grid = (P2/kColTile, batch); batch index selects the Doppler bin / code table.
__global__ void __launch_bounds__(kAcqThreads) acq_fwd_col_kernel(AcqPlan pl, int kind, const int8_t* src,
                                                                  size_t srcStride, const unsigned long long* dphi,
                                                                  float2* spec) {
    extern __shared__ __align__(16) unsigned char smraw[];
    float2* buf = reinterpret_cast<float2*>(smraw);
    const int col0 = blockIdx.x * kColTile;
    const int bi = blockIdx.y;
    const int8_t* x = src + (size_t)bi * srcStride;
    const unsigned long long dp = kind == 0 ? dphi[bi] : 0ull;
    const int total = pl.P1 * kColTile;
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
        int c = e & (kColTile - 1), r = e >> 3;
        unsigned n = (unsigned)r * pl.P2 + col0 + c;
        float2 v = make_float2(0.f, 0.f);
        if (kind == 0) {
            if (n < (unsigned)pl.Next) {
                unsigned m = n >= (unsigned)pl.N ? n - pl.N : n;
                float xs = (float)x[m];
                unsigned long long ph = (unsigned long long)m * dp;
                float sn, cs;
                sincospif((float)(int)(ph >> 32) * 4.656612873077392578125e-10f, &sn, &cs);
                v = make_float2(xs * cs, xs * sn);
            }
        } else if (n < (unsigned)pl.M) {
            v.x = (float)x[n];
        }
        buf[e] = v;
    }
    __syncthreads();
    smem_fft<false>(buf, pl.log2P1, kColTile, kColTile, 1, pl.tw);
    float2* out = spec + (size_t)bi * pl.P;
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
        int c = e & (kColTile - 1), r = e >> 3;
        unsigned k1 = __brev((unsigned)r) >> (32 - pl.log2P1);
        unsigned n2 = col0 + c;
        float2 w = twiddleP(pl, k1 * n2);
        out[(size_t)r * pl.P2 + n2] = cmul(buf[e], w);
    }
}

// ---- forward row pass (in place) -----------------------------------------------------
// conjScale != 0: store conj(X) * conjScale (code spectra: acquisition.m:180 with the 1/P of
// the inverse transform folded in).
__global__ void __launch_bounds__(kAcqThreads) acq_fwd_row_kernel(AcqPlan pl, float2* spec, float conjScale) {
    extern __shared__ __align__(16) unsigned char smraw[];
    float2* buf = reinterpret_cast<float2*>(smraw);
    float2* row = spec + (size_t)blockIdx.y * pl.P + (size_t)blockIdx.x * pl.P2;
    for (int i = threadIdx.x; i < pl.P2; i += blockDim.x) buf[i] = row[i];
    __syncthreads();
    smem_fft<false>(buf, pl.log2P2, 1, 1, 0, pl.tw);
    for (int i = threadIdx.x; i < pl.P2; i += blockDim.x) {
        float2 v = buf[i];
        if (conjScale != 0.f) v = make_float2(v.x * conjScale, -v.y * conjScale);
        row[i] = v;
    }
}

// ---- inverse row pass ------------------------------------------------------------------
// work[bin][dp] row r = IDFT_row( sig[bin] row r .* code[dp] row r ) .* W_P^{-n2 k1}
// grid = (P1, nbins, ncodes)
__global__ void __launch_bounds__(kAcqThreads) acq_inv_row_kernel(AcqPlan pl, const float2* sig, const float2* code,
                                                                  float2* work, int ncodes) {
    extern __shared__ __align__(16) unsigned char smraw[];
    float2* buf = reinterpret_cast<float2*>(smraw);
    const int r = blockIdx.x, bin = blockIdx.y, dp = blockIdx.z;
    const float2* srow = sig + (size_t)bin * pl.P + (size_t)r * pl.P2;
    const float2* crow = code + (size_t)dp * pl.P + (size_t)r * pl.P2;
    for (int i = threadIdx.x; i < pl.P2; i += blockDim.x) buf[i] = cmul(srow[i], __ldg(crow + i));
    __syncthreads();
    smem_fft<true>(buf, pl.log2P2, 1, 1, 0, pl.tw);
    const unsigned k1 = __brev((unsigned)r) >> (32 - pl.log2P1);
    float2* orow = work + ((size_t)bin * ncodes + dp) * pl.P + (size_t)r * pl.P2;
    for (int i = threadIdx.x; i < pl.P2; i += blockDim.x) {
        float2 w = twiddleP(pl, k1 * (unsigned)i);
        orow[i] = cmulc(buf[i], w);
    }
}

// ---- inverse column pass + magnitude + combine + max -------------------------------------
struct AcqPeak {
    float val;
    int lag;  // 0-based
};
// combine: 0 data only; 1 B1C (|d| sqrt11 + |p| sqrt29)/sqrt40 (acquisition.m:215-220);
//          2 B2a |d|+|p| (B2a acquisition.m:205-209)
// Lags are restricted to [0,N) and, when ex0 <= ex1 (second-peak search, B2a
// acquisition.m:224-249, 0-based inclusive bounds), to [lo0,hi0] U [lo1,hi1].
__global__ void __launch_bounds__(kAcqThreads) acq_inv_col_kernel(AcqPlan pl, const float2* work, int ncodes,
                                                                  int combine, int lo0, int hi0, int lo1, int hi1,
                                                                  int useRanges, AcqPeak* peaks /*[bin][gridDim.x]*/) {
    extern __shared__ __align__(16) unsigned char smraw[];
    float2* buf = reinterpret_cast<float2*>(smraw);
    const int total = pl.P1 * kColTile;
    float* mag = reinterpret_cast<float*>(buf + total);
    const int col0 = blockIdx.x * kColTile, bin = blockIdx.y;
    for (int dp = 0; dp < ncodes; ++dp) {
        const float2* w = work + ((size_t)bin * ncodes + dp) * pl.P;
        for (int e = threadIdx.x; e < total; e += blockDim.x) {
            int c = e & (kColTile - 1), r = e >> 3;
            buf[e] = w[(size_t)r * pl.P2 + col0 + c];
        }
        __syncthreads();
        smem_fft<true>(buf, pl.log2P1, kColTile, kColTile, 1, pl.tw);
        for (int e = threadIdx.x; e < total; e += blockDim.x) {
            float2 v = buf[e];
            float m = sqrtf(v.x * v.x + v.y * v.y);
            if (dp == 0)
                mag[e] = m;
            else if (combine == 1)
                mag[e] = (mag[e] * 3.3166247903554f + m * 5.3851648071345f) / 6.3245553203368f;
            else
                mag[e] = mag[e] + m;
        }
        __syncthreads();
    }
    float best = -1.f;
    int bl = 0x7fffffff;
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
        int c = e & (kColTile - 1), r = e >> 3;
        int lag = r * pl.P2 + col0 + c;
        bool ok = lag < pl.N;
        if (useRanges) ok = ok && ((lag >= lo0 && lag <= hi0) || (lag >= lo1 && lag <= hi1));
        float m = mag[e];
        if (ok && (m > best || (m == best && lag < bl))) {
            best = m;
            bl = lag;
        }
    }
    // block reduce (max value, then min lag)
    __shared__ float sv[kAcqThreads / 32];
    __shared__ int sl[kAcqThreads / 32];
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, best, o);
        int ol = __shfl_xor_sync(0xffffffffu, bl, o);
        if (ov > best || (ov == best && ol < bl)) {
            best = ov;
            bl = ol;
        }
    }
    if ((threadIdx.x & 31) == 0) {
        sv[threadIdx.x >> 5] = best;
        sl[threadIdx.x >> 5] = bl;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
            if (sv[w] > best || (sv[w] == best && sl[w] < bl)) {
                best = sv[w];
                bl = sl[w];
            }
        peaks[(size_t)bin * gridDim.x + blockIdx.x] = AcqPeak{best, bl};
    }
}

// per-bin reduction of the column-group peaks
__global__ void acq_peak_reduce_kernel(const AcqPeak* peaks, int groups, AcqPeak* binPeak) {
    int bin = blockIdx.x;
    float best = -1.f;
    int bl = 0x7fffffff;
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
        AcqPeak p = peaks[(size_t)bin * groups + g];
        if (p.val > best || (p.val == best && p.lag < bl)) {
            best = p.val;
            bl = p.lag;
        }
    }
    __shared__ float sv[256];
    __shared__ int sl[256];
    sv[threadIdx.x] = best;
    sl[threadIdx.x] = bl;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int t = 1; t < (int)blockDim.x; ++t)
            if (sv[t] > best || (sv[t] == best && sl[t] < bl)) {
                best = sv[t];
                bl = sl[t];
            }
        binPeak[bin] = AcqPeak{best, bl};
    }
}

// ---- integer power sums for sigPower (acquisition.m:150) ------------------------------------
__global__ void acq_power_kernel(const int8_t* x, int n, long long* sums /*[2]*/) {
    long long s1 = 0, s2 = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int v = x[i];
        s1 += v;
        s2 += v * v;
    }
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd((unsigned long long*)&sums[0], (unsigned long long)s1);
        atomicAdd((unsigned long long*)&sums[1], (unsigned long long)s2);
    }
}

// ---- fine search ---------------------------------------------------------------------------
// B1C (acquisition.m:253-300): for fine bin j, component dp:
//   A = sum_n x[n] T[n] e^{i phi_j n},  B = sum_n T[n] e^{i phi_j n}   (DC removal applied on host: A - mean*B)
// out[(j*ncodes+dp)*4 + {0..3}] = {Re A, Im A, Re B, Im B}; grid = (slices, nfine, ncodes)
__global__ void acq_fine_b1c_kernel(const int8_t* x, const int8_t* tables, int spc, const unsigned long long* dphi,
                                    int ncodes, double* out) {
    const int j = blockIdx.y, dp = blockIdx.z;
    const int8_t* T = tables + (size_t)dp * spc;
    const unsigned long long d = dphi[j];
    float ar = 0, ai = 0, br = 0, bi = 0;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < spc; n += gridDim.x * blockDim.x) {
        unsigned long long ph = (unsigned long long)n * d;
        float sn, cs;
        sincospif((float)(int)(ph >> 32) * 4.656612873077392578125e-10f, &sn, &cs);
        float t = (float)T[n], xs = (float)x[n] * t;
        ar += xs * cs;
        ai += xs * sn;
        br += t * cs;
        bi += t * sn;
    }
    double v[4] = {warp_sum((double)ar), warp_sum((double)ai), warp_sum((double)br), warp_sum((double)bi)};
    if ((threadIdx.x & 31) == 0) {
        double* o = out + ((size_t)j * ncodes + dp) * 4;
#pragma unroll
        for (int k = 0; k < 4; ++k) atomicAdd(o + k, v[k]);
    }
}

// B2a (acquisition.m:279-320): K = fineNoncoh*spc samples; code index floor((ts*n)/tc), n = 1..K,
// rem(.,10230); per-ms coherent sums.  out[((j*2+dp)*nseg + seg)*2 + {re,im}]; grid = (slices, nfine, nseg)
__global__ void acq_fine_b2a_kernel(const int8_t* x, const uint32_t* bits /*[2][320]*/, int spc, double ts,
                                    double tc, const unsigned long long* dphi, int nseg, double* out) {
    const int j = blockIdx.y, seg = blockIdx.z;
    const unsigned long long d = dphi[j];
    float dr = 0, di = 0, pr = 0, pi = 0;
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < spc; m += gridDim.x * blockDim.x) {
        long long n0 = (long long)seg * spc + m;  // 0-based sample index into sigFineACQ
        double ci = floor(__ddiv_rn(__dmul_rn(ts, (double)(n0 + 1)), tc));
        int chip = (int)(ci - floor(ci / 10230.0) * 10230.0);
        float cd = (bits[chip >> 5] >> (chip & 31)) & 1 ? -1.f : 1.f;
        float cp = (bits[kPackedWords + (chip >> 5)] >> (chip & 31)) & 1 ? -1.f : 1.f;
        unsigned long long ph = (unsigned long long)n0 * d;
        float sn, cs;
        sincospif((float)(int)(ph >> 32) * 4.656612873077392578125e-10f, &sn, &cs);
        float xs = (float)x[n0];
        dr += cd * xs * cs;
        di += cd * xs * sn;
        pr += cp * xs * cs;
        pi += cp * xs * sn;
    }
    double v[4] = {warp_sum((double)dr), warp_sum((double)di), warp_sum((double)pr), warp_sum((double)pi)};
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(out + (((size_t)j * 2 + 0) * nseg + seg) * 2 + 0, v[0]);
        atomicAdd(out + (((size_t)j * 2 + 0) * nseg + seg) * 2 + 1, v[1]);
        atomicAdd(out + (((size_t)j * 2 + 1) * nseg + seg) * 2 + 0, v[2]);
        atomicAdd(out + (((size_t)j * 2 + 1) * nseg + seg) * 2 + 1, v[3]);
    }
}

}  // namespace bds

// ===========================================================================================
using namespace bds;

namespace {

struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { cudaFree(p); }
    template <typename T>
    T* as() { return reinterpret_cast<T*>(p); }
    cudaError_t alloc(size_t bytes) {
        cudaFree(p);
        p = nullptr;
        return cudaMalloc(&p, bytes);
    }
};

unsigned long long freq_to_dphi(double f, double fs) {
    double r = f / fs;
    r -= std::floor(r);
    return (unsigned long long)(r * 18446744073709551616.0);
}

long mround(double x) { return std::lround(x); }

int sampled_table(int component, int prn, double fs, double fcb, int L, std::vector<int8_t>& t) {
    long spc = mround(fs / (fcb / L));
    t.resize(spc);
    return bds_make_code_table(component, prn, fs, fcb, L, t.data(), (int)spc);
}

}  // namespace

extern "C" int bds_acquire(int signal, const int8_t* x, size_t n, int x_loc, const bds_acq_cfg* cfg,
                           const int32_t* prn, int n_prn, int prn_lo, int prn_hi, double* carrFreq,
                           double* codePhase, double* peakMetric, int max_prn, double* dbg) {
    if (!x || !cfg || !prn || !carrFreq || !codePhase || !peakMetric || n_prn <= 0)
        return set_error(BDS_ERR_ARG, "bds_acquire: null/empty argument");
    if (signal != BDS_SIG_B1C && signal != BDS_SIG_B2A) return set_error(BDS_ERR_ARG, "unknown signal %d", signal);
    if (cfg->codeLength != kCodeLen) return set_error(BDS_ERR_UNSUPPORTED, "codeLength must be 10230");
    for (int i = 0; i < n_prn; ++i)
        if (prn[i] < 1 || prn[i] > 63 || prn[i] > max_prn)
            return set_error(BDS_ERR_ARG, "PRN %d out of range (max_prn %d)", prn[i], max_prn);
    int rc = require_device();
    if (rc) return rc;
    std::memset(carrFreq, 0, sizeof(double) * max_prn);
    std::memset(codePhase, 0, sizeof(double) * max_prn);
    std::memset(peakMetric, 0, sizeof(double) * max_prn);
    if (dbg) std::memset(dbg, 0, sizeof(double) * max_prn * 4);
    prn_lo = std::max(prn_lo, 0);
    prn_hi = std::min(prn_hi, n_prn);

    const bool b1c = signal == BDS_SIG_B1C;
    const double fs = cfg->samplingFreq;
    const long spc = mround(fs / (cfg->codeFreqBasis / cfg->codeLength));
    long M, N;
    if (b1c) {
        M = mround((double)spc / 10 * cfg->acqCohT);          // samplesXmsLen  acquisition.m:132
        N = mround((double)spc / 10 * (10 + cfg->acqCohT));   // len10PlusXms   acquisition.m:135
    } else {
        M = spc;                                               // B2a acquisition.m:134,179
        N = 2 * spc;
    }
    if ((size_t)N > n) return set_error(BDS_ERR_ARG, "acquisition needs %ld samples, got %zu", N, n);
    if (M > spc) return set_error(BDS_ERR_UNSUPPORTED, "acqCohT > 10 ms is not supported by the reference tables");
    const int nbins = (int)mround(cfg->acqSearchBand * 2 / cfg->acqStep) + 1;
    int lgP = 1;
    while ((1L << lgP) < N + M - 1) ++lgP;
    if (lgP > 23 || lgP < 8) return set_error(BDS_ERR_UNSUPPORTED, "FFT length 2^%d out of range", lgP);
    AcqPlan pl{};
    pl.log2P = lgP;
    pl.log2P1 = lgP / 2;
    pl.log2P2 = lgP - pl.log2P1;
    pl.P = 1 << lgP;
    pl.P1 = 1 << pl.log2P1;
    pl.P2 = 1 << pl.log2P2;
    pl.N = (int)N;
    pl.M = (int)M;
    pl.Next = (int)(N + M - 1);
    const int ncodes = b1c ? (cfg->pilotACQflag == 1 ? 2 : 1) : 2;
    const int combine = b1c ? (ncodes == 2 ? 1 : 0) : 2;

    // ---- twiddle tables (double -> float)
    std::vector<float2> tw(kTwN), twHi(std::max(1, pl.P >> 11)), twLo(2048);
    const double twoPi = 6.283185307179586476925286766559;
    for (int k = 0; k < kTwN; ++k) tw[k] = make_float2((float)std::cos(twoPi * k / kTwN), (float)-std::sin(twoPi * k / kTwN));
    for (size_t k = 0; k < twHi.size(); ++k) {
        double a = twoPi * (double)(k << 11) / pl.P;
        twHi[k] = make_float2((float)std::cos(a), (float)-std::sin(a));
    }
    for (int k = 0; k < 2048; ++k) {
        double a = twoPi * (double)k / pl.P;
        twLo[k] = make_float2((float)std::cos(a), (float)-std::sin(a));
    }
    DevBuf dTw, dTwHi, dTwLo, dX, dSig, dCode, dWork, dDphi, dTab, dPeaks, dBinPeak, dSums, dFine, dFineDphi, dBits;
#define TRYA(x_)                                                                                    \
    do {                                                                                            \
        cudaError_t _e = (x_);                                                                      \
        if (_e != cudaSuccess)                                                                      \
            return set_error(_e == cudaErrorMemoryAllocation ? BDS_ERR_NOMEM : BDS_ERR_CUDA,        \
                             "acquire: %s: %s", #x_, cudaGetErrorString(_e));                       \
    } while (0)
    TRYA(dTw.alloc(tw.size() * 8));
    TRYA(dTwHi.alloc(twHi.size() * 8));
    TRYA(dTwLo.alloc(twLo.size() * 8));
    TRYA(cudaMemcpy(dTw.p, tw.data(), tw.size() * 8, cudaMemcpyHostToDevice));
    TRYA(cudaMemcpy(dTwHi.p, twHi.data(), twHi.size() * 8, cudaMemcpyHostToDevice));
    TRYA(cudaMemcpy(dTwLo.p, twLo.data(), twLo.size() * 8, cudaMemcpyHostToDevice));
    pl.tw = dTw.as<float2>();
    pl.twHi = dTwHi.as<float2>();
    pl.twLo = dTwLo.as<float2>();

    // ---- IF record on the device
    const int8_t* dx = x;
    if (x_loc == BDS_LOC_HOST) {
        TRYA(dX.alloc(n));
        TRYA(cudaMemcpy(dX.p, x, n, cudaMemcpyHostToDevice));
        dx = dX.as<int8_t>();
    }

    // ---- Doppler bins: frqBins = IF - band + step*(k-1)   acquisition.m:194-195
    std::vector<double> frq(nbins);
    std::vector<unsigned long long> dphi(nbins);
    for (int k = 0; k < nbins; ++k) {
        frq[k] = cfg->IF - cfg->acqSearchBand + cfg->acqStep * k;
        dphi[k] = freq_to_dphi(frq[k], fs);
    }
    TRYA(dDphi.alloc(sizeof(unsigned long long) * nbins));
    TRYA(cudaMemcpy(dDphi.p, dphi.data(), sizeof(unsigned long long) * nbins, cudaMemcpyHostToDevice));

    // ---- forward FFT of every bin, resident for all PRNs (the reference recomputes it per PRN)
    const size_t specBytes = sizeof(float2) * (size_t)pl.P;
    TRYA(dSig.alloc(specBytes * nbins));
    const size_t smemCol = sizeof(float2) * pl.P1 * kColTile;
    const size_t smemColInv = smemCol + sizeof(float) * pl.P1 * kColTile;
    const size_t smemRow = sizeof(float2) * pl.P2;
    TRYA(cudaFuncSetAttribute(acq_fwd_col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemCol));
    TRYA(cudaFuncSetAttribute(acq_inv_col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemColInv));
    TRYA(cudaFuncSetAttribute(acq_fwd_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemRow));
    TRYA(cudaFuncSetAttribute(acq_inv_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemRow));
    const int colGroups = pl.P2 / kColTile;
    acq_fwd_col_kernel<<<dim3(colGroups, nbins), kAcqThreads, smemCol>>>(pl, 0, dx, 0, dDphi.as<unsigned long long>(),
                                                                       dSig.as<float2>());
    acq_fwd_row_kernel<<<dim3(pl.P1, nbins), kAcqThreads, smemRow>>>(pl, dSig.as<float2>(), 0.f);
    count_launch(2);
    TRYA(cudaGetLastError());

    // ---- signal power (B1C metric normaliser)  acquisition.m:150
    double sigPower = 0;
    if (b1c) {
        TRYA(dSums.alloc(16));
        TRYA(cudaMemset(dSums.p, 0, 16));
        acq_power_kernel<<<g_num_sms * 4, 256>>>(dx, (int)M, dSums.as<long long>());
        count_launch();
        long long hs[2];
        TRYA(cudaMemcpy(hs, dSums.p, 16, cudaMemcpyDeviceToHost));
        double mean = (double)hs[0] / (double)M;
        double var = ((double)hs[1] - (double)hs[0] * mean) / (double)(M - 1);
        sigPower = std::sqrt(var * (double)M);
    }

    // ---- per-PRN buffers; work buffer sized to the free memory (bins per batch)
    TRYA(dCode.alloc(specBytes * ncodes));
    TRYA(dTab.alloc((size_t)spc * ncodes));
    size_t freeB = 0, totalB = 0;
    TRYA(cudaMemGetInfo(&freeB, &totalB));
    int binsPerBatch = (int)std::min<size_t>(nbins, std::max<size_t>(1, (freeB / 2) / (specBytes * ncodes)));
    binsPerBatch = std::min(binsPerBatch, 8);
    TRYA(dWork.alloc(specBytes * ncodes * binsPerBatch));
    TRYA(dPeaks.alloc(sizeof(AcqPeak) * (size_t)colGroups * binsPerBatch));
    TRYA(dBinPeak.alloc(sizeof(AcqPeak) * nbins));
    std::vector<AcqPeak> binPeak(nbins);
    std::vector<int8_t> tabD, tabP;

    for (int li = prn_lo; li < prn_hi; ++li) {
        const int PRN = prn[li];
        // code tables and their (conjugated, 1/P scaled) spectra
        rc = sampled_table(b1c ? BDS_CODE_B1C_DATA_BOC11 : BDS_CODE_B2A_DATA, PRN, fs, cfg->codeFreqBasis, cfg->codeLength, tabD);
        if (rc) return rc;
        TRYA(cudaMemcpy(dTab.p, tabD.data(), spc, cudaMemcpyHostToDevice));
        if (ncodes == 2) {
            rc = sampled_table(b1c ? BDS_CODE_B1C_PILOT_BOC11 : BDS_CODE_B2A_PILOT, PRN, fs, cfg->codeFreqBasis, cfg->codeLength, tabP);
            if (rc) return rc;
            TRYA(cudaMemcpy(dTab.as<int8_t>() + spc, tabP.data(), spc, cudaMemcpyHostToDevice));
        }
        acq_fwd_col_kernel<<<dim3(colGroups, ncodes), kAcqThreads, smemCol>>>(pl, 1, dTab.as<int8_t>(), (size_t)spc, nullptr,
                                                                            dCode.as<float2>());
        acq_fwd_row_kernel<<<dim3(pl.P1, ncodes), kAcqThreads, smemRow>>>(pl, dCode.as<float2>(), 1.0f / (float)pl.P);
        count_launch(2);
        for (int b0 = 0; b0 < nbins; b0 += binsPerBatch) {
            int nb = std::min(binsPerBatch, nbins - b0);
            acq_inv_row_kernel<<<dim3(pl.P1, nb, ncodes), kAcqThreads, smemRow>>>(
                pl, dSig.as<float2>() + (size_t)b0 * pl.P, dCode.as<float2>(), dWork.as<float2>(), ncodes);
            acq_inv_col_kernel<<<dim3(colGroups, nb), kAcqThreads, smemColInv>>>(pl, dWork.as<float2>(), ncodes, combine, 0, 0,
                                                                               0, 0, 0, dPeaks.as<AcqPeak>());
            acq_peak_reduce_kernel<<<nb, 256>>>(dPeaks.as<AcqPeak>(), colGroups, dBinPeak.as<AcqPeak>() + b0);
            count_launch(3);
        }
        TRYA(cudaGetLastError());
        TRYA(cudaMemcpy(binPeak.data(), dBinPeak.p, sizeof(AcqPeak) * nbins, cudaMemcpyDeviceToHost));
        // [~, bin] = max(max(results,[],2)); [peak, codePhase] = max(max(results))  (first index wins)
        float peak = -1.f;
        int bestBin = 0;
        for (int k = 0; k < nbins; ++k)
            if (binPeak[k].val > peak) {
                peak = binPeak[k].val;
                bestBin = k;
            }
        int lag = 0x7fffffff;
        for (int k = 0; k < nbins; ++k)
            if (binPeak[k].val == peak) lag = std::min(lag, binPeak[k].lag);
        long cp = (long)lag + 1;  // 1-based
        double metric, norm;
        if (b1c) {
            norm = sigPower;
            metric = (double)peak / sigPower;                       // acquisition.m:235
        } else {
            // second peak in the best bin's row, excluding +-samples2CodeChip around the peak and
            // its one-period image   (B2a acquisition.m:224-252)
            long s2cc = (long)std::ceil(fs / cfg->codeFreqBasis) * 2;
            long e1 = cp - s2cc, e2 = cp + s2cc, e3 = cp - spc + s2cc, e4 = cp + spc - s2cc;
            int lo0 = 1, hi0 = 0, lo1 = 1, hi1 = 0;  // empty
            if (e1 >= 1) {
                lo0 = (int)std::max(1L, e3) - 1;
                hi0 = (int)e1 - 1;
            }
            if (e2 < N) {
                lo1 = (int)e2 - 1;
                hi1 = (int)std::min(e4, N) - 1;
            }
            acq_inv_row_kernel<<<dim3(pl.P1, 1, ncodes), kAcqThreads, smemRow>>>(
                pl, dSig.as<float2>() + (size_t)bestBin * pl.P, dCode.as<float2>(), dWork.as<float2>(), ncodes);
            acq_inv_col_kernel<<<dim3(colGroups, 1), kAcqThreads, smemColInv>>>(pl, dWork.as<float2>(), ncodes, combine, lo0,
                                                                              hi0, lo1, hi1, 1, dPeaks.as<AcqPeak>());
            acq_peak_reduce_kernel<<<1, 256>>>(dPeaks.as<AcqPeak>(), colGroups, dBinPeak.as<AcqPeak>());
            count_launch(3);
            AcqPeak second;
            TRYA(cudaMemcpy(&second, dBinPeak.p, sizeof(AcqPeak), cudaMemcpyDeviceToHost));
            norm = second.val;
            metric = (double)peak / (double)second.val;
        }
        peakMetric[PRN - 1] = metric;
        if (b1c && cp + spc - 1 > (long)n) cp -= spc;               // acquisition.m:239-241
        if (dbg) {
            dbg[(PRN - 1) * 4 + 0] = bestBin;
            dbg[(PRN - 1) * 4 + 1] = (double)cp;
            dbg[(PRN - 1) * 4 + 2] = peak;
            dbg[(PRN - 1) * 4 + 3] = norm;
        }
        if (!(metric > cfg->acqThreshold)) continue;
        if (cp < 1) continue;  // the reference would index longSignal(<=0) here and abort

        if (b1c) {
            // ---- fine search, acquisition.m:253-307
            if ((size_t)(cp - 1 + spc) > n) continue;
            const int nfine = (int)mround(cfg->acqStep / 25) * 2 + 1;
            std::vector<double> ff(nfine);
            std::vector<unsigned long long> fd(nfine);
            for (int j = 0; j < nfine; ++j) {
                ff[j] = frq[bestBin] - cfg->acqStep + 25.0 * j;
                fd[j] = freq_to_dphi(ff[j], fs);
            }
            TRYA(dFineDphi.alloc(sizeof(unsigned long long) * nfine));
            TRYA(cudaMemcpy(dFineDphi.p, fd.data(), sizeof(unsigned long long) * nfine, cudaMemcpyHostToDevice));
            TRYA(dFine.alloc(sizeof(double) * nfine * ncodes * 4));
            TRYA(cudaMemset(dFine.p, 0, sizeof(double) * nfine * ncodes * 4));
            TRYA(cudaMemset(dSums.p, 0, 16));
            acq_power_kernel<<<g_num_sms * 4, 256>>>(dx + (cp - 1), (int)spc, dSums.as<long long>());
            acq_fine_b1c_kernel<<<dim3(g_num_sms, nfine, ncodes), 256>>>(dx + (cp - 1), dTab.as<int8_t>(), (int)spc,
                                                                        dFineDphi.as<unsigned long long>(), ncodes,
                                                                        dFine.as<double>());
            count_launch(2);
            long long hs[2];
            TRYA(cudaMemcpy(hs, dSums.p, 16, cudaMemcpyDeviceToHost));
            std::vector<double> fo((size_t)nfine * ncodes * 4);
            TRYA(cudaMemcpy(fo.data(), dFine.p, fo.size() * 8, cudaMemcpyDeviceToHost));
            const double mean = (double)hs[0] / (double)spc;
            int best = 0;
            double bestV = -1;
            for (int j = 0; j < nfine; ++j) {
                double v[2] = {0, 0};
                for (int dp = 0; dp < ncodes; ++dp) {
                    const double* o = &fo[((size_t)j * ncodes + dp) * 4];
                    v[dp] = std::hypot(o[0] - mean * o[2], o[1] - mean * o[3]);
                }
                double r = ncodes == 2 ? (v[0] * 11 + v[1] * 29) / 40 : v[0];
                if (r > bestV) {
                    bestV = r;
                    best = j;
                }
            }
            carrFreq[PRN - 1] = ff[best];
        } else {
            // ---- fine search, B2a acquisition.m:256-335
            const int nseg = cfg->fineNoncoh;
            if (nseg <= 0 || (size_t)(cp - 1 + (long)nseg * spc) > n) continue;
            const int nfine = (int)mround(cfg->acqStep / 25) + 1;
            std::vector<double> ff(nfine);
            std::vector<unsigned long long> fd(nfine);
            for (int j = 0; j < nfine; ++j) {
                ff[j] = frq[bestBin] - cfg->acqStep / 2 + 25.0 * j;
                fd[j] = freq_to_dphi(ff[j], fs);
            }
            std::vector<uint32_t> bits(2 * kPackedWords);
            std::vector<uint8_t> chips;
            primary_bits(BDS_CODE_B2A_DATA, PRN, chips);
            pack_bits(chips, bits.data());
            primary_bits(BDS_CODE_B2A_PILOT, PRN, chips);
            pack_bits(chips, bits.data() + kPackedWords);
            TRYA(dBits.alloc(bits.size() * 4));
            TRYA(cudaMemcpy(dBits.p, bits.data(), bits.size() * 4, cudaMemcpyHostToDevice));
            TRYA(dFineDphi.alloc(sizeof(unsigned long long) * nfine));
            TRYA(cudaMemcpy(dFineDphi.p, fd.data(), sizeof(unsigned long long) * nfine, cudaMemcpyHostToDevice));
            const size_t no = (size_t)nfine * 2 * nseg * 2;
            TRYA(dFine.alloc(sizeof(double) * no));
            TRYA(cudaMemset(dFine.p, 0, sizeof(double) * no));
            acq_fine_b2a_kernel<<<dim3(32, nfine, nseg), 256>>>(dx + (cp - 1), dBits.as<uint32_t>(), (int)spc, 1.0 / fs,
                                                               1.0 / cfg->codeFreqBasis,
                                                               dFineDphi.as<unsigned long long>(), nseg, dFine.as<double>());
            count_launch();
            std::vector<double> fo(no);
            TRYA(cudaMemcpy(fo.data(), dFine.p, no * 8, cudaMemcpyDeviceToHost));
            int best = 0;
            double bestV = -1;
            for (int j = 0; j < nfine; ++j) {
                double r = 0;
                for (int dp = 0; dp < 2; ++dp)
                    for (int s = 0; s < nseg; ++s) {
                        const double* o = &fo[(((size_t)j * 2 + dp) * nseg + s) * 2];
                        r += std::hypot(o[0], o[1]);
                    }
                if (r > bestV) {
                    bestV = r;
                    best = j;
                }
            }
            carrFreq[PRN - 1] = ff[best];
        }
        if (carrFreq[PRN - 1] == 0) carrFreq[PRN - 1] = 1;          // acquisition.m:303-305
        codePhase[PRN - 1] = (double)cp;
    }
    TRYA(cudaDeviceSynchronize());
#undef TRYA
    return BDS_OK;
}

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
#include <map>
#include <mutex>
#include <vector>

#include "bds_codes.h"
#include "bds_common.cuh"

namespace bds {

constexpr int kTwN = 4096;           // butterfly twiddle table size (covers sub-FFTs <= 4096)
constexpr int kColTile = 8;          // columns per CTA in the column passes
constexpr int kAcqThreads = 512;
constexpr int kLgColTile = 3;        // log2(kColTile)

struct AcqPlan {
    int log2P, log2P1, log2P2;   // P = P1*P2 ; P1 = column (strided) length, P2 = row length
    int lgRowTile;               // log2 of the rows per CTA in the row passes
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

// ---- shared-memory FFT ---------------------------------------------------------------
// In-place FFT over len = 2^lg elements for 2^lgBatch independent sequences held in shared memory.
// kInv=false: DIF forward (natural -> bit reversed).  kInv=true: DIT inverse (bit reversed -> natural), unnormalised.
// Up to four radix-2 stages are fused into one step on 16 elements held in registers, so a 2048-point
// transform makes three trips through shared memory (4 + 4 + 3 stages) instead of eleven.  Within a step the
// twiddle of a butterfly factors into a per-thread part W_{2h}^j (table lookup, one per stage) and a
// compile-time 16th root of unity.
// Layout: element i of a sequence sits at the padded index pidx(i) = i + (i >> 4), which makes both the
// stride-1 steps (a thread owns 16 consecutive elements) and the strided ones bank-conflict free:
//   kCols = false (rows)   : element (i, b) at buf[b * seqStride + pidx(i)]
//   kCols = true  (columns): element (i, b) at buf[pidx(i) * 2^lgBatch + b]
__device__ __forceinline__ int pidx(int i) { return i + (i >> 4); }

// x * exp(-2 pi i k / 16), k = 0..7; k is a compile-time constant after unrolling, the switch folds away
__device__ __forceinline__ float2 mul_w16(float2 x, int k) {
    constexpr float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, r2 = 0.70710678118654752f;
    switch (k) {
        case 0: return x;
        case 1: return make_float2(x.x * c1 + x.y * s1, x.y * c1 - x.x * s1);
        case 2: return make_float2((x.x + x.y) * r2, (x.y - x.x) * r2);
        case 3: return make_float2(x.x * s1 + x.y * c1, x.y * s1 - x.x * c1);
        case 4: return make_float2(x.y, -x.x);
        case 5: return make_float2(x.y * c1 - x.x * s1, -x.y * s1 - x.x * c1);
        case 6: return make_float2((x.y - x.x) * r2, -(x.x + x.y) * r2);
        default: return make_float2(x.y * s1 - x.x * c1, -x.y * c1 - x.x * s1);   // 7
    }
}
// x * exp(+2 pi i k / 16)
__device__ __forceinline__ float2 mul_w16c(float2 x, int k) {
    const float2 y = mul_w16(make_float2(x.x, -x.y), k);   // conj(conj(x) * w) = x * conj(w)
    return make_float2(y.x, -y.y);
}

// Q fused radix-2 stages on N = 2^Q register values; w[l] = W_{2 h_l}^j of stage l
template <bool kInv, int Q>
__device__ __forceinline__ void fft_regs(float2* r, const float2* w) {
    constexpr int N = 1 << Q;
#pragma unroll
    for (int l = 0; l < Q; ++l) {
        const int dist = kInv ? (1 << l) : (N >> (l + 1));
#pragma unroll
        for (int m = 0; m < N; ++m) {
            if (m & dist) continue;
            const int k = (m & (dist - 1)) * (8 / dist);     // exp(-+2 pi i mp / (2 dist)) as a 16th root of unity
            const float2 a = r[m], b = r[m + dist];
            if (!kInv) {
                r[m] = make_float2(a.x + b.x, a.y + b.y);
                r[m + dist] = cmul(mul_w16(make_float2(a.x - b.x, a.y - b.y), k), w[l]);
            } else {
                const float2 t = cmulc(mul_w16c(b, k), w[l]);
                r[m] = make_float2(a.x + t.x, a.y + t.y);
                r[m + dist] = make_float2(a.x - t.x, a.y - t.y);
            }
        }
    }
}

// one fused step of Q stages; `done` stages have been applied before it
template <bool kInv, bool kCols, int Q>
__device__ __forceinline__ void fft_step(float2* buf, int lg, int lgBatch, int seqStride, int done, const float2* tw) {
    constexpr int N = 1 << Q;
    const int lgItems = lg - Q;                       // groups per sequence
    const int items = 1 << (lgItems + lgBatch);
    for (int u = threadIdx.x; u < items; u += blockDim.x) {
        int b, v;
        if (kCols) {
            b = u & ((1 << lgBatch) - 1);
            v = u >> lgBatch;
        } else {
            b = u >> lgItems;
            v = u & ((1 << lgItems) - 1);
        }
        int i, lstride, j, lh;
        if (!kInv) {     // DIF: half spans h, h/2, .. ; elements i + m * (h >> (Q-1))
            lh = lg - 1 - done;
            lstride = lh - (Q - 1);
            j = v & ((1 << lstride) - 1);
            i = ((v >> lstride) << (lh + 1)) + j;
        } else {         // DIT: half spans h, 2h, .. ; elements i + m * h
            lh = done;
            lstride = lh;
            j = v & ((1 << lh) - 1);
            i = ((v >> lh) << (lh + Q)) + j;
        }
        float2 w[Q];
#pragma unroll
        for (int l = 0; l < Q; ++l) {
            const int l2h = kInv ? (lh + l + 1) : (lh - l + 1);          // log2 of the full span of stage l
            w[l] = __ldg(tw + ((unsigned)j << (12 - l2h)));               // W_{2h_l}^j, kTwN = 4096
        }
        float2 r[N];
        float2* base = kCols ? buf + b : buf + b * seqStride;
#pragma unroll
        for (int m = 0; m < N; ++m) {
            const int e = pidx(i + (m << lstride));
            r[m] = kCols ? base[e << lgBatch] : base[e];
        }
        fft_regs<kInv, Q>(r, w);
#pragma unroll
        for (int m = 0; m < N; ++m) {
            const int e = pidx(i + (m << lstride));
            if (kCols) base[e << lgBatch] = r[m];
            else base[e] = r[m];
        }
    }
}

template <bool kInv, bool kCols>
__device__ void smem_fft(float2* buf, int lg, int lgBatch, int seqStride, const float2* tw) {
    int done = 0;
    while (done < lg) {
        const int q = min(4, lg - done);
        if (q == 4) fft_step<kInv, kCols, 4>(buf, lg, lgBatch, seqStride, done, tw);
        else if (q == 3) fft_step<kInv, kCols, 3>(buf, lg, lgBatch, seqStride, done, tw);
        else if (q == 2) fft_step<kInv, kCols, 2>(buf, lg, lgBatch, seqStride, done, tw);
        else fft_step<kInv, kCols, 1>(buf, lg, lgBatch, seqStride, done, tw);
        done += q;
        __syncthreads();
    }
}

// Global loads of the FFT kernels are issued in batches of kLdBatch per thread before the first one is used: a plain
// "load, use, next element" loop keeps one or two loads in flight per thread and stalls on every one of them
// (ncu, round 2: 44 % of the stall samples of the inverse passes sat on the instruction after such a load).
constexpr int kLdBatch = 8;

// ---- forward column pass -------------------------------------------------------------
// kind 0: signal  y[n] = x[n mod N] * exp(+i 2 pi f n'/fs), n' = n mod N, n < N+M-1, else 0
//                 (acquisition.m:194-205; periodic extension see file header)
// kind 1: code    y[n] = table[n] for n < M else 0   (acquisition.m:176-180)
// grid = (P2/kColTile, batch); batch index selects the Doppler bin / code table.
// sample m of the record as (I, Q).  fmt bit 0: interleaved I/Q pairs (settings.fileType == 2: longSignal = I + 1i*Q,
// postProcessing.m:96-99); fmt bit 1: float samples (the output of the resampling pre-conditioner) instead of int8
enum { kFmtI8 = 0, kFmtI8IQ = 1, kFmtF32 = 2, kFmtF32IQ = 3 };
__device__ __forceinline__ float2 acq_sample(const int8_t* x, size_t m, int fmt) {
    if (fmt == kFmtI8) return make_float2((float)x[m], 0.f);
    if (fmt == kFmtI8IQ) {
        const char2 v = reinterpret_cast<const char2*>(x)[m];
        return make_float2((float)v.x, (float)v.y);
    }
    if (fmt == kFmtF32) return make_float2(reinterpret_cast<const float*>(x)[m], 0.f);
    return reinterpret_cast<const float2*>(x)[m];
}
__host__ __device__ inline size_t acq_sample_bytes(int fmt) { return (fmt & 2 ? 4 : 1) << (fmt & 1); }

__global__ void __launch_bounds__(kAcqThreads, 2) acq_fwd_col_kernel(AcqPlan pl, int kind, const int8_t* src,
                                                                  size_t srcStride, const unsigned long long* dphi,
                                                                  float2* spec, int iq) {
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
                const float2 xs = acq_sample(x, m, iq);
                unsigned long long ph = (unsigned long long)m * dp;
                float sn, cs;
                sincospif((float)(int)(ph >> 32) * 4.656612873077392578125e-10f, &sn, &cs);
                v = make_float2(xs.x * cs - xs.y * sn, xs.x * sn + xs.y * cs);   // sigCarr .* sig, acquisition.m:198-205
            }
        } else if (n < (unsigned)pl.M) {
            v.x = (float)x[n];
        }
        buf[pidx(r) * kColTile + c] = v;
    }
    __syncthreads();
    smem_fft<false, true>(buf, pl.log2P1, kLgColTile, 0, pl.tw);
    float2* out = spec + (size_t)bi * pl.P;
    for (int e0 = threadIdx.x; e0 < total; e0 += blockDim.x * kLdBatch) {
        float2 hi[kLdBatch], lo[kLdBatch];
#pragma unroll
        for (int u = 0; u < kLdBatch; ++u) {
            const int e = e0 + u * blockDim.x;
            if (e < total) {
                const unsigned m = (__brev((unsigned)(e >> 3)) >> (32 - pl.log2P1)) * (unsigned)(col0 + (e & (kColTile - 1)));
                hi[u] = __ldg(pl.twHi + (m >> 11));
                lo[u] = __ldg(pl.twLo + (m & 2047));
            }
        }
#pragma unroll
        for (int u = 0; u < kLdBatch; ++u) {
            const int e = e0 + u * blockDim.x;
            if (e < total) {
                const int c = e & (kColTile - 1), r = e >> 3;
                out[(size_t)r * pl.P2 + col0 + c] = cmul(buf[pidx(r) * kColTile + c], cmul(hi[u], lo[u]));
            }
        }
    }
}

// ---- forward row pass (in place) -----------------------------------------------------
// conjScale != 0: store conj(X) * conjScale (code spectra: acquisition.m:180 with the 1/P of
// the inverse transform folded in).  grid = (P1 >> lgRowTile, batch).
// perm = 1: scrambled element 16 v + u of a row is stored at position u * (P2 / 16) + v (the layout the compile-time
// specialised inverse row pass reads, see acq_inv_row_ct_kernel); stores stay contiguous, the shared-memory reads stride.
__global__ void __launch_bounds__(kAcqThreads, 2) acq_fwd_row_kernel(AcqPlan pl, float2* spec, float conjScale, int perm) {
    extern __shared__ __align__(16) unsigned char smraw[];
    float2* buf = reinterpret_cast<float2*>(smraw);
    float2* rows = spec + (size_t)blockIdx.y * pl.P + ((size_t)blockIdx.x << pl.lgRowTile) * pl.P2;   // 2^lgRowTile contiguous rows
    const int total = pl.P2 << pl.lgRowTile, rowStride = pl.P2 + (pl.P2 >> 4);
    for (int i0 = threadIdx.x; i0 < total; i0 += blockDim.x * kLdBatch) {
        float2 v[kLdBatch];
#pragma unroll
        for (int u = 0; u < kLdBatch; ++u) {
            const int i = i0 + u * blockDim.x;
            if (i < total) v[u] = rows[i];
        }
#pragma unroll
        for (int u = 0; u < kLdBatch; ++u) {
            const int i = i0 + u * blockDim.x;
            if (i < total) buf[(i >> pl.log2P2) * rowStride + pidx(i & (pl.P2 - 1))] = v[u];
        }
    }
    __syncthreads();
    smem_fft<false, false>(buf, pl.log2P2, pl.lgRowTile, rowStride, pl.tw);
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int d = i & (pl.P2 - 1);
        const int src = perm ? ((d & ((pl.P2 >> 4) - 1)) << 4) + (d >> (pl.log2P2 - 4)) : d;
        float2 v = buf[(i >> pl.log2P2) * rowStride + pidx(src)];
        if (conjScale != 0.f) v = make_float2(v.x * conjScale, -v.y * conjScale);
        rows[i] = v;
    }
}

// ---- inverse row pass ------------------------------------------------------------------
// work[bin][dp] row r = IDFT_row( sig[bin] row r .* code[dp] row r ) .* W_P^{-n2 k1}
// grid = (P1 >> lgRowTile, nbins, ncodes); binMap (optional): Doppler bin of batch entry blockIdx.y
__global__ void __launch_bounds__(kAcqThreads, 2) acq_inv_row_kernel(AcqPlan pl, const float2* sig, const float2* code,
                                                                  float2* work, int ncodes, const int* binMap) {
    extern __shared__ __align__(16) unsigned char smraw[];
    float2* buf = reinterpret_cast<float2*>(smraw);
    const int r0 = blockIdx.x << pl.lgRowTile, bi = blockIdx.y, dp = blockIdx.z;
    const int bin = binMap ? binMap[bi] : bi;
    const float2* srow = sig + (size_t)bin * pl.P + (size_t)r0 * pl.P2;
    const float2* crow = code + (size_t)dp * pl.P + (size_t)r0 * pl.P2;
    const int total = pl.P2 << pl.lgRowTile, rowStride = pl.P2 + (pl.P2 >> 4);
    for (int i0 = threadIdx.x; i0 < total; i0 += blockDim.x * kLdBatch) {
        float2 a[kLdBatch], c[kLdBatch];
#pragma unroll
        for (int u = 0; u < kLdBatch; ++u) {
            const int i = i0 + u * blockDim.x;
            if (i < total) {
                a[u] = srow[i];
                c[u] = __ldg(crow + i);
            }
        }
#pragma unroll
        for (int u = 0; u < kLdBatch; ++u) {
            const int i = i0 + u * blockDim.x;
            if (i < total) buf[(i >> pl.log2P2) * rowStride + pidx(i & (pl.P2 - 1))] = cmul(a[u], c[u]);
        }
    }
    __syncthreads();
    smem_fft<true, false>(buf, pl.log2P2, pl.lgRowTile, rowStride, pl.tw);
    float2* orow = work + ((size_t)bi * ncodes + dp) * pl.P + (size_t)r0 * pl.P2;
    for (int i0 = threadIdx.x; i0 < total; i0 += blockDim.x * kLdBatch) {
        float2 hi[kLdBatch], lo[kLdBatch];
#pragma unroll
        for (int u = 0; u < kLdBatch; ++u) {     // the two factors of W_P^{k1 n2} (twiddleP), all fetched before the first use
            const int i = i0 + u * blockDim.x;
            if (i < total) {
                const int r = r0 + (i >> pl.log2P2), n2 = i & (pl.P2 - 1);
                const unsigned m = (__brev((unsigned)r) >> (32 - pl.log2P1)) * (unsigned)n2;
                hi[u] = __ldg(pl.twHi + (m >> 11));
                lo[u] = __ldg(pl.twLo + (m & 2047));
            }
        }
#pragma unroll
        for (int u = 0; u < kLdBatch; ++u) {
            const int i = i0 + u * blockDim.x;
            if (i < total) orow[i] = cmulc(buf[(i >> pl.log2P2) * rowStride + pidx(i & (pl.P2 - 1))], cmul(hi[u], lo[u]));
        }
    }
}

// ---- inverse column pass + magnitude + combine + max -------------------------------------
struct AcqPeak {   // 8 bytes (the specialised column pass reuses the slot for a packed 64-bit maximum)
    float val;
    int lag;  // 0-based
};
// combine: 0 data only; 1 B1C (|d| sqrt11 + |p| sqrt29)/sqrt40 (acquisition.m:215-220);
//          2 B2a |d|+|p| (B2a acquisition.m:205-209)
// Lags are restricted to [0,N) and, when ex0 <= ex1 (second-peak search, B2a
// acquisition.m:224-249, 0-based inclusive bounds), to [lo0,hi0] U [lo1,hi1].
__global__ void __launch_bounds__(kAcqThreads, 2) acq_inv_col_kernel(AcqPlan pl, const float2* work, int ncodes,
                                                                  int combine, int lo0, int hi0, int lo1, int hi1,
                                                                  int useRanges, AcqPeak* peaks /*[bin][gridDim.x]*/) {
    extern __shared__ __align__(16) unsigned char smraw[];
    float2* buf = reinterpret_cast<float2*>(smraw);
    const int total = pl.P1 * kColTile;
    float* mag = reinterpret_cast<float*>(buf + (pl.P1 + (pl.P1 >> 4)) * kColTile);   // after the padded FFT buffer
    const int col0 = blockIdx.x * kColTile, bin = blockIdx.y;
    for (int dp = 0; dp < ncodes; ++dp) {
        const float2* w = work + ((size_t)bin * ncodes + dp) * pl.P;
        for (int e0 = threadIdx.x; e0 < total; e0 += blockDim.x * kLdBatch) {
            float2 v[kLdBatch];
#pragma unroll
            for (int u = 0; u < kLdBatch; ++u) {
                const int e = e0 + u * blockDim.x;
                if (e < total) v[u] = w[(size_t)(e >> 3) * pl.P2 + col0 + (e & (kColTile - 1))];
            }
#pragma unroll
            for (int u = 0; u < kLdBatch; ++u) {
                const int e = e0 + u * blockDim.x;
                if (e < total) buf[pidx(e >> 3) * kColTile + (e & (kColTile - 1))] = v[u];
            }
        }
        __syncthreads();
        smem_fft<true, true>(buf, pl.log2P1, kLgColTile, 0, pl.tw);
        for (int e = threadIdx.x; e < total; e += blockDim.x) {
            const int c = e & (kColTile - 1), r = e >> 3;
            float2 v = buf[pidx(r) * kColTile + c];
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

// =====================================================================================================
// Inverse passes specialised at compile time for the transform shapes of the shipped configurations.
// The PRN x Doppler grid spends > 95 % of an acquisition in the two inverse passes (25 326 transforms of 2^22 points
// for the 63-PRN B1C grid), and ncu (profiles/r02/acq_inv_*_after.csv) showed them bound by instruction issue at
// 160-180 thread-instructions per point, not by HBM.  Three things remove more than half of those instructions:
//   * every length, stride and stage count is a template parameter, so the index arithmetic of the padded layout folds
//     into immediate offsets (16 LDS/STS from one base address per fused step);
//   * a fused step is a plain radix-2^Q butterfly (constant 16th roots of unity only) applied to inputs that were
//     multiplied once by their step twiddle exp(+2 pi i j m' / 2^(D+Q)), read from a table laid out [m'][j] so that a
//     warp's loads are contiguous - instead of one general complex multiply per butterfly and stage; the first step
//     (D = 0) has no twiddles at all;
//   * the magnitude / combine / max epilogue works on the rows that can hold a lag < N only (less than half of them:
//     N < P/2), with MUFU square roots and the (sqrt11, sqrt29) / sqrt40 weights as two constants.
// Same scrambled spectrum layout, same padded shared-memory layout (pidx) and same results to float rounding as the
// generic kernels above, which remain the path of every other transform shape.
template <int Q>
__host__ __device__ constexpr int brevq(int m) {
    int r = 0;
    for (int k = 0; k < Q; ++k) r |= ((m >> k) & 1) << (Q - 1 - k);
    return r;
}
// Steps of a 2^LG-point inverse: D = 0, 4, 8, .. radix-2 stages done before the step, Q = min(4, LG - D) fused.
// Twiddle table of step D > 0: T[(m' - 1) * 2^D + j], m' = 1 .. 2^Q - 1, j < 2^D; the tables of a transform are
// concatenated in step order.
__host__ __device__ constexpr int inv_tw_offset(int LG, int D) {
    int off = 0;
    for (int d = 4; d < D; d += 4) off += (((LG - d) < 4 ? (1 << (LG - d)) : 16) - 1) << d;
    return off;
}
__host__ __device__ constexpr int inv_tw_size(int LG) { return inv_tw_offset(LG, ((LG + 3) / 4) * 4); }

// Q fused inverse DIT stages on 2^Q register values (bit-reversed in, natural out), constant twiddles only
template <int Q>
__device__ __forceinline__ void ifft_regs_const(float2* r) {
    constexpr int N = 1 << Q;
#pragma unroll
    for (int l = 0; l < Q; ++l) {
        const int dist = 1 << l;
#pragma unroll
        for (int m = 0; m < N; ++m) {
            if (m & dist) continue;
            const int k = (m & (dist - 1)) * (8 / dist);
            const float2 a = r[m], t = mul_w16c(r[m + dist], k);
            r[m] = make_float2(a.x + t.x, a.y + t.y);
            r[m + dist] = make_float2(a.x - t.x, a.y - t.y);
        }
    }
}

template <int LG, int LGB, bool kCols, int SEQ, int NT, int D, int Q>
__device__ __forceinline__ void ifft_step_ct(float2* buf, const float2* __restrict__ T) {
    constexpr int N = 1 << Q, h = 1 << D, lgItems = LG - Q, items = 1 << (lgItems + LGB);
    constexpr bool lin = (D >= 4) || (D == 0 && Q == 4);   // the padded offsets of the step's elements do not depend on i
#pragma unroll 1
    for (int u = threadIdx.x; u < items; u += NT) {
        const int b = kCols ? (u & ((1 << LGB) - 1)) : (u >> lgItems);
        const int v = kCols ? (u >> LGB) : (u & ((1 << lgItems) - 1));
        const int j = v & (h - 1);
        const int i = ((v >> D) << (D + Q)) + j;
        float2* base = kCols ? buf + (pidx(i) << LGB) + b : buf + b * SEQ + pidx(i);
        float2 r[N];
#pragma unroll
        for (int m = 0; m < N; ++m) {
            const int off = lin ? ((m << D) + ((m << D) >> 4)) : (pidx(i + (m << D)) - pidx(i));
            r[m] = base[kCols ? (off << LGB) : off];
        }
        if (D > 0) {
#pragma unroll
            for (int m = 1; m < N; ++m) r[m] = cmul(r[m], __ldg(T + (brevq<Q>(m) - 1) * h + j));
        }
        ifft_regs_const<Q>(r);
#pragma unroll
        for (int m = 0; m < N; ++m) {
            const int off = lin ? ((m << D) + ((m << D) >> 4)) : (pidx(i + (m << D)) - pidx(i));
            base[kCols ? (off << LGB) : off] = r[m];
        }
    }
}
template <int LG, int LGB, bool kCols, int SEQ, int NT, int D = 0>
__device__ __forceinline__ void ifft_ct(float2* buf, const float2* __restrict__ tws) {
    if constexpr (D < LG) {
        constexpr int Q = (LG - D) < 4 ? (LG - D) : 4;
        ifft_step_ct<LG, LGB, kCols, SEQ, NT, D, Q>(buf, tws + inv_tw_offset(LG, D));
        __syncthreads();
        ifft_ct<LG, LGB, kCols, SEQ, NT, D + Q>(buf, tws);
    }
}

// L2 residency hints of the specialised inverse passes: the signal and code spectra stream through once per cell
// (evict first), the row pass's output is read back by the column pass right after (evict last) and dead after that
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
    unsigned long long p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
    unsigned long long p;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ float2 ld_hint(const float2* p, unsigned long long pol) {
    float2 v;
    asm volatile("ld.global.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void st_hint(float2* p, float2 v, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// one thread per 16-element group of the first step
constexpr int inv_row_threads(int LG2, int LGRT) { return (1 << (LG2 - 4)) << LGRT; }
constexpr int inv_col_threads(int LG1) { return (1 << (LG1 - 4)) * kColTile; }
constexpr int inv_last_d(int LG) { return ((LG - 1) / 4) * 4; }   // stages done before the last fused step

// The first and the last step of a pass never touch shared memory on their outer side:
//   * the first step (D = 0: 16 adjacent elements, no twiddles) takes its inputs straight from global memory.  For the
//     row pass the spectra are stored with the 16 elements of a group 2^(LG2-4) apart (acq_fwd_row_kernel, perm = 1:
//     scrambled element 16 v + u of a row at position u * 2^(LG2-4) + v), so that the loads of a warp are contiguous;
//     for the column pass the elements of a group are rows, a warp reads 64-byte row segments as before;
//   * the last step leaves its outputs (elements j + m * 2^D, contiguous in j across a warp) in registers, where the
//     row pass multiplies them with the four-step twiddle and writes them to `work`, and the column pass takes the
//     magnitude, combines data and pilot and keeps the running maximum.
// Shared-memory traffic per point and pass drops from 64 to 32 bytes, the barriers from four to two.

// inverse row pass, see acq_inv_row_kernel.  grid = ((P1 >> LGRT) * ncodes, nbins): the CTAs of the codes that read the
// same signal rows are neighbours in launch order (the second read of the rows hits L2).
template <int LG1, int LG2, int LGRT>
__global__ void __launch_bounds__(inv_row_threads(LG2, LGRT), 1024 / inv_row_threads(LG2, LGRT))
acq_inv_row_ct_kernel(AcqPlan pl, const float2* __restrict__ sig, const float2* __restrict__ code, float2* __restrict__ work,
                      int ncodes, const int* __restrict__ binMap, const float2* __restrict__ twRow) {
    constexpr int NT = inv_row_threads(LG2, LGRT), P2 = 1 << LG2, rowStride = P2 + (P2 >> 4), G = P2 >> 4;
    constexpr int DL = inv_last_d(LG2), QL = LG2 - DL, NL = 1 << QL, HL = 1 << DL;
    constexpr int itemsL = (P2 >> QL) << LGRT;   // groups of the last step
    static_assert(NT >= 64 && NT <= 512 && DL >= 4, "row tiling");
    extern __shared__ __align__(16) unsigned char smraw[];
    float2* buf = reinterpret_cast<float2*>(smraw);
    __shared__ float2 sS[1 << LGRT][NL];          // W_P^{k1 m 2^DL}: the four-step twiddle's advance from output m-1 to m
    const int tid = threadIdx.x;
    const int dp = blockIdx.x % ncodes, rt = blockIdx.x / ncodes, r0 = rt << LGRT, bi = blockIdx.y;
    const int bin = binMap ? binMap[bi] : bi;
    const unsigned long long polFirst = l2_policy_evict_first(), polLast = l2_policy_evict_last();
    if (tid < (NL << LGRT)) {
        const int row = tid >> QL, m = tid & (NL - 1);
        const unsigned k1 = __brev((unsigned)(r0 + row)) >> (32 - LG1);
        sS[row][m] = twiddleP(pl, (k1 * (unsigned)m) << DL);
    }
    {   // ---- first step from global memory
        const int b = tid >> (LG2 - 4), v = tid & (G - 1);
        const float2* srow = sig + (size_t)bin * pl.P + (size_t)(r0 + b) * P2 + v;
        const float2* crow = code + (size_t)dp * pl.P + (size_t)(r0 + b) * P2 + v;
        float2 r[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) r[u] = ld_hint(srow + u * G, polFirst);
#pragma unroll
        for (int u0 = 0; u0 < 16; u0 += 8) {
            float2 c[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) c[u] = ld_hint(crow + (u0 + u) * G, polFirst);
#pragma unroll
            for (int u = 0; u < 8; ++u) r[u0 + u] = cmul(r[u0 + u], c[u]);
        }
        ifft_regs_const<4>(r);
        float2* base = buf + b * rowStride + 17 * v;     // pidx(16 v + m) = 17 v + m
#pragma unroll
        for (int m = 0; m < 16; ++m) base[m] = r[m];
    }
    __syncthreads();
    // ---- middle steps in shared memory
    if constexpr (DL > 4) {
        ifft_step_ct<LG2, LGRT, false, rowStride, NT, 4, 4>(buf, twRow + inv_tw_offset(LG2, 4));
        __syncthreads();
    }
    static_assert(DL <= 8, "row transforms up to 4096 points");
    // ---- last step: outputs to global memory
    const float2* T = twRow + inv_tw_offset(LG2, DL);
#pragma unroll 1
    for (int u = tid; u < itemsL; u += NT) {
        const int b = u >> DL, j = u & (HL - 1);
        const float2* base = buf + b * rowStride + pidx(j);
        float2 r[NL];
#pragma unroll
        for (int m = 0; m < NL; ++m) r[m] = base[(m << DL) + ((m << DL) >> 4)];
#pragma unroll
        for (int m = 1; m < NL; ++m) r[m] = cmul(r[m], __ldg(T + (brevq<QL>(m) - 1) * HL + j));
        ifft_regs_const<QL>(r);
        // W_P^{k1 n2}, k1 = bitrev(row), n2 = j + m 2^DL
        const unsigned k1 = __brev((unsigned)(r0 + b)) >> (32 - LG1);
        const float2 t0 = twiddleP(pl, k1 * (unsigned)j);
        float2* orow = work + ((size_t)bi * ncodes + dp) * pl.P + (size_t)(r0 + b) * P2 + j;
        st_hint(orow, cmulc(r[0], t0), polLast);
#pragma unroll
        for (int m = 1; m < NL; ++m) st_hint(orow + (m << DL), cmulc(r[m], cmul(t0, sS[b][m])), polLast);
    }
}

// inverse column pass + magnitude + combine + max, see acq_inv_col_kernel
template <int LG1, int LG2>
__global__ void __launch_bounds__(inv_col_threads(LG1), 1024 / inv_col_threads(LG1))
acq_inv_col_ct_kernel(AcqPlan pl, const float2* __restrict__ work, int ncodes, int combine, int lo0, int hi0, int lo1, int hi1,
                      int useRanges, AcqPeak* __restrict__ peaks /*[bin]: packed, see below*/, const float2* __restrict__ twCol) {
    constexpr int NT = inv_col_threads(LG1), P1 = 1 << LG1, P2 = 1 << LG2;
    constexpr int DL = inv_last_d(LG1), QL = LG1 - DL, NL = 1 << QL, HL = 1 << DL;
    constexpr int itemsL = (P1 >> QL) * kColTile, itersL = itemsL / NT;
    static_assert(NT >= 64 && NT <= 512 && DL >= 4 && DL <= 8 && itemsL % NT == 0, "column tiling");
    extern __shared__ __align__(16) unsigned char smraw[];
    float2* buf = reinterpret_cast<float2*>(smraw);
    float* mag = reinterpret_cast<float*>(buf + (P1 + (P1 >> 4)) * kColTile);   // [itersL][NL][NT]
    const int tid = threadIdx.x, c = tid & (kColTile - 1);
    const int col0 = blockIdx.x * kColTile, bin = blockIdx.y;
    // rows that can hold a lag < N (lag = r * P2 + col): r < rEnd
    const int rEnd = min(P1, (pl.N - col0 + P2 - 1) >> LG2);
    const float2* T = twCol + inv_tw_offset(LG1, DL);
    const unsigned long long polFirst = l2_policy_evict_first();
    float best = -1.f;
    int bl = 0x7fffffff;
    for (int dp = 0; dp < ncodes; ++dp) {
        {   // ---- first step from global memory: rows 16 v .. 16 v + 15 of column c
            const int v = tid >> kLgColTile;
            const float2* w = work + ((size_t)bin * ncodes + dp) * pl.P + (size_t)(16 * v) * P2 + col0 + c;
            float2 r[16];
#pragma unroll
            for (int m = 0; m < 16; ++m) r[m] = ld_hint(w + (size_t)m * P2, polFirst);
            ifft_regs_const<4>(r);
            float2* base = buf + 17 * v * kColTile + c;
#pragma unroll
            for (int m = 0; m < 16; ++m) base[m * kColTile] = r[m];
        }
        __syncthreads();
        if constexpr (DL > 4) {
            ifft_step_ct<LG1, kLgColTile, true, 0, NT, 4, 4>(buf, twCol + inv_tw_offset(LG1, 4));
            __syncthreads();
        }
        // ---- last step: magnitude, combine, running maximum from registers
        const bool lastCode = dp == ncodes - 1;
#pragma unroll
        for (int it = 0; it < itersL; ++it) {
            const int u = tid + it * NT, j = u >> kLgColTile;      // column c again: NT is a multiple of kColTile
            if (j < rEnd) {
                const float2* base = buf + pidx(j) * kColTile + c;
                float2 r[NL];
#pragma unroll
                for (int m = 0; m < NL; ++m) r[m] = base[((m << DL) + ((m << DL) >> 4)) * kColTile];
#pragma unroll
                for (int m = 1; m < NL; ++m) r[m] = cmul(r[m], __ldg(T + (brevq<QL>(m) - 1) * HL + j));
                ifft_regs_const<QL>(r);
#pragma unroll
                for (int m = 0; m < NL; ++m) {
                    const int row = j + (m << DL);
                    if (row < rEnd) {
                        float a = sqrt_approx(r[m].x * r[m].x + r[m].y * r[m].y);
                        float* mg = mag + (it * NL + m) * NT + tid;
                        if (dp > 0) a = combine == 1 ? *mg * 0.52440442408507577f + a * 0.85146931829632011f   // sqrt(11/40), sqrt(29/40)
                                                     : *mg + a;
                        if (!lastCode) *mg = a;
                        else {
                            const int lag = row * P2 + col0 + c;
                            bool ok = lag < pl.N;
                            if (useRanges) ok = ok && ((lag >= lo0 && lag <= hi0) || (lag >= lo1 && lag <= hi1));
                            if (ok && (a > best || (a == best && lag < bl))) {
                                best = a;
                                bl = lag;
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();   // the next code's first step overwrites buf
    }
    __shared__ float sv[NT / 32];
    __shared__ int sl[NT / 32];
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, best, o);
        int ol = __shfl_xor_sync(0xffffffffu, bl, o);
        if (ov > best || (ov == best && ol < bl)) {
            best = ov;
            bl = ol;
        }
    }
    if ((tid & 31) == 0) {
        sv[tid >> 5] = best;
        sl[tid >> 5] = bl;
    }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < NT / 32; ++w)
            if (sv[w] > best || (sv[w] == best && sl[w] < bl)) {
                best = sv[w];
                bl = sl[w];
            }
        // the bin's peak over all column tiles: largest value, then smallest lag, as ONE unsigned maximum of
        // (value bits << 32 | 0x7fffffff - lag) - magnitudes are non-negative floats, whose bit patterns order like
        // the values (acq_unpack_peak undoes it on the host); the slot is zeroed before the launch
        if (best >= 0.f)
            atomicMax(reinterpret_cast<unsigned long long*>(peaks) + bin,
                      ((unsigned long long)__float_as_uint(best) << 32) | (unsigned)(0x7fffffff - bl));
    }
}

// Sampled code tables on the device: table[t][k] = code[t][idx[k] - 1]  (makeDataTable.m:49-66; idx is computed once on
// the host with the reference's own expression, it does not depend on the PRN).  grid = (blocks, tables)
__global__ void acq_sample_tables_kernel(const int32_t* idx, const int8_t* codes, int codeLen, int spc, int8_t* tables) {
    const int8_t* code = codes + (size_t)blockIdx.y * codeLen;
    int8_t* out = tables + (size_t)blockIdx.y * spc;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < spc; k += gridDim.x * blockDim.x) out[k] = code[__ldg(idx + k) - 1];
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
    for (int o = 16; o > 0; o >>= 1) {   // max value, then min lag (first index wins, acquisition.m:229-232)
        float ov = __shfl_xor_sync(0xffffffffu, best, o);
        int ol = __shfl_xor_sync(0xffffffffu, bl, o);
        if (ov > best || (ov == best && ol < bl)) {
            best = ov;
            bl = ol;
        }
    }
    __shared__ float sv[8];
    __shared__ int sl[8];
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
        binPeak[bin] = AcqPeak{best, bl};
    }
}

// ---- integer power sums for sigPower (acquisition.m:150) ------------------------------------
// sums = {sum I, sum (I^2 + Q^2), sum Q, -}: mean and var() of a real or complex record in exact integers
__global__ void acq_power_kernel(const int8_t* x, int n, long long* sums /*[4]*/, int iq) {
    long long s1 = 0, s2 = 0, s3 = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int v = iq ? x[2 * (size_t)i] : x[i], w = iq ? x[2 * (size_t)i + 1] : 0;
        s1 += v;
        s2 += v * v + w * w;
        s3 += w;
    }
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        s3 += __shfl_xor_sync(0xffffffffu, s3, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd((unsigned long long*)&sums[0], (unsigned long long)s1);
        atomicAdd((unsigned long long*)&sums[1], (unsigned long long)s2);
        if (iq) atomicAdd((unsigned long long*)&sums[2], (unsigned long long)s3);
    }
}

// the same for float samples (resampled records): sums = {sum I, sum (I^2 + Q^2), sum Q, -} as doubles
__global__ void acq_power_f32_kernel(const int8_t* x, int n, double* sums /*[4]*/, int fmt) {
    double s1 = 0, s2 = 0, s3 = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float2 v = acq_sample(x, (size_t)i, fmt);
        s1 += (double)v.x;
        s2 += (double)v.x * (double)v.x + (double)v.y * (double)v.y;
        s3 += (double)v.y;
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    s3 = warp_sum(s3);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&sums[0], s1);
        atomicAdd(&sums[1], s2);
        atomicAdd(&sums[2], s3);
    }
}

// ---- resampling pre-conditioner (acquisition.m:56-123) --------------------------------------------------------------
// longSignal = filtfilt(b, 1, longSignal); longSignal = longSignal(index), index = ceil((0:len-1)/fsNew*fsOld), index(1) = 1.
// For an FIR b the zero-phase result on the original samples is exactly the correlation of the record, extended by
// filtfilt's odd reflection (nfact = 3*(numel(b)-1) >= 2*(numel(b)-1) samples at both ends), with g = b (*) flip(b): the
// start-up states filtfilt adds only shape the reflected margins it strips again.  Only the samples the index vector
// picks are evaluated (float64 accumulation; the acquisition continues on float samples).  grid-stride, one output each.
constexpr int kResampleThreads = 256;
__global__ void __launch_bounds__(kResampleThreads) acq_resample_kernel(const int8_t* x, int iq, long long n, const double* g,
                                                                       int half /* numel(b) - 1 */, double fsNew, double fsOld,
                                                                       long long outLen, float* out) {
    extern __shared__ double gs[];
    for (int i = threadIdx.x; i < 2 * half + 1; i += blockDim.x) gs[i] = g[i];
    __syncthreads();
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < outLen; k += (long long)gridDim.x * blockDim.x) {
        long long idx = (long long)ceil(__dmul_rn(__ddiv_rn((double)k, fsNew), fsOld)) - 1;   // 0-based sample
        if (k == 0) idx = 0;
        double ar = 0.0, ai = 0.0;
        if (idx - half >= 0 && idx + half < n) {   // interior: plain correlation
            if (iq) {
                const char2* p = reinterpret_cast<const char2*>(x) + (idx - half);
                for (int j = 0; j <= 2 * half; ++j) {
                    const char2 v = p[j];
                    ar = fma(gs[j], (double)v.x, ar);
                    ai = fma(gs[j], (double)v.y, ai);
                }
            } else {
                const int8_t* p = x + (idx - half);
                for (int j = 0; j <= 2 * half; ++j) ar = fma(gs[j], (double)p[j], ar);
            }
        } else {                                   // margins: odd reflection about the first / last sample
            for (int j = 0; j <= 2 * half; ++j) {
                long long m = idx - half + j;
                double sr, si = 0.0;
                if (m < 0) {
                    const long long r = -m;
                    sr = 2.0 * (double)x[0] - (double)(iq ? x[2 * r] : x[r]);
                    if (iq) si = 2.0 * (double)x[1] - (double)x[2 * r + 1];
                } else if (m >= n) {
                    const long long r = 2 * (n - 1) - m;
                    sr = 2.0 * (double)(iq ? x[2 * (n - 1)] : x[n - 1]) - (double)(iq ? x[2 * r] : x[r]);
                    if (iq) si = 2.0 * (double)x[2 * (n - 1) + 1] - (double)x[2 * r + 1];
                } else {
                    sr = (double)(iq ? x[2 * m] : x[m]);
                    if (iq) si = (double)x[2 * m + 1];
                }
                ar = fma(gs[j], sr, ar);
                ai = fma(gs[j], si, ai);
            }
        }
        if (iq) reinterpret_cast<float2*>(out)[k] = make_float2((float)ar, (float)ai);
        else out[k] = (float)ar;
    }
}

// ---- fine search ---------------------------------------------------------------------------
// B1C (acquisition.m:253-300): for fine bin j, component dp:
//   A = sum_n x[n] T[n] e^{i phi_j n},  B = sum_n T[n] e^{i phi_j n}   (DC removal applied on host: A - mean*B)
// out[(j*ncodes+dp)*4 + {0..3}] = {Re A, Im A, Re B, Im B}; grid = (slices, nfine, ncodes)
__global__ void acq_fine_b1c_kernel(const int8_t* x, const int8_t* tables, int spc, const unsigned long long* dphi,
                                    int ncodes, double* out, int iq) {
    const int j = blockIdx.y, dp = blockIdx.z;
    const int8_t* T = tables + (size_t)dp * spc;
    const unsigned long long d = dphi[j];
    float ar = 0, ai = 0, br = 0, bi = 0;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < spc; n += gridDim.x * blockDim.x) {
        unsigned long long ph = (unsigned long long)n * d;
        float sn, cs;
        sincospif((float)(int)(ph >> 32) * 4.656612873077392578125e-10f, &sn, &cs);
        const float t = (float)T[n];
        const float2 xv = acq_sample(x, n, iq);
        const float xs = xv.x * t, xq = xv.y * t;
        ar += xs * cs - xq * sn;
        ai += xs * sn + xq * cs;
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
                                    double tc, const unsigned long long* dphi, int nseg, double* out, int iq) {
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
        const float2 xv = acq_sample(x, (size_t)n0, iq);
        const float mr = xv.x * cs - xv.y * sn, mi = xv.x * sn + xv.y * cs;
        dr += cd * mr;
        di += cd * mi;
        pr += cp * mr;
        pi += cp * mi;
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

// Work buffers come from a small process-wide pool: cudaMalloc / cudaFree of the 0.7 - 10 GB an acquisition needs
// costs tens of milliseconds (and varies wildly), more than the search itself.  The pool is emptied by bds_shutdown.
struct PoolBlock {
    void* p;
    size_t cap;
};
std::mutex g_poolMu;
std::vector<PoolBlock> g_pool;
// streams of the inverse passes (acquire_core), made once per device, destroyed by bds_shutdown
std::mutex g_streamsMu;
cudaStream_t g_streams[4] = {nullptr, nullptr, nullptr, nullptr};
int g_streamsDev = -1;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    ~DevBuf() { release(); }
    template <typename T>
    T* as() { return reinterpret_cast<T*>(p); }
    void release() {
        if (!p) return;
        std::lock_guard<std::mutex> lock(g_poolMu);
        g_pool.push_back(PoolBlock{p, cap});
        p = nullptr;
        cap = 0;
    }
    cudaError_t alloc(size_t bytes) {
        release();
        if (bytes == 0) bytes = 16;
        {
            std::lock_guard<std::mutex> lock(g_poolMu);
            int best = -1;
            for (int i = 0; i < (int)g_pool.size(); ++i)   // smallest block that fits, but not one wastefully larger
                if (g_pool[i].cap >= bytes && g_pool[i].cap <= 2 * bytes + (1 << 20) && (best < 0 || g_pool[i].cap < g_pool[best].cap))
                    best = i;
            if (best >= 0) {
                p = g_pool[best].p;
                cap = g_pool[best].cap;
                g_pool.erase(g_pool.begin() + best);
                return cudaSuccess;
            }
        }
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) {   // out of memory: give the pooled blocks back and retry once
            cudaGetLastError();
            bds::acq_pool_release();
            e = cudaMalloc(&p, bytes);
        }
        cap = e == cudaSuccess ? bytes : 0;
        if (e != cudaSuccess) p = nullptr;
        return e;
    }
};

// ---- compile-time specialised inverse passes: one instantiation per transform shape of the shipped configurations
typedef void (*InvRowFn)(AcqPlan, const float2*, const float2*, float2*, int, const int*, const float2*);
typedef void (*InvColFn)(AcqPlan, const float2*, int, int, int, int, int, int, int, AcqPeak*, const float2*);
struct InvCt {
    InvRowFn row = nullptr;
    InvColFn col = nullptr;
    int thrRow = 0, thrCol = 0, lgRT = 0;
};
template <int LG1, int LG2, int LGRT>
InvCt make_inv_ct() {
    InvCt f;
    f.row = acq_inv_row_ct_kernel<LG1, LG2, LGRT>;
    f.col = acq_inv_col_ct_kernel<LG1, LG2>;
    f.thrRow = inv_row_threads(LG2, LGRT);
    f.thrCol = inv_col_threads(LG1);
    f.lgRT = LGRT;
    return f;
}
// (log2 P1, log2 P2) as acquire_core plans them, and the rows per CTA of the specialised row pass:
//   (10, 12)  B1C at 99.375 MHz (BASELINE), P = 2^22      (9, 12)  B1C at the reference's shipped 53 MHz, P = 2^21
//   (9, 10)   B2a at 99.375 MHz, P = 2^19
// every other shape (e.g. the band-pass rates of the resampling branch) runs on the generic kernels, and so does every
// shape when bit 0 of cfg.tune is set (tests compare the two implementations that way)
InvCt find_inv_ct(const AcqPlan& pl, int tune) {
    if (tune & 1) return InvCt{};
    if (pl.log2P1 == 10 && pl.log2P2 == 12) return (tune & 2) ? make_inv_ct<10, 12, 0>() : make_inv_ct<10, 12, 1>();
    if (pl.log2P1 == 9 && pl.log2P2 == 12) return (tune & 2) ? make_inv_ct<9, 12, 0>() : make_inv_ct<9, 12, 1>();
    if (pl.log2P1 == 9 && pl.log2P2 == 10) return (tune & 2) ? make_inv_ct<9, 10, 1>() : make_inv_ct<9, 10, 2>();
    return InvCt{};
}
// step twiddles of a 2^LG-point inverse, see inv_tw_offset
std::vector<float2> inv_step_twiddles(int LG) {
    std::vector<float2> t;
    const double twoPi = 6.283185307179586476925286766559;
    for (int D = 4; D < LG; D += 4) {
        const int Q = std::min(4, LG - D), h = 1 << D;
        const double L = (double)(1 << (D + Q));
        for (int mp = 1; mp < (1 << Q); ++mp)
            for (int j = 0; j < h; ++j) {
                const double a = twoPi * (double)j * (double)mp / L;
                t.push_back(make_float2((float)std::cos(a), (float)std::sin(a)));
            }
    }
    if (t.empty()) t.push_back(make_float2(1.f, 0.f));
    return t;
}

// peaks of the specialised column pass arrive packed (value bits << 32 | 0x7fffffff - lag), see acq_inv_col_ct_kernel
void acq_unpack_peaks(std::vector<AcqPeak>& v) {
    for (auto& p : v) {
        unsigned long long u;
        std::memcpy(&u, &p, 8);
        const unsigned hi = (unsigned)(u >> 32), lo = (unsigned)u;
        float val;
        std::memcpy(&val, &hi, 4);
        p = u == 0 ? AcqPeak{-1.f, 0x7fffffff} : AcqPeak{val, (int)(0x7fffffffu - lo)};
    }
}

unsigned long long freq_to_dphi(double f, double fs) {
    double r = f / fs;
    r -= std::floor(r);
    return (unsigned long long)(r * 18446744073709551616.0);
}

long mround(double x) { return std::lround(x); }

// The ranging codes are constants of the signal: generate each (component, PRN) once per process.
const std::vector<int8_t>* cached_component(int component, int prn) {
    static std::mutex mu;
    static std::map<std::pair<int, int>, std::vector<int8_t>> cache;
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find({component, prn});
    if (it != cache.end()) return &it->second;
    std::vector<int8_t> v;
    if (!gen_component(component, prn, v)) return nullptr;
    return &cache.emplace(std::make_pair(component, prn), std::move(v)).first->second;
}

// The search proper (acquisition.m:129-338 / B2a acquisition.m:130-365) on a DEVICE record dx of n samples in format fmt
// (kFmt*: int8 or float, real or I/Q).  Outputs are zero-filled by the caller.
int acquire_core(int signal, const int8_t* dx, int fmt, size_t n, const bds_acq_cfg* cfg, const int32_t* prn, int n_prn,
                 int prn_lo, int prn_hi, double* carrFreq, double* codePhase, double* peakMetric, int max_prn, double* dbg) {
    const bool b1c = signal == BDS_SIG_B1C;
    const int iq = fmt;   // the kernels' sample-format argument
    const size_t sampleBytes = acq_sample_bytes(fmt);
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
    // long transforms: shorter columns (more column tiles per launch, two CTAs per SM) and 4096-point rows
    pl.log2P2 = lgP >= 21 ? std::min(12, lgP - lgP / 2 + 1) : lgP - lgP / 2;
    pl.log2P1 = lgP - pl.log2P2;
    pl.lgRowTile = pl.log2P2 >= 12 ? 1 : 2;
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
    const size_t smemCol = sizeof(float2) * (pl.P1 + (pl.P1 >> 4)) * kColTile;   // padded, see pidx()
    const size_t smemColInv = smemCol + sizeof(float) * pl.P1 * kColTile;
    const int rowTile = 1 << pl.lgRowTile;
    const size_t smemRow = sizeof(float2) * (pl.P2 + (pl.P2 >> 4)) * rowTile;   // padded rows, see pidx()
    TRYA(cudaFuncSetAttribute(acq_fwd_col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemCol));
    TRYA(cudaFuncSetAttribute(acq_inv_col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemColInv));
    TRYA(cudaFuncSetAttribute(acq_fwd_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemRow));
    TRYA(cudaFuncSetAttribute(acq_inv_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemRow));
    const int colGroups = pl.P2 / kColTile;
    // one thread per 16-element register group (fewer for the short B2a transforms)
    const int thrRow = std::min(kAcqThreads, std::max(128, (pl.P2 >> 4) * rowTile));
    const int thrCol = std::min(kAcqThreads, std::max(128, (pl.P1 >> 4) * kColTile));
    const InvCt ct = find_inv_ct(pl, cfg->tune);
    // blocking streams (ordered against the legacy default stream the rest of the call uses), made once per device
    cudaStream_t streams[4] = {nullptr, nullptr, nullptr, nullptr};
    {
        std::lock_guard<std::mutex> lock(g_streamsMu);
        int dev = -1;
        TRYA(cudaGetDevice(&dev));
        if (dev != g_streamsDev) {    // first call, or the caller moved to another device (bds_init)
            for (auto& st : g_streams) st = nullptr;   // streams of another device stay with that device's context
            for (auto& st : g_streams)
                if (cudaStreamCreate(&st) != cudaSuccess) st = nullptr;
            g_streamsDev = dev;
        }
        for (int k = 0; k < 4; ++k) streams[k] = g_streams[k];
    }
    const bool haveStreams = streams[0] && streams[1] && streams[2] && streams[3];
    const int nStreams = !haveStreams || (cfg->tune & 4) ? 1 : ((cfg->tune & 8) ? 4 : 2);
    size_t workElems = 0, peakElems = 0;
    const size_t smemRowCt = sizeof(float2) * ((size_t)(pl.P2 + (pl.P2 >> 4)) << ct.lgRT);
    DevBuf dTwRowS, dTwColS;
    if (ct.row) {
        const std::vector<float2> tr = inv_step_twiddles(pl.log2P2), tc = inv_step_twiddles(pl.log2P1);
        TRYA(dTwRowS.alloc(tr.size() * 8));
        TRYA(dTwColS.alloc(tc.size() * 8));
        TRYA(cudaMemcpy(dTwRowS.p, tr.data(), tr.size() * 8, cudaMemcpyHostToDevice));
        TRYA(cudaMemcpy(dTwColS.p, tc.data(), tc.size() * 8, cudaMemcpyHostToDevice));
        TRYA(cudaFuncSetAttribute((const void*)ct.row, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemRowCt));
        TRYA(cudaFuncSetAttribute((const void*)ct.col, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemColInv));
    }
    // the inverse passes of nb bins against the ncodes spectra of one PRN; per-bin peaks -> out[nb].  Consecutive calls
    // alternate between two streams with a work buffer each: the row pass of one cell fills the SMs the tail of the
    // other cell's column pass leaves idle (1024 / 512 CTAs per launch are 3.5 / 1.7 waves of 296 resident CTAs)
    int slot = 0;
    auto inverse_passes = [&](const float2* sigBins, int nb, const float2* code, const int* binMap, int lo0, int hi0, int lo1,
                              int hi1, int useRanges, AcqPeak* out) {
        cudaStream_t st = nStreams > 1 ? streams[slot] : (cudaStream_t)0;
        float2* work = dWork.as<float2>() + (size_t)slot * workElems;
        AcqPeak* pk = dPeaks.as<AcqPeak>() + (size_t)slot * peakElems;
        slot = (slot + 1) % nStreams;
        if (ct.row) {
            ct.row<<<dim3((pl.P1 >> ct.lgRT) * ncodes, nb), ct.thrRow, smemRowCt, st>>>(pl, sigBins, code, work, ncodes, binMap,
                                                                                     dTwRowS.as<float2>());
            ct.col<<<dim3(colGroups, nb), ct.thrCol, smemColInv, st>>>(pl, work, ncodes, combine, lo0, hi0, lo1, hi1, useRanges, out,
                                                                       dTwColS.as<float2>());
            count_launch(2);
            return;
        } else {
            acq_inv_row_kernel<<<dim3(pl.P1 / rowTile, nb, ncodes), thrRow, smemRow, st>>>(pl, sigBins, code, work, ncodes, binMap);
            acq_inv_col_kernel<<<dim3(colGroups, nb), thrCol, smemColInv, st>>>(pl, work, ncodes, combine, lo0, hi0, lo1, hi1,
                                                                                useRanges, pk);
        }
        acq_peak_reduce_kernel<<<nb, 256, 0, st>>>(pk, colGroups, out);
        count_launch(3);
    };
    acq_fwd_col_kernel<<<dim3(colGroups, nbins), thrCol, smemCol>>>(pl, 0, dx, 0, dDphi.as<unsigned long long>(),
                                                                       dSig.as<float2>(), iq);
    acq_fwd_row_kernel<<<dim3(pl.P1 / rowTile, nbins), thrRow, smemRow>>>(pl, dSig.as<float2>(), 0.f, ct.row ? 1 : 0);
    count_launch(2);
    TRYA(cudaGetLastError());

    // ---- signal power (B1C metric normaliser)  acquisition.m:150
    double sigPower = 0;
    if (b1c) {
        TRYA(dSums.alloc(32));
        TRYA(cudaMemset(dSums.p, 0, 32));
        double ps[4];   // sum I, sum |x|^2, sum Q
        if (fmt & 2) {
            acq_power_f32_kernel<<<g_num_sms * 4, 256>>>(dx, (int)M, dSums.as<double>(), fmt);
            TRYA(cudaMemcpy(ps, dSums.p, 32, cudaMemcpyDeviceToHost));
        } else {
            acq_power_kernel<<<g_num_sms * 4, 256>>>(dx, (int)M, dSums.as<long long>(), iq);
            long long hs[4];
            TRYA(cudaMemcpy(hs, dSums.p, 32, cudaMemcpyDeviceToHost));
            for (int k = 0; k < 4; ++k) ps[k] = (double)hs[k];
        }
        count_launch();
        // var() of a real or complex vector: (sum |x|^2 - |sum x|^2 / M) / (M - 1)
        double var = (ps[1] - (ps[0] * ps[0] + ps[2] * ps[2]) / (double)M) / (double)(M - 1);
        sigPower = std::sqrt(var * (double)M);
    }

    // ---- the PRN list is processed in chunks whose code spectra fit comfortably in HBM; inside a chunk every
    //      phase queues the kernels of all its PRNs back to back and synchronises once (no host round trip per PRN)
    const int nSelAll = std::max(0, prn_hi - prn_lo);
    size_t freeB = 0, totalB = 0;
    TRYA(cudaMemGetInfo(&freeB, &totalB));
    const size_t perPrn = specBytes * ncodes + (size_t)spc * ncodes;
    const int chunkMax = (int)std::max<size_t>(1, std::min<size_t>(64, (freeB / 3) / perPrn));
    // bins per launch: the row pass's output of a launch (two launches are in flight, one per stream) has to stay in L2
    // until the column pass has read it back.  Measured on the B2a grid (4 MB transforms): 4 bins 16.1 ms, 8 bins 16.6 ms,
    // 16 bins 18.4 ms, 26 bins 19.2 ms; B1C (32 MB transforms, 64 MB per bin): 1 bin 129 ms, 2 bins 132 ms
    const int binsPerBatch = (int)std::max<size_t>(1, std::min<size_t>({(size_t)nbins, (size_t)8, ((size_t)32 << 20) / (specBytes * ncodes)}));
    workElems = (size_t)pl.P * ncodes * binsPerBatch;
    peakElems = (size_t)colGroups * binsPerBatch;
    TRYA(dWork.alloc(sizeof(float2) * workElems * nStreams));
    TRYA(dPeaks.alloc(sizeof(AcqPeak) * peakElems * nStreams));
    DevBuf dBinMap, dSecond, dIdx;
    {   // makeDataTable.m:49-63 / makeB2aDataTable.m:46-62: idx = ceil((ts*k)/tc), k = 1..spc; idx(end) forced to the last
        // element; B1C also forces idx(1) = 1  (same expression as bds_make_code_table)
        std::vector<int32_t> idx(spc);
        const double ts = 1.0 / fs;
        const double tc = b1c ? (1.0 / cfg->codeFreqBasis) / 2.0 : 1.0 / cfg->codeFreqBasis;
        const long last = b1c ? 2L * cfg->codeLength : (long)cfg->codeLength;
        for (long k = 1; k <= spc; ++k) {
            long v = (long)std::ceil((ts * (double)k) / tc);
            if (k == spc) v = last;
            if (b1c && k == 1) v = 1;
            if (v < 1 || v > last) return set_error(BDS_ERR_ARG, "code index out of range (fs/codeFreqBasis mismatch)");
            idx[k - 1] = (int32_t)v;
        }
        TRYA(dIdx.alloc(sizeof(int32_t) * spc));
        TRYA(cudaMemcpy(dIdx.p, idx.data(), sizeof(int32_t) * spc, cudaMemcpyHostToDevice));
    }
    struct Cand {
        int PRN, bestBin;
        long cp;
        float peak;
        double metric, norm;
    };
    for (int c0 = 0; c0 < nSelAll; c0 += chunkMax) {
        const int nSel = std::min(chunkMax, nSelAll - c0);
        // ---- phase 0: sampled code tables of the chunk (host, integer) and their conjugated, 1/P-scaled spectra
        const int compD = b1c ? BDS_CODE_B1C_DATA_BOC11 : BDS_CODE_B2A_DATA, compP = b1c ? BDS_CODE_B1C_PILOT_BOC11 : BDS_CODE_B2A_PILOT;
        const int codeLen = component_length(compD);
        std::vector<int8_t> codes((size_t)nSel * ncodes * codeLen);
        for (int i = 0; i < nSel; ++i) {
            const int PRN = prn[prn_lo + c0 + i];
            for (int dp = 0; dp < ncodes; ++dp) {
                const std::vector<int8_t>* cc = cached_component(dp == 0 ? compD : compP, PRN);
                if (!cc) return set_error(BDS_ERR_ARG, "PRN %d out of range 1..63", PRN);
                std::memcpy(&codes[((size_t)i * ncodes + dp) * codeLen], cc->data(), codeLen);
            }
        }
        DevBuf dCodes;
        TRYA(dCodes.alloc(codes.size()));
        TRYA(cudaMemcpy(dCodes.p, codes.data(), codes.size(), cudaMemcpyHostToDevice));
        TRYA(dTab.alloc((size_t)nSel * ncodes * spc));
        acq_sample_tables_kernel<<<dim3(std::max(1L, std::min(64L, spc / 4096)), nSel * ncodes), 256>>>(
            dIdx.as<int32_t>(), dCodes.as<int8_t>(), codeLen, (int)spc, dTab.as<int8_t>());
        count_launch();
        TRYA(dCode.alloc(specBytes * ncodes * nSel));
        acq_fwd_col_kernel<<<dim3(colGroups, ncodes * nSel), thrCol, smemCol>>>(pl, 1, dTab.as<int8_t>(), (size_t)spc, nullptr,
                                                                                   dCode.as<float2>(), 0);
        acq_fwd_row_kernel<<<dim3(pl.P1 / rowTile, ncodes * nSel), thrRow, smemRow>>>(pl, dCode.as<float2>(), 1.0f / (float)pl.P,
                                                                                      ct.row ? 1 : 0);
        count_launch(2);
        // ---- phase 1: coarse PRN x Doppler grid; per (PRN, bin) peak and first lag
        TRYA(dBinPeak.alloc(sizeof(AcqPeak) * (size_t)nbins * nSel));
        if (ct.row) TRYA(cudaMemset(dBinPeak.p, 0, sizeof(AcqPeak) * (size_t)nbins * nSel));
        for (int i = 0; i < nSel; ++i) {
            const float2* code = dCode.as<float2>() + (size_t)i * ncodes * pl.P;
            for (int b0 = 0; b0 < nbins; b0 += binsPerBatch) {
                const int nb = std::min(binsPerBatch, nbins - b0);
                inverse_passes(dSig.as<float2>() + (size_t)b0 * pl.P, nb, code, nullptr, 0, 0, 0, 0, 0,
                               dBinPeak.as<AcqPeak>() + (size_t)i * nbins + b0);
            }
        }
        TRYA(cudaGetLastError());
        std::vector<AcqPeak> binPeak((size_t)nbins * nSel);
        TRYA(cudaMemcpy(binPeak.data(), dBinPeak.p, sizeof(AcqPeak) * binPeak.size(), cudaMemcpyDeviceToHost));
        if (ct.row) acq_unpack_peaks(binPeak);
        std::vector<Cand> cand(nSel);
        for (int i = 0; i < nSel; ++i) {
            // [~, bin] = max(max(results,[],2)); [peak, codePhase] = max(max(results))  (first index wins)
            const AcqPeak* bp = &binPeak[(size_t)i * nbins];
            float peak = -1.f;
            int bestBin = 0;
            for (int k = 0; k < nbins; ++k)
                if (bp[k].val > peak) {
                    peak = bp[k].val;
                    bestBin = k;
                }
            int lag = 0x7fffffff;
            for (int k = 0; k < nbins; ++k)
                if (bp[k].val == peak) lag = std::min(lag, bp[k].lag);
            cand[i] = Cand{prn[prn_lo + c0 + i], bestBin, (long)lag + 1 /* 1-based */, peak, 0.0, 0.0};
        }
        // ---- phase 2: metric normaliser
        if (b1c) {
            for (auto& cd : cand) {
                cd.norm = sigPower;
                cd.metric = (double)cd.peak / sigPower;                       // acquisition.m:235
            }
        } else {
            // second peak in the best bin's row, excluding +-samples2CodeChip around the peak and
            // its one-period image   (B2a acquisition.m:224-252)
            TRYA(dSecond.alloc(sizeof(AcqPeak) * nSel));
            if (ct.row) TRYA(cudaMemset(dSecond.p, 0, sizeof(AcqPeak) * nSel));
            TRYA(dBinMap.alloc(sizeof(int) * nSel));
            std::vector<int> bm(nSel);
            for (int i = 0; i < nSel; ++i) bm[i] = cand[i].bestBin;
            TRYA(cudaMemcpy(dBinMap.p, bm.data(), sizeof(int) * nSel, cudaMemcpyHostToDevice));
            const long s2cc = (long)std::ceil(fs / cfg->codeFreqBasis) * 2;
            for (int i = 0; i < nSel; ++i) {
                const long cp = cand[i].cp;
                const long e1 = cp - s2cc, e2 = cp + s2cc, e3 = cp - spc + s2cc, e4 = cp + spc - s2cc;
                int lo0 = 1, hi0 = 0, lo1 = 1, hi1 = 0;  // empty
                if (e1 >= 1) {
                    lo0 = (int)std::max(1L, e3) - 1;
                    hi0 = (int)e1 - 1;
                }
                if (e2 < N) {
                    lo1 = (int)e2 - 1;
                    hi1 = (int)std::min(e4, N) - 1;
                }
                inverse_passes(dSig.as<float2>(), 1, dCode.as<float2>() + (size_t)i * ncodes * pl.P, dBinMap.as<int>() + i, lo0, hi0,
                               lo1, hi1, 1, dSecond.as<AcqPeak>() + i);
            }
            std::vector<AcqPeak> second(nSel);
            TRYA(cudaMemcpy(second.data(), dSecond.p, sizeof(AcqPeak) * nSel, cudaMemcpyDeviceToHost));
            if (ct.row) acq_unpack_peaks(second);
            for (int i = 0; i < nSel; ++i) {
                cand[i].norm = second[i].val;
                cand[i].metric = (double)cand[i].peak / (double)second[i].val;
            }
        }
        // ---- phase 3: threshold, fine frequency search of the detected PRNs (queued together, one sync)
        struct FineJob {
            int i, nfine;
            size_t off;        // offset into dFine (doubles)
            long cp;
            std::vector<double> ff;
        };
        std::vector<FineJob> jobs;
        size_t fineDoubles = 0, fineDphis = 0;
        const int nseg = cfg->fineNoncoh;
        for (int i = 0; i < nSel; ++i) {
            Cand& cd = cand[i];
            const int PRN = cd.PRN;
            peakMetric[PRN - 1] = cd.metric;
            if (b1c && cd.cp + spc - 1 > (long)n) cd.cp -= spc;               // acquisition.m:239-241
            if (dbg) {
                dbg[(PRN - 1) * 4 + 0] = cd.bestBin;
                dbg[(PRN - 1) * 4 + 1] = (double)cd.cp;
                dbg[(PRN - 1) * 4 + 2] = cd.peak;
                dbg[(PRN - 1) * 4 + 3] = cd.norm;
            }
            if (!(cd.metric > cfg->acqThreshold)) continue;
            if (cd.cp < 1) continue;  // the reference would index longSignal(<=0) here and abort
            FineJob j;
            j.i = i;
            j.cp = cd.cp;
            if (b1c) {
                if ((size_t)(cd.cp - 1 + spc) > n) continue;
                j.nfine = (int)mround(cfg->acqStep / 25) * 2 + 1;             // acquisition.m:266-267
                for (int q = 0; q < j.nfine; ++q) j.ff.push_back(frq[cd.bestBin] - cfg->acqStep + 25.0 * q);
                j.off = fineDoubles;
                fineDoubles += (size_t)j.nfine * ncodes * 4 + 4;              // + 4 slots for the power sums (as long long)
            } else {
                if (nseg <= 0 || (size_t)(cd.cp - 1 + (long)nseg * spc) > n) continue;
                j.nfine = (int)mround(cfg->acqStep / 25) + 1;                 // B2a acquisition.m:265
                for (int q = 0; q < j.nfine; ++q) j.ff.push_back(frq[cd.bestBin] - cfg->acqStep / 2 + 25.0 * q);
                j.off = fineDoubles;
                fineDoubles += (size_t)j.nfine * 2 * nseg * 2;
            }
            fineDphis += j.nfine;
            jobs.push_back(std::move(j));
        }
        if (!jobs.empty()) {
            std::vector<unsigned long long> fd;
            fd.reserve(fineDphis);
            for (auto& j : jobs)
                for (double f : j.ff) fd.push_back(freq_to_dphi(f, fs));
            TRYA(dFineDphi.alloc(sizeof(unsigned long long) * fd.size()));
            TRYA(cudaMemcpy(dFineDphi.p, fd.data(), sizeof(unsigned long long) * fd.size(), cudaMemcpyHostToDevice));
            TRYA(dFine.alloc(sizeof(double) * fineDoubles));
            TRYA(cudaMemset(dFine.p, 0, sizeof(double) * fineDoubles));
            std::vector<uint32_t> bitsAll;
            if (!b1c) {
                bitsAll.resize((size_t)jobs.size() * 2 * kPackedWords);
                std::vector<uint8_t> chips;
                for (size_t q = 0; q < jobs.size(); ++q) {
                    const int PRN = cand[jobs[q].i].PRN;
                    primary_bits(BDS_CODE_B2A_DATA, PRN, chips);
                    pack_bits(chips, &bitsAll[(q * 2 + 0) * kPackedWords]);
                    primary_bits(BDS_CODE_B2A_PILOT, PRN, chips);
                    pack_bits(chips, &bitsAll[(q * 2 + 1) * kPackedWords]);
                }
                TRYA(dBits.alloc(bitsAll.size() * 4));
                TRYA(cudaMemcpy(dBits.p, bitsAll.data(), bitsAll.size() * 4, cudaMemcpyHostToDevice));
            }
            size_t dphiOff = 0;
            for (size_t q = 0; q < jobs.size(); ++q) {
                const FineJob& j = jobs[q];
                double* out = dFine.as<double>() + j.off;
                const unsigned long long* dph = dFineDphi.as<unsigned long long>() + dphiOff;
                if (b1c) {   // acquisition.m:253-307
                    double* pw = out + (size_t)j.nfine * ncodes * 4;   // power sums: long long (int8 records) or double
                    const int8_t* xs = dx + (size_t)(j.cp - 1) * sampleBytes;
                    if (fmt & 2) acq_power_f32_kernel<<<g_num_sms * 4, 256>>>(xs, (int)spc, pw, fmt);
                    else acq_power_kernel<<<g_num_sms * 4, 256>>>(xs, (int)spc, reinterpret_cast<long long*>(pw), iq);
                    acq_fine_b1c_kernel<<<dim3(g_num_sms, j.nfine, ncodes), 256>>>(
                        xs, dTab.as<int8_t>() + (size_t)j.i * ncodes * spc, (int)spc, dph, ncodes, out, iq);
                    count_launch(2);
                } else {     // B2a acquisition.m:256-335
                    acq_fine_b2a_kernel<<<dim3(32, j.nfine, nseg), 256>>>(dx + (size_t)(j.cp - 1) * sampleBytes, dBits.as<uint32_t>() + q * 2 * kPackedWords,
                                                                         (int)spc, 1.0 / fs, 1.0 / cfg->codeFreqBasis, dph, nseg, out, iq);
                    count_launch();
                }
                dphiOff += j.nfine;
            }
            TRYA(cudaGetLastError());
            std::vector<double> fo(fineDoubles);
            TRYA(cudaMemcpy(fo.data(), dFine.p, sizeof(double) * fineDoubles, cudaMemcpyDeviceToHost));
            for (const FineJob& j : jobs) {
                const int PRN = cand[j.i].PRN;
                const double* o0 = &fo[j.off];
                int best = 0;
                double bestV = -1;
                if (b1c) {
                    double ps[4];
                    if (fmt & 2) std::memcpy(ps, o0 + (size_t)j.nfine * ncodes * 4, 32);
                    else {
                        long long hs[4];
                        std::memcpy(hs, o0 + (size_t)j.nfine * ncodes * 4, 32);
                        for (int k = 0; k < 4; ++k) ps[k] = (double)hs[k];
                    }
                    const double mr = ps[0] / (double)spc, mi = ps[2] / (double)spc;   // mean(signal), complex for I/Q
                    for (int q = 0; q < j.nfine; ++q) {
                        double v[2] = {0, 0};
                        for (int dp = 0; dp < ncodes; ++dp) {
                            const double* o = o0 + ((size_t)q * ncodes + dp) * 4;   // sum (x - mean) T e = A - mean * B
                            v[dp] = std::hypot(o[0] - (mr * o[2] - mi * o[3]), o[1] - (mr * o[3] + mi * o[2]));
                        }
                        const double r = ncodes == 2 ? (v[0] * 11 + v[1] * 29) / 40 : v[0];
                        if (r > bestV) {
                            bestV = r;
                            best = q;
                        }
                    }
                } else {
                    for (int q = 0; q < j.nfine; ++q) {
                        double r = 0;
                        for (int dp = 0; dp < 2; ++dp)
                            for (int sg = 0; sg < nseg; ++sg) {
                                const double* o = o0 + (((size_t)q * 2 + dp) * nseg + sg) * 2;
                                r += std::hypot(o[0], o[1]);
                            }
                        if (r > bestV) {
                            bestV = r;
                            best = q;
                        }
                    }
                }
                carrFreq[PRN - 1] = j.ff[best];
                if (carrFreq[PRN - 1] == 0) carrFreq[PRN - 1] = 1;          // acquisition.m:303-305
                codePhase[PRN - 1] = (double)j.cp;
            }
        }
    }
    TRYA(cudaDeviceSynchronize());
    return BDS_OK;
}

// b = fir1(order, [w1 w2]) (band-pass, Hamming window, scaled to unit gain at the centre of the pass band) and
// g = b (*) flip(b), the zero-phase kernel filtfilt(b, 1, .) applies (acquisition.m:70-74)
void fir1_bandpass_autocorr(int order, double w1, double w2, std::vector<double>& g) {
    const int L = order + 1;
    const double pi = 3.14159265358979323846;
    std::vector<double> b(L);
    auto sinc = [&](double x) { return x == 0.0 ? 1.0 : std::sin(pi * x) / (pi * x); };
    for (int i = 0; i < L; ++i) {
        const double m = (double)i - 0.5 * order;
        const double ideal = w2 * sinc(w2 * m) - w1 * sinc(w1 * m);           // least-squares fit of the ideal band
        const double win = 0.54 - 0.46 * std::cos(2.0 * pi * (double)i / (double)order);   // hamming(L)
        b[i] = ideal * win;
    }
    const double f0 = 0.5 * (w1 + w2);   // b = b / abs(exp(-j*2*pi*(0:L-1)*(f0/2)) * b.')
    double re = 0, im = 0;
    for (int i = 0; i < L; ++i) {
        re += b[i] * std::cos(pi * f0 * i);
        im -= b[i] * std::sin(pi * f0 * i);
    }
    const double sc = std::hypot(re, im);
    for (auto& v : b) v /= sc;
    g.assign(2 * order + 1, 0.0);
    for (int k = -order; k <= order; ++k) {
        double a = 0;
        for (int i = std::max(0, -k); i < L && i + k < L; ++i) a += b[i] * b[i + k];
        g[k + order] = a;
    }
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
    if (cfg->fileType != 0 && cfg->fileType != 1 && cfg->fileType != 2)
        return set_error(BDS_ERR_ARG, "fileType must be 1 (real) or 2 (I/Q), got %d", cfg->fileType);
    int rc = require_device();
    if (rc) return rc;
    std::memset(carrFreq, 0, sizeof(double) * max_prn);
    std::memset(codePhase, 0, sizeof(double) * max_prn);
    std::memset(peakMetric, 0, sizeof(double) * max_prn);
    if (dbg) std::memset(dbg, 0, sizeof(double) * max_prn * 4);
    prn_lo = std::max(prn_lo, 0);
    prn_hi = std::min(prn_hi, n_prn);
    const int iq = cfg->fileType == 2;   // x holds n I/Q pairs = 2n bytes

    // ---- IF record on the device
    DevBuf dX, dRes, dG;
    const int8_t* dx = x;
    if (x_loc == BDS_LOC_HOST) {
        TRYA(dX.alloc(n << iq));
        TRYA(cudaMemcpy(dX.p, x, n << iq, cudaMemcpyHostToDevice));
        dx = dX.as<int8_t>();
    }
    if (!(cfg->resamplingflag == 1 && cfg->samplingFreq > cfg->resamplingThreshold))
        return acquire_core(signal, dx, iq, n, cfg, prn, n_prn, prn_lo, prn_hi, carrFreq, codePhase, peakMetric, max_prn, dbg);

    // ---- resampling pre-conditioner (acquisition.m:56-123 / B2a acquisition.m:56-124): zero-phase band-pass around the IF,
    //      then the search runs at a band-pass sampling rate on every index-th sample
    const double fsOld = cfg->samplingFreq, IF = cfg->IF;
    const double BW = signal == BDS_SIG_B1C ? 9e6 : cfg->codeFreqBasis * 2 + 0.5e6;     // :63 / B2a :64
    const double w1 = IF - BW / 2, w2 = IF + BW / 2;
    const double wp1 = w1 * 2 / fsOld - 0.002, wp2 = w2 * 2 / fsOld + 0.002;            // :67
    if (!(wp1 > 0.0 && wp2 < 1.0)) return set_error(BDS_ERR_ARG, "resampling: pass band [%g, %g] outside (0, 1)", wp1, wp2);
    const int order = 700;                                                              // fir1(700, wp), :69
    if (n <= (size_t)3 * order) return set_error(BDS_ERR_ARG, "resampling: filtfilt needs more than %d samples", 3 * order);
    std::vector<double> g;
    fir1_bandpass_autocorr(order, wp1, wp2, g);
    const double fu = IF + BW / 2, fl = IF - BW / 2;                                    // :79-95
    double nn = std::floor(fu / BW);
    if (nn < 1) nn = 1;
    const double lowerFreq = 2 * fu / nn;
    const double upperFreq = nn > 1 ? 2 * fl / (nn - 1) : lowerFreq;
    const double fsNew = std::ceil((lowerFreq + upperFreq) / 2);                        // :105
    const long long outLen = (long long)std::floor((double)(n - 1) / fsOld * fsNew);    // :109
    if (outLen <= 0) return set_error(BDS_ERR_ARG, "resampling: empty record");
    TRYA(dG.alloc(sizeof(double) * g.size()));
    TRYA(cudaMemcpy(dG.p, g.data(), sizeof(double) * g.size(), cudaMemcpyHostToDevice));
    TRYA(dRes.alloc(sizeof(float) * ((size_t)outLen << iq)));
    acq_resample_kernel<<<g_num_sms * 8, kResampleThreads, sizeof(double) * g.size()>>>(
        dx, iq, (long long)n, dG.as<double>(), order, fsNew, fsOld, outLen, dRes.as<float>());
    count_launch();
    TRYA(cudaGetLastError());
    bds_acq_cfg c2 = *cfg;
    c2.samplingFreq = fsNew;
    c2.IF = std::fmod(IF, fsNew);                                                       // :122
    c2.resamplingflag = 0;
    rc = acquire_core(signal, dRes.as<int8_t>(), kFmtF32 | iq, (size_t)outLen, &c2, prn, n_prn, prn_lo, prn_hi, carrFreq,
                      codePhase, peakMetric, max_prn, dbg);
    if (rc) return rc;
    // ---- results back at the original sampling rate (acquisition.m:321-338)
    for (int p = 0; p < max_prn; ++p) {
        if (carrFreq[p] == 0) continue;   // not detected
        codePhase[p] = std::floor((codePhase[p] - 1) / fsNew * fsOld) + 1;
        const double doppler = c2.IF >= fsNew / 2 ? (fsNew - c2.IF) - carrFreq[p] : carrFreq[p] - c2.IF;
        carrFreq[p] = doppler + IF;
    }
#undef TRYA
    return BDS_OK;
}

namespace bds {
void acq_pool_release() {
    std::lock_guard<std::mutex> lock(g_poolMu);
    for (auto& b : g_pool) cudaFree(b.p);
    g_pool.clear();
}
void acq_streams_release() {   // bds_shutdown only: no acquisition is in flight
    std::lock_guard<std::mutex> lk(g_streamsMu);
    int dev = -1;
    if (cudaGetDevice(&dev) == cudaSuccess && dev == g_streamsDev)
        for (auto& st : g_streams)
            if (st) cudaStreamDestroy(st);
    for (auto& st : g_streams) st = nullptr;
    g_streamsDev = -1;
}
}  // namespace bds

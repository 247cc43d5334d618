// Chip-synchronous fast path of the B1C wide-band correlator (sm_100a).
//
// Same arithmetic as BDS-3_B1C/WB_tracking.m:289-372 (nine +-1 replicas x carrier-wiped
// samples -> 18 sums), reorganised so that the per-sample work is ~3 instructions:
//   * one thread integrates one primary-code chip (~97 samples at 99.375 MHz);
//   * inside a chip all nine replicas are constant on 36 segments (gen_fast_wb.py); the
//     per-sample work is a fused int8 x Q16-carrier multiply-accumulate into the running
//     segment sum (2 IMAD + 1 PRMT); segment sums are folded into nine complex
//     "basis" sums (SA,SB,SC,H1,H2,W1a,W1b,W2a,W2b) from which E/P/L of data, BOC(1,1)
//     pilot and BOC(6,1) pilot follow by +-1 combinations with the chip signs;
//   * the carrier is exp(-i theta(n_c)) * exp(-i 2 pi r dphi): a per-epoch table of the
//     second factor (r = 0..97, Q16) is broadcast from shared memory, the first factor is
//     applied once per chip in fp32;
//   * the IF samples of a pass are staged into shared memory with one TMA bulk copy.
// Which of two neighbouring segments the sample at a boundary belongs to depends only on
// the sub-sample phase psi of the thread's first sample; it is looked up from a per-epoch
// table of sorted thresholds.  When psi is within 4e-9 of a threshold the chip is
// re-evaluated sample by sample with the exact float64 expressions of the general kernel,
// so chip-edge decisions are identical to the oracle's.
#pragma once
#include "bds_track.cuh"

namespace bds {

#include "bds_track_fast_gen.inc"

#ifndef BDS_FAST_BINREC
#define BDS_FAST_BINREC 0   // 1: rank search through a per-bin record (one 8-byte load instead of four dependent ones); opt-in build
#endif
constexpr int kFastBins = 128;
constexpr unsigned kFastGuard = 16u;  // fixed-point guard band (2^-32 units of one sample)

struct __align__(16) FastTab {
    int4 w[FAST_NWORDS + 1];            // per 4-sample word: {wr01, wr23, wi01, wi23}, int16 pairs, Q15 exp(-i 2 pi r dphi)
    unsigned thr[40];                   // sorted thresholds (2^32 fixed point), thr[36..] = 0xffffffff
    uint2 mask[40];                     // decision masks by rank
    unsigned char binStart[kFastBins + 16];
    unsigned char posbin[80];           // [0..35] sorted position of threshold k-1, [40..75] its bin (thr >> 25)
#if BDS_FAST_BINREC
    uint2 rec[kFastBins / 2];           // per pair of bins (Psi >> 26): the (at most two) thresholds inside, else 0xffffffff
#endif
    double u0, sigma, S;                // 12*rem, 12*step, 1/sigma
    unsigned long long dphi, phi0;      // carrier NCO, 2^-64 turns
    int valid;
    int pad[3];
};
static_assert(sizeof(FastTab) % 16 == 0, "FastTab must be a 16-byte multiple");

// B1C with a pilot: wide band (data + BOC(1,1) + BOC(6,1) pilot, 18 sums) or narrow band (the same minus the
// BOC(6,1) replica, 12 sums — the unused sums are zeroed by the epilogue warp).
inline bool fast_wb_supported(int mode, int hasPilot, int hasP61, double fs, double fc, int codeLength, double d) {
    const bool modeOk = (mode == BDS_TRK_B1C_WB && hasP61) || (mode == BDS_TRK_B1C_NB && hasPilot);
    return modeOk && codeLength == 10230 && fs == FAST_FS_HZ && fc == FAST_FC_HZ && d == FAST_D;
}

// ---- per-epoch table construction (one warp) -------------------------------------------------
// tab lives in global memory (one per channel, rewritten in place: every slice of epoch e has finished before
// epoch e+1's table is built); scratch: >= 128 words of shared memory private to the calling warp.
// prev (optional, shared memory): the posbin[] array of the table being replaced.  The sorted order of the 36
// thresholds and their bins almost never change from one epoch to the next (the code rate moves by ~1e-9), in
// which case masks / binStart / posbin are still right and only the threshold values, the carrier rotation
// table and the scalars are rewritten.
__device__ void fast_build_tab_warp(FastTab* tab, const EpochParams& np, double fs, unsigned* scratch,
                                    const unsigned char* prev = nullptr) {
    const int lane = threadIdx.x & 31;
    const double sigma = 12.0 * np.step, S = 1.0 / sigma;
    double r = np.carrFreq / fs;
    r -= floor(r);
    const unsigned long long dphi = __double2ull_rn(r * 18446744073709551616.0);
    double r0 = np.remCarr / 6.283185307179586476925286766559;   // independent of the above: overlaps its latency
    r0 -= floor(r0);
    const unsigned long long phi0 = __double2ull_rn(r0 * 18446744073709551616.0);
    short* w = reinterpret_cast<short*>(tab->w);   // word i: [wr(4i..4i+3) | wi(4i..4i+3)] as int16
    for (int t = lane; t < 4 * (FAST_NWORDS + 1); t += 32) {
        unsigned long long ph = (unsigned long long)t * dphi;
        float sn, cs;  // fp32 sincospi: 1e-7 accuracy, far below the Q15 quantisation step
        sincospif((float)(int)(ph >> 32) * 4.656612873077392578125e-10f, &sn, &cs);
        w[(t >> 2) * 8 + (t & 3)] = (short)__float2int_rn(cs * 32767.0f);
        w[(t >> 2) * 8 + 4 + (t & 3)] = (short)__float2int_rn(-sn * 32767.0f);
    }
    unsigned* thr = scratch;        // [36] unsorted thresholds
    unsigned* pos = scratch + 40;   // [36] sorted position of threshold k-1
    unsigned* srt = scratch + 80;   // [36] thresholds in sorted order (reuse path)
    int ok = 1;
    for (int t = lane; t < 36; t += 32) {
        const int k = t + 1;
        double th = kFastBeta[k] * S - (double)kFastR[k];  // theta_k / sigma
        ok &= (th > 1e-6 && th < 1.0 - 1e-6);
        th = fmin(fmax(th, 0.0), 1.0);
        thr[t] = (unsigned)fmin(th * 4294967296.0, 4294967295.0);
    }
    ok = __all_sync(0xffffffffu, ok);
    __syncwarp();
    bool reuse = prev != nullptr && ok;
    if (reuse) {   // same order, same bins as the table being replaced?
        int same = 1;
        for (int t = lane; t < 36; t += 32) {
            const unsigned p = prev[t];
            same &= p < 36u && (thr[t] >> 25) == (unsigned)prev[40 + t];
            srt[p < 36u ? p : 0] = thr[t];
        }
        same = __all_sync(0xffffffffu, same);
        __syncwarp();
        for (int t = lane; t < 35; t += 32) same &= srt[t] < srt[t + 1];
        reuse = __all_sync(0xffffffffu, same);
    }
    if (reuse) {
        for (int t = lane; t < 36; t += 32) tab->thr[t] = srt[t];
    } else {
        for (int t = lane; t < 40; t += 32) {
            if (t < 36) {  // rank sort (ties broken by index)
                const unsigned v = thr[t];
                int rank = 0;
                for (int j = 0; j < 36; ++j) rank += (thr[j] < v) || (thr[j] == v && j < t);
                tab->thr[rank] = v;
                pos[t] = rank;
                tab->posbin[t] = (unsigned char)rank;
                tab->posbin[40 + t] = (unsigned char)(v >> 25);
            } else {
                tab->thr[t] = 0xffffffffu;
            }
        }
        __syncwarp();
        for (int t = lane; t < 37; t += 32) {
            // mask[j]: bit (k-1) set  <=>  boundary sample R_k belongs to the OLD segment  <=>  Theta_k >= Psi
            //          <=> sorted position of k >= j   (j = number of thresholds < Psi)
            unsigned lo = 0, hi = 0;
            for (int k = 1; k <= 36; ++k)
                if ((int)pos[k - 1] >= t) {
                    if (k <= 32) lo |= 1u << (k - 1);
                    else hi |= 1u << (k - 33);
                }
            tab->mask[t] = make_uint2(lo, hi);
        }
        for (int t = lane; t < kFastBins + 1; t += 32) {
            int cnt = 0, here = 0;
            for (int j = 0; j < 36; ++j) {
                cnt += (thr[j] >> 25) < (unsigned)t;
                here += (thr[j] >> 25) == (unsigned)t;
            }
            tab->binStart[t] = (unsigned char)cnt;
            ok &= here <= 4;  // the rank refinement in the correlator does 4 steps
        }
        ok = __all_sync(0xffffffffu, ok);
    }
#if BDS_FAST_BINREC
    {
        __syncwarp();   // tab->thr (sorted) and tab->binStart are complete (this build or the table being reused)
        int okr = 1;
        for (int b = lane; b < kFastBins / 2; b += 32) {
            const int s0 = tab->binStart[2 * b], s1 = tab->binStart[2 * b + 2];
            tab->rec[b] = make_uint2(s0 < s1 ? tab->thr[s0] : 0xffffffffu, s0 + 1 < s1 ? tab->thr[s0 + 1] : 0xffffffffu);
            okr &= s1 - s0 <= 2;
        }
        okr = __all_sync(0xffffffffu, okr);
        if (!okr) {   // never with the nominal geometry (thresholds are >= 0.008 apart); keeps the result right regardless
            ok = 0;
            reuse = false;
        }
    }
#endif
    if (lane == 0) {
        tab->u0 = 12.0 * np.rem;
        tab->sigma = sigma;
        tab->S = S;
        tab->dphi = dphi;
        tab->phi0 = phi0;
        // reuse: `here <= 4` held for the replaced table (same bins), whose valid flag is left in place
        if (!reuse) tab->valid = ok;
    }
    __syncwarp();
}

// ---- exact per-sample evaluation (shared with the general kernel's arithmetic) --------------
struct ExactCtx {
    double a[3], stop[3], dd;
    unsigned long long dphi, phi0;
    int n;  // colon steps = blksize-1
};
__device__ inline void make_exact_ctx(const EpochParams& p, double d, double fs, ExactCtx& c) {
    c.n = p.blksize - 1;
    c.dd = __dmul_rn(p.step, 2.0);
    const double base = __dadd_rn(__dmul_rn((double)c.n, p.step), p.rem);
    c.a[0] = __dmul_rn(__dadd_rn(p.rem, -d), 2.0);
    c.a[1] = __dmul_rn(p.rem, 2.0);
    c.a[2] = __dmul_rn(__dadd_rn(p.rem, d), 2.0);
    c.stop[0] = __dmul_rn(__dadd_rn(base, -d), 2.0);
    c.stop[1] = __dmul_rn(base, 2.0);
    c.stop[2] = __dmul_rn(__dadd_rn(base, d), 2.0);
    double r = p.carrFreq / fs;
    r -= floor(r);
    c.dphi = __double2ull_rn(r * 18446744073709551616.0);
    double r0 = p.remCarr / 6.283185307179586476925286766559;
    r0 -= floor(r0);
    c.phi0 = __double2ull_rn(r0 * 18446744073709551616.0);
}
__device__ __forceinline__ double colon_elem_f(double a, double dd, double stop, int n, int k) {
    int h = n >> 1;
    if (!(n & 1) && k == h) return __dmul_rn(__dadd_rn(a, stop), 0.5);
    if (k <= h) return __dadd_rn(a, __dmul_rn((double)k, dd));
    return __dadd_rn(stop, -__dmul_rn((double)(n - k), dd));
}
__device__ __forceinline__ int bit_of(const uint32_t* w, int chip) { return (w[chip >> 5] >> (chip & 31)) & 1; }

// Accumulates block-relative samples k in [k0, k1] whose exact prompt BOC(6,1) index lies in
// [idxLo, idxHi] (chip membership exactly as the oracle decides it).
__device__ __noinline__ void fast_exact_range(const ExactCtx& c, const int8_t* xblk, const uint32_t* bitsData,
                                              const uint32_t* bitsPilot, int k0, int k1, int idxLo, int idxHi,
                                              float* acc) {
    for (int k = k0; k <= k1; ++k) {
        double tP = colon_elem_f(c.a[1], c.dd, c.stop[1], c.n, k);
        int i6 = (int)ceil(__dmul_rn(tP, 6.0));
        if (i6 < idxLo || i6 > idxHi) continue;
        float xs = (float)xblk[k];
        unsigned long long ph = c.phi0 + (unsigned long long)k * c.dphi;
        float sn, cs;
        sincospif((float)(int)(ph >> 32) * 4.656612873077392578125e-10f, &sn, &cs);
        float iB = xs * cs, qB = -xs * sn;
#pragma unroll
        for (int o = 0; o < 3; ++o) {
            double t = colon_elem_f(c.a[o], c.dd, c.stop[o], c.n, k);
            int h = (int)ceil(t) - 1;
            if (h < 0) h = 20459;
            if (h >= 20460) h = 0;
            float e0 = (h & 1) ? 1.f : -1.f;
            float sd = bit_of(bitsData, h >> 1) ? -e0 : e0;
            float sp = bit_of(bitsPilot, h >> 1) ? -e0 : e0;
            int s = (int)ceil(__dmul_rn(t, 6.0)) - 1;
            if (s < 0) s = 122759;
            if (s >= 122760) s = 0;
            int chip = s / 12;
            float e6 = ((s - chip * 12) & 1) ? 1.f : -1.f;
            float s6 = bit_of(bitsPilot, chip) ? -e6 : e6;
            acc[sum_idx(0, o, 0)] += sd * iB;
            acc[sum_idx(0, o, 1)] += sd * qB;
            acc[sum_idx(1, o, 0)] += sp * iB;
            acc[sum_idx(1, o, 1)] += sp * qB;
            acc[sum_idx(2, o, 0)] += s6 * iB;
            acc[sum_idx(2, o, 1)] += s6 * qB;
        }
    }
}

// ---- TMA bulk copy helpers -------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra WAIT_%=;\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}

// sign-extended byte b of a 32-bit word (PRMT with sign replication)
// (__byte_perm masks the selector nibbles to 3 bits, so the PTX form is needed for the sign mode)
template <int B>
__device__ __forceinline__ int sext_byte(unsigned w) {
    int r;
    asm("prmt.b32 %0, %1, 0, %2;" : "=r"(r) : "r"(w), "n"(B | ((B | 8) << 4) | ((B | 8) << 8) | ((B | 8) << 12)));
    return r;
}
// v (a constant) if bit BIT of m is set, else 0
template <int BIT>
__device__ __forceinline__ unsigned sel_bit_u(unsigned v, unsigned m) {
    unsigned r;
    asm("{\n .reg .pred p;\n .reg .b32 t;\n and.b32 t, %2, %3;\n setp.ne.u32 p, t, 0;\n selp.u32 %0, %1, 0, p;\n}"
        : "=r"(r)
        : "r"(v), "r"(m), "n"(1u << BIT));
    return r;
}
// s if bit BIT of m is set, else 0 (LOP3 with predicate result + SEL)
template <int BIT>
__device__ __forceinline__ int sel_bit(int s, unsigned m) {
    int r;
    asm("{\n .reg .pred p;\n .reg .b32 t;\n and.b32 t, %2, %3;\n setp.ne.u32 p, t, 0;\n selp.s32 %0, %1, 0, p;\n}"
        : "=r"(r)
        : "r"(s), "r"(m), "n"(1u << BIT));
    return r;
}

// ---- packed fp32 pairs (FFMA2 / FADD2 / FMUL2 of sm_100): an (I, Q) pair per instruction -------------
#ifndef BDS_FAST_F32X2
#define BDS_FAST_F32X2 0
#endif
#ifndef BDS_ABL
#define BDS_ABL 0   // developer ablations of fast_chip, bit mask (1: no rotation, 2: no rank search, 4: no per-sample body); never shipped
#endif
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t f2_pk(float a, float b) {
    f2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void f2_unpk(f2_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f2_t f2_fma(f2_t a, f2_t b, f2_t c) {
    f2_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f2_t f2_mul(f2_t a, f2_t b) {
    f2_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f2_t f2_add(f2_t a, f2_t b) {
    f2_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f2_t f2_sub(f2_t a, f2_t b) {
    f2_t d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f2_t f2_sfma(float s, f2_t b, f2_t c) { return f2_fma(f2_pk(s, s), b, c); }   // s * b + c

// the per-thread running sums of the compute warps: 18 floats, or 9 (I, Q) pairs in the packed build
#if BDS_FAST_F32X2
typedef f2_t fast_acc_t;
constexpr int kFastAccN = kNSum / 2;
__device__ __forceinline__ void fast_acc_zero(fast_acc_t* a) {
#pragma unroll
    for (int i = 0; i < kFastAccN; ++i) a[i] = 0ull;
}
__device__ __forceinline__ void fast_acc_add(fast_acc_t* a, const float* t) {
#pragma unroll
    for (int i = 0; i < kFastAccN; ++i) a[i] = f2_add(a[i], f2_pk(t[2 * i], t[2 * i + 1]));
}
__device__ __forceinline__ float fast_acc_get(const fast_acc_t* a, int i) {
    float x, y;
    f2_unpk(a[i >> 1], x, y);
    return (i & 1) ? y : x;
}
#else
typedef float fast_acc_t;
constexpr int kFastAccN = kNSum;
__device__ __forceinline__ void fast_acc_zero(fast_acc_t* a) {
#pragma unroll
    for (int i = 0; i < kNSum; ++i) a[i] = 0.f;
}
__device__ __forceinline__ void fast_acc_add(fast_acc_t* a, const float* t) {
#pragma unroll
    for (int i = 0; i < kNSum; ++i) a[i] += t[i];
}
__device__ __forceinline__ float fast_acc_get(const fast_acc_t* a, int i) { return a[i]; }
#endif

// ---- one chip (one thread) ---------------------------------------------------------------------
// Integrates chip c of the epoch described by (tab, p) into acc[18].  tile/tileBase: staged IF
// bytes (window byte offset of tile[0]); xblk = g.x + B0 for the exact path.  Returns true if the
// chip went through the exact per-sample path.
__device__ __forceinline__ bool fast_chip(const FastTab& tab, const EpochParams& p, const uint32_t* bitsData,
                                          const uint32_t* bitsPilot, const unsigned char* tile, long long tileBase,
                                          int tileBytes, long long B0, const int8_t* xblk, double dSpacing, double fs, int c,
                                          unsigned guard, fast_acc_t* acc) {
    // ---- per-chip phase bookkeeping (fp64) ----
    const double q = ((double)(12 * c) - tab.u0) * tab.S;  // sample position of the chip start
    const int nc = (int)floor(q) + 1;                        // first sample of the chip
    const double psi = (double)nc - q;                       // in (0,1] samples
    const unsigned Psi = (unsigned)fmin(psi * 4294967296.0, 4294967295.0);
#if BDS_FAST_BINREC
    // thresholds of this pair of bins come with one load; a threshold of a neighbouring pair within the guard band of
    // Psi implies that Psi is within the guard band of the pair's edge
    const uint2 rc = tab.rec[Psi >> 26];
    const int j = tab.binStart[(Psi >> 26) * 2] + (rc.x < Psi) + (rc.y < Psi);
    const uint2 mk = tab.mask[j];
    const unsigned eg = guard < 4096u ? guard : 4096u, low = Psi & 0x3ffffffu;
    bool exact = !tab.valid || rc.x - Psi + guard <= 2u * guard || rc.y - Psi + guard <= 2u * guard || low <= eg ||
                 low >= 0x3ffffffu - eg || Psi >= 0xffffffffu - guard;
#else
    int j = tab.binStart[Psi >> 25];
#if BDS_ABL & 2   // developer ablation (wrong results): no rank refinement, no guard band
    const uint2 mk = tab.mask[j];
    bool exact = !tab.valid;
#else
#pragma unroll
    for (int it = 0; it < 4; ++it) j += (tab.thr[j] < Psi);
    const uint2 mk = tab.mask[j];
    // near-miss of any decision (including the chip start/end) -> exact path
    const unsigned below = j > 0 ? Psi - tab.thr[j - 1] : Psi;
    const unsigned above = j < 36 ? tab.thr[j] - Psi : 0xffffffffu - Psi;
    bool exact = !tab.valid || below <= guard || above <= guard || Psi >= 0xffffffffu - guard;
#endif
#endif  // BDS_FAST_BINREC
    const int len = FAST_RLAST + ((mk.y >> 3) & 1);          // bit 35 (k = 36): last sample still mine
    if (nc < 0 || nc + len > p.blksize) exact = true;
    const long long o = B0 + nc - tileBase;   // the chip's first sample inside the staged bytes
    if (o < 0 || o + 4 * (FAST_NWORDS + 1) > (long long)tileBytes) exact = true;
    if (!exact) {
        const unsigned* raw = reinterpret_cast<const unsigned*>(tile) + (o >> 2);
        const unsigned sh = (unsigned)(o & 3) * 8u;
        const int4* wt = tab.w;
        FAST_DECL_ACCS
#define FAST_RAW(i) raw[i]
#define FAST_FSH(lo, hi) __funnelshift_r(lo, hi, sh)
#define FAST_WTAB(i) wt[i]
#define FAST_DP_LO(a, b, c) __dp2a_lo((int)(a), (int)(b), (c))
#define FAST_DP_HI(a, b, c) __dp2a_hi((int)(a), (int)(b), (c))
#define FAST_SELU(k, v) ((k) <= 32 ? sel_bit_u<((k)-1) & 31>(v, mk.x) : sel_bit_u<((k)-33) & 31>(v, mk.y))
#if BDS_ABL & 4   // developer ablation (wrong results): no per-sample body, one word of the tile per chip
        const int v0 = (int)__funnelshift_r(raw[0], raw[1], sh) + wt[0].x;
        const int SAr = v0, SAi = v0, SBr = v0, SBi = v0, SCr = v0, SCi = v0, H1r = v0, H1i = v0, H2r = v0, H2i = v0,
                  W1ar = v0, W1ai = v0, W1br = v0, W1bi = v0, W2ar = v0, W2ai = v0, W2br = v0, W2bi = v0;
        (void)mk;
#else
        FAST_CHIP_BODY
        FAST_COMBINE
#endif
#undef FAST_RAW
#undef FAST_FSH
#undef FAST_WTAB
#undef FAST_DP_LO
#undef FAST_DP_HI
#undef FAST_SELU
#if BDS_ABL & 1   // developer ablation (wrong results): no rotation, no chip-sign combination
        {
            float tmp[kNSum] = {(float)SAr, (float)SAi, (float)SBr, (float)SBi, (float)SCr, (float)SCi,
                                (float)H1r, (float)H1i, (float)H2r, (float)H2i, (float)W1ar, (float)W1ai,
                                (float)W1br, (float)W1bi, (float)W2ar, (float)W2ai, (float)W2br, (float)W2bi};
            fast_acc_add(acc, tmp);
        }
#else
        // ---- chip-level: rotate by exp(-i theta(nc)), combine with chip signs ----
        const unsigned long long ph = tab.phi0 + (unsigned long long)(long long)nc * tab.dphi;
        const float ang = (float)(int)(ph >> 32) * 1.4629180792671596e-9f;  // 2*pi / 2^32, |ang| <= pi
        const float sn = __sinf(ang), cs = __cosf(ang);
        const float rr = cs * (1.0f / 32767.0f), ri = -sn * (1.0f / 32767.0f);
#if BDS_FAST_F32X2
        // the same sums with an (I, Q) pair per instruction
        const f2_t rotA = f2_pk(rr, ri), rotB = f2_pk(-ri, rr);
#define ROT(N) const f2_t N##p = f2_sfma((float)N##i, rotB, f2_mul(f2_pk((float)N##r, (float)N##r), rotA));
        ROT(SA) ROT(SB) ROT(SC) ROT(H1) ROT(H2) ROT(W1a) ROT(W1b) ROT(W2a) ROT(W2b)
#undef ROT
        const int cp_ = c == 0 ? 10229 : c - 1, cn_ = c == 10229 ? 0 : c + 1;
        const float cd = bit_of(bitsData, c) ? -1.f : 1.f, cdp = bit_of(bitsData, cp_) ? -1.f : 1.f,
                    cdn = bit_of(bitsData, cn_) ? -1.f : 1.f;
        const float cp = bit_of(bitsPilot, c) ? -1.f : 1.f, cpp = bit_of(bitsPilot, cp_) ? -1.f : 1.f,
                    cpn = bit_of(bitsPilot, cn_) ? -1.f : 1.f;
        const f2_t Xp = f2_sub(H2p, H1p);
        const f2_t XEp = f2_sfma(-2.f, W1bp, f2_add(Xp, W1ap));
        const f2_t XLp = f2_sfma(2.f, W2ap, f2_sub(Xp, W2bp));
        const f2_t sAB = f2_add(SAp, SBp), sBC = f2_add(SBp, SCp);
        const f2_t SPp = f2_add(sAB, SCp), SEp = f2_sub(SCp, sAB), SLp = f2_sub(SAp, sBC);
#define ACC2(fam, epl, expr)                                       \
    {                                                              \
        const f2_t a0 = acc[sum_idx(fam, epl, 0) >> 1];            \
        acc[sum_idx(fam, epl, 0) >> 1] = expr;                     \
    }
        ACC2(0, EPL_P, f2_sfma(cd, Xp, a0))
        ACC2(0, EPL_E, f2_sfma(cd, XEp, f2_sfma(cdp, W1ap, a0)))
        ACC2(0, EPL_L, f2_sfma(cd, XLp, f2_sfma(-cdn, W2bp, a0)))
        ACC2(1, EPL_P, f2_sfma(cp, Xp, a0))
        ACC2(1, EPL_E, f2_sfma(cp, XEp, f2_sfma(cpp, W1ap, a0)))
        ACC2(1, EPL_L, f2_sfma(cp, XLp, f2_sfma(-cpn, W2bp, a0)))
        ACC2(2, EPL_P, f2_sfma(cp, SPp, a0))
        ACC2(2, EPL_E, f2_sfma(cp, SEp, f2_sfma(cpp - cp, W1ap, a0)))
        ACC2(2, EPL_L, f2_sfma(cp, SLp, f2_sfma(cp - cpn, W2bp, a0)))
#undef ACC2
#else
#define ROT(N) const float N##x = (float)N##r * rr - (float)N##i * ri, N##y = (float)N##r * ri + (float)N##i * rr;
        ROT(SA) ROT(SB) ROT(SC) ROT(H1) ROT(H2) ROT(W1a) ROT(W1b) ROT(W2a) ROT(W2b)
#undef ROT
        const int cp_ = c == 0 ? 10229 : c - 1, cn_ = c == 10229 ? 0 : c + 1;
        const float cd = bit_of(bitsData, c) ? -1.f : 1.f, cdp = bit_of(bitsData, cp_) ? -1.f : 1.f,
                    cdn = bit_of(bitsData, cn_) ? -1.f : 1.f;
        const float cp = bit_of(bitsPilot, c) ? -1.f : 1.f, cpp = bit_of(bitsPilot, cp_) ? -1.f : 1.f,
                    cpn = bit_of(bitsPilot, cn_) ? -1.f : 1.f;
        const float Xx = H2x - H1x, Xy = H2y - H1y;
        const float XEx = Xx + W1ax - 2.f * W1bx, XEy = Xy + W1ay - 2.f * W1by;
        const float XLx = Xx + 2.f * W2ax - W2bx, XLy = Xy + 2.f * W2ay - W2by;
        acc[sum_idx(0, EPL_P, 0)] += cd * Xx;
        acc[sum_idx(0, EPL_P, 1)] += cd * Xy;
        acc[sum_idx(0, EPL_E, 0)] += cd * XEx + cdp * W1ax;
        acc[sum_idx(0, EPL_E, 1)] += cd * XEy + cdp * W1ay;
        acc[sum_idx(0, EPL_L, 0)] += cd * XLx - cdn * W2bx;
        acc[sum_idx(0, EPL_L, 1)] += cd * XLy - cdn * W2by;
        acc[sum_idx(1, EPL_P, 0)] += cp * Xx;
        acc[sum_idx(1, EPL_P, 1)] += cp * Xy;
        acc[sum_idx(1, EPL_E, 0)] += cp * XEx + cpp * W1ax;
        acc[sum_idx(1, EPL_E, 1)] += cp * XEy + cpp * W1ay;
        acc[sum_idx(1, EPL_L, 0)] += cp * XLx - cpn * W2bx;
        acc[sum_idx(1, EPL_L, 1)] += cp * XLy - cpn * W2by;
        const float SPx = SAx + SBx + SCx, SPy = SAy + SBy + SCy;
        const float SEx = SCx - SAx - SBx, SEy = SCy - SAy - SBy;
        const float SLx = SAx - SBx - SCx, SLy = SAy - SBy - SCy;
        acc[sum_idx(2, EPL_P, 0)] += cp * SPx;
        acc[sum_idx(2, EPL_P, 1)] += cp * SPy;
        acc[sum_idx(2, EPL_E, 0)] += cp * SEx + (cpp - cp) * W1ax;
        acc[sum_idx(2, EPL_E, 1)] += cp * SEy + (cpp - cp) * W1ay;
        acc[sum_idx(2, EPL_L, 0)] += cp * SLx + (cp - cpn) * W2bx;
        acc[sum_idx(2, EPL_L, 1)] += cp * SLy + (cp - cpn) * W2by;
#endif
#endif  // BDS_ABL & 1
    } else {
        // rare (~1e-5 of chips): keep the fast path's accumulators in registers
        ExactCtx ex;
        make_exact_ctx(p, dSpacing, fs, ex);
        const double qe = ((double)(12 * c + 12) - tab.u0) * tab.S;
        int k0 = max(0, nc - 2), k1 = min(p.blksize - 1, (int)floor(qe) + 3);
        float tmp[kNSum];
#pragma unroll
        for (int i = 0; i < kNSum; ++i) tmp[i] = 0.f;
        fast_exact_range(ex, xblk, bitsData, bitsPilot, k0, k1, 12 * c + 1, 12 * c + 12, tmp);
        fast_acc_add(acc, tmp);
    }
    return exact;
}

}  // namespace bds

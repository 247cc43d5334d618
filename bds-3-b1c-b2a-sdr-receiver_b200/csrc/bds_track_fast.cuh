// Chip-synchronous fast path of the B1C wide-band correlator (sm_100a).
//
// Same arithmetic as BDS-3_B1C/WB_tracking.m:289-372 (nine +-1 replicas x carrier-wiped
// samples -> 18 sums), reorganised so that the per-sample work is ~3 instructions:
//   * one thread integrates one primary-code chip (~97 samples at 99.375 MHz);
//   * inside a chip all nine replicas are constant on 36 segments (gen_fast_wb.py); a re-aligned word of four
//     int8 samples is AND-masked to a segment (one LOP3 with the prefix masks of the segment's boundaries, each a
//     SEL between two constants on a bit of the rank's decision mask) and multiplied with the Q15 carrier by IDP.2A into the segment's class accumulator; class sums are
//     folded into eight complex "basis" sums (X = H2 - H1, SA, SB, SC, W1a, W1b, W2a, W2b) from which E/P/L of data, BOC(1,1)
//     pilot and BOC(6,1) pilot follow by +-1 combinations with the chip signs;
//   * the carrier is exp(-i theta(n_c)) * exp(-i 2 pi r dphi): a per-epoch table of the
//     second factor (r = 0..103, Q15) is broadcast from shared memory, the first factor is
//     applied once per chip in fp32;
//   * the IF samples of a pass are staged into shared memory with one TMA bulk copy.
// Which of two neighbouring segments the sample at a boundary belongs to depends only on
// the sub-sample phase psi of the thread's first sample: its rank among the epoch's 36 thresholds (one table
// lookup + one compare, the thresholds' order being a generation-time constant).  When psi is within 4e-9 of a
// threshold the chip is re-evaluated sample by sample with the exact float64 expressions of the general kernel,
// so chip-edge decisions are identical to the oracle's.
//
// The file also compiles for the host (tests/test_fast_b1c_hostcompile.py includes it after a shim that gives the
// CUDA intrinsics host definitions), so the arithmetic of a build is checked against the oracle before it gets GPU time.
//
// Geometries: the chip body, its tables and everything sized by them live in namespace bds::FAST_GEOM_NS and come from
// the generated file FAST_GEN_INC (gen_fast_wb.py).  Included as it is, the header gives the BASELINE geometry
// (g99: fs = 99.375 MHz) and makes its names visible in namespace bds.  bds_track.cu defines FAST_GEOM_MULTI and includes
// the header once per geometry (g99, and g53 = the reference's shipped fs = 53 MHz, B1C/initSettings.m:57); the
// geometry-independent helpers are compiled by the first inclusion only.
#include "bds_track.cuh"

#ifndef FAST_GEOM_NS
#define FAST_GEOM_NS g99
#define FAST_GEN_INC "bds_track_fast_gen.inc"
#endif
#ifndef FAST_CONST
#define FAST_CONST static __constant__
#endif

#ifndef BDS_TRACK_FAST_COMMON
namespace bds {
constexpr int kFastBins = 128;        // (B2a body: bins of its rank index)
constexpr unsigned kFastGuard = 16u;  // fixed-point guard band (2^-32 units of one sample)
}
#endif

namespace bds {
namespace FAST_GEOM_NS {
#include FAST_GEN_INC
constexpr unsigned kFastNomTol = 1u << (32 - FAST_RANK_BITS - 1);   // a threshold may sit this far from its nominal value (half a rank bin)

// Per-epoch table of one channel, built in shared memory by the table-builder warp of the consuming CTA from the
// epoch's NCO parameters alone (48 bytes travel from the loop closure to the correlator, nothing else).
struct __align__(16) FastTab {
    int4 w[FAST_NWORDS + 1];            // per 4-sample word: {wr01, wr23, wi01, wi23}, int16 pairs, Q15 exp(-i 2 pi r dphi)
    unsigned thr[40];                   // thresholds in the (generation-time) sorted order, 2^32 fixed point; [FAST_NSEG..] = 0xffffffff
    double u0, sigma, S;                // 12*rem, 12*step, 1/sigma
    unsigned long long dphi, phi0;      // carrier NCO, 2^-64 turns
    int valid;                          // thresholds within kFastNomTol of nominal (else every chip takes the exact path)
    int pad[3];
};
static_assert(sizeof(FastTab) % 16 == 0, "FastTab must be a 16-byte multiple");

// Generation-time tables (gen_fast_wb.py), copied once per CTA into shared memory: lanes index them with different
// ranks, which constant memory would serialise.
struct __align__(16) FastStatic {
    uint2 mask[40];                     // [rank]: bit k-1 set <=> the jitter sample of boundary k is old
    unsigned thrNom[36];                // nominal sorted thresholds
    unsigned char rankLo[FAST_RANK_BINS];
    unsigned char pos[36];              // sorted position of threshold k-1
    unsigned char pad_[12];
};
static_assert(sizeof(FastStatic) % 16 == 0, "FastStatic must be a 16-byte multiple");

// B1C with a pilot: wide band (data + BOC(1,1) + BOC(6,1) pilot, 18 sums) or narrow band (the same minus the
// BOC(6,1) replica, 12 sums — the unused sums are zeroed by the epilogue warp).
inline bool fast_wb_supported(int mode, int hasPilot, int hasP61, double fs, double fc, int codeLength, double d) {
#if FAST_NB   // narrow-band body: data + BOC(1,1) pilot only
    (void)hasP61;
    const bool modeOk = mode == BDS_TRK_B1C_NB && hasPilot;
#else
    const bool modeOk = (mode == BDS_TRK_B1C_WB && hasP61) || (mode == BDS_TRK_B1C_NB && hasPilot);
#endif
    return modeOk && codeLength == 10230 && fs == FAST_FS_HZ && fc == FAST_FC_HZ && d == FAST_D;
}

// one thread's share of the copy of the static tables (tid = 0 .. nthreads-1; followed by a CTA barrier)
__device__ inline void fast_load_static(FastStatic* fsx, int tid, int nthreads) {
    for (int i = tid; i < FAST_NSEG + 1; i += nthreads)
        fsx->mask[i] = make_uint2((unsigned)kFastMask[i], (unsigned)(kFastMask[i] >> 32));
    for (int i = tid; i < FAST_NSEG; i += nthreads) {
        fsx->thrNom[i] = kFastThrNom[i];
        fsx->pos[i] = kFastPos[i];
    }
    for (int i = tid; i < FAST_RANK_BINS; i += nthreads) fsx->rankLo[i] = kFastRankLo[i];
}

// ---- per-epoch table construction --------------------------------------------------------------
// One lane's share (lane = 0..31): carrier rotation entries lane, lane+32, ... and thresholds lane+1, lane+33.
// Returns 0 if one of its thresholds is further than kFastNomTol from nominal.  No cross-lane dependency.
__device__ inline int fast_build_tab_lane(FastTab* tab, const FastStatic& fsx, const EpochParams& np, double fs, int lane) {
    const double sigma = 12.0 * np.step, S = 1.0 / sigma;
    double r = np.carrFreq / fs;
    r -= floor(r);
    const unsigned long long dphi = __double2ull_rn(r * 18446744073709551616.0);
    short* w = reinterpret_cast<short*>(tab->w);   // word i: [wr(4i..4i+3) | wi(4i..4i+3)] as int16
    for (int t = lane; t < 4 * (FAST_NWORDS + 1); t += 32) {
        // MUFU sine / cosine of the exactly reduced phase (|angle| <= pi, abs. error 4e-7): a tenth of the Q15 step,
        // i.e. at most the odd last bit of a table entry; this runs once per pass on a service warp, so it is kept short
        const unsigned long long ph = (unsigned long long)t * dphi;
        const float ang = (float)(int)(ph >> 32) * 1.4629180792671596e-9f;  // 2*pi / 2^32
        w[(t >> 2) * 8 + (t & 3)] = (short)__float2int_rn(__cosf(ang) * 32767.0f);
        w[(t >> 2) * 8 + 4 + (t & 3)] = (short)__float2int_rn(-__sinf(ang) * 32767.0f);
    }
    int ok = 1;
    for (int t = lane; t < 40; t += 32) {
        if (t < FAST_NSEG) {
            const int k = t + 1;
            double th = kFastBeta[k] * S - (double)kFastR[k];  // theta_k / sigma
            th = fmin(fmax(th, 0.0), 1.0);
            const unsigned v = (unsigned)fmin(th * 4294967296.0, 4294967295.0);
            const int p = fsx.pos[t];
            const unsigned nom = fsx.thrNom[p];
            ok &= (v - nom + kFastNomTol) < 2u * kFastNomTol;
            tab->thr[p] = v;
        } else {
            tab->thr[t] = 0xffffffffu;
        }
    }
    if (lane == 0) {
        double r0 = np.remCarr / 6.283185307179586476925286766559;
        r0 -= floor(r0);
        tab->u0 = 12.0 * np.rem;
        tab->sigma = sigma;
        tab->S = S;
        tab->dphi = dphi;
        tab->phi0 = __double2ull_rn(r0 * 18446744073709551616.0);
    }
    return ok;
}
#ifdef __CUDACC__
// whole warp; tab in shared memory
__device__ inline void fast_build_tab_warp(FastTab* tab, const FastStatic& fsx, const EpochParams& np, double fs) {
    const int lane = threadIdx.x & 31;
    const int ok = __all_sync(0xffffffffu, fast_build_tab_lane(tab, fsx, np, fs, lane));
    if (lane == 0) tab->valid = ok;
    __syncwarp();
}
#endif

}  // namespace FAST_GEOM_NS
}  // namespace bds

#ifndef BDS_TRACK_FAST_COMMON
#define BDS_TRACK_FAST_COMMON
namespace bds {
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

#ifdef __CUDACC__
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

// hi if bit BIT of m is set, else lo (both constants): R2P + SEL
template <int BIT>
__device__ __forceinline__ unsigned fast_psel(unsigned m, unsigned lo, unsigned hi) {
    unsigned r;
    asm("{\n .reg .pred p;\n .reg .b32 t;\n and.b32 t, %3, %4;\n setp.ne.u32 p, t, 0;\n selp.u32 %0, %1, %2, p;\n}"
        : "=r"(r)
        : "r"(hi), "r"(lo), "r"(m), "n"(1u << BIT));
    return r;
}
// v (a constant) if bit BIT of m is set, else 0   (B2a body)
template <int BIT>
__device__ __forceinline__ unsigned sel_bit_u(unsigned v, unsigned m) {
    unsigned r;
    asm("{\n .reg .pred p;\n .reg .b32 t;\n and.b32 t, %2, %3;\n setp.ne.u32 p, t, 0;\n selp.u32 %0, %1, 0, p;\n}"
        : "=r"(r)
        : "r"(v), "r"(m), "n"(1u << BIT));
    return r;
}

// ---- packed fp32 pairs (FFMA2 / FADD2 / FMUL2 of sm_100): an (I, Q) pair per instruction -------------
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
#endif  // __CUDACC__  (the host build brings its own definitions of the helpers above)
__device__ __forceinline__ f2_t f2_sfma(float s, f2_t b, f2_t c) { return f2_fma(f2_pk(s, s), b, c); }   // s * b + c

#ifndef BDS_ABL
#define BDS_ABL 0   // developer ablations of fast_chip, bit mask (1: no rotation, 2: no rank search, 4: no per-sample body); never shipped
#endif

// the per-thread running sums of the compute warps: 9 (I, Q) pairs
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

}  // namespace bds
#endif  // BDS_TRACK_FAST_COMMON

namespace bds {
namespace FAST_GEOM_NS {
// ---- one chip (one thread) ---------------------------------------------------------------------
// Integrates chip c of the epoch described by (tab, p) into acc[9 pairs].  tile/tileBase: staged IF
// bytes (window byte offset of tile[0]); xblk = g.x + B0 for the exact path.  Returns true if the
// chip went through the exact per-sample path.
__device__ __forceinline__ bool fast_chip(const FastTab& tab, const FastStatic& fsx, const EpochParams& p,
                                          const uint32_t* bitsData, const uint32_t* bitsPilot, const unsigned char* tile,
                                          long long tileBase, int tileBytes, long long B0, const int8_t* xblk, double dSpacing,
                                          double fs, int c, unsigned guard, fast_acc_t* acc) {
    // ---- per-chip phase bookkeeping (fp64) ----
    const double q = ((double)(12 * c) - tab.u0) * tab.S;  // sample position of the chip start
    const int nc = (int)floor(q) + 1;                        // first sample of the chip
    const double psi = (double)nc - q;                       // in (0,1] samples
    const unsigned Psi = (unsigned)fmin(psi * 4294967296.0, 4294967295.0);
    // rank = number of thresholds < Psi.  The thresholds keep their generation-time order and stay within half a
    // rank bin (1/512 or 1/1024 sample) of nominal (tab.valid), nominal neighbours are > 3 bins apart: the only threshold that can lie within
    // a bin of Psi - and hence the only one that can decide the rank beyond rankLo, or come within the guard band -
    // is thr[rankLo[bin]].
    const int jlo = fsx.rankLo[Psi >> (32 - FAST_RANK_BITS)];
    static_assert(FAST_RANK_BINS == 1 << FAST_RANK_BITS, "rank bins");
    const unsigned tnear = tab.thr[jlo];
    const int j = jlo + (tnear < Psi);
    const uint2 mk = fsx.mask[j];
    // near-miss of any decision (including the chip start/end) -> exact path
    bool exact = !tab.valid || tnear - Psi + guard <= 2u * guard || Psi <= guard || Psi >= 0xffffffffu - guard;
    const int len = FAST_RLAST + (j <= FAST_POS_LAST);       // boundary 36 old: the last sample is still mine
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
#define FAST_PSEL(k, lo, hi) ((k) <= 32 ? fast_psel<((k)-1) & 31>(mk.x, lo, hi) : fast_psel<((k)-33) & 31>(mk.y, lo, hi))
#define FAST_DP_LO(a, b, c) __dp2a_lo((int)(a), (int)(b), (c))
#define FAST_DP_HI(a, b, c) __dp2a_hi((int)(a), (int)(b), (c))
#if BDS_ABL & 4   // developer ablation (wrong results): no per-sample body, one word of the tile per chip
        const int v0 = (int)__funnelshift_r(raw[0], raw[1], sh) + wt[0].x + (int)mk.x;
        const int SAr = v0, SAi = v0, SBr = v0, SBi = v0, SCr = v0, SCi = v0, Xr = v0, Xi = v0,
                  W1ar = v0, W1ai = v0, W1br = v0, W1bi = v0, W2ar = v0, W2ai = v0, W2br = v0, W2bi = v0;
#else
        FAST_CHIP_BODY
        FAST_COMBINE
#endif
#undef FAST_RAW
#undef FAST_FSH
#undef FAST_WTAB
#undef FAST_PSEL
#undef FAST_DP_LO
#undef FAST_DP_HI
#if BDS_ABL & 1   // developer ablation (wrong results): no rotation, no chip-sign combination
        {
            float tmp[kNSum] = {(float)SAr, (float)SAi, (float)SBr, (float)SBi, (float)SCr, (float)SCi,
                                (float)Xr, (float)Xi, (float)Xr, (float)Xi, (float)W1ar, (float)W1ai,
                                (float)W1br, (float)W1bi, (float)W2ar, (float)W2ai, (float)W2br, (float)W2bi};
            fast_acc_add(acc, tmp);
        }
#else
        // ---- chip-level: rotate by exp(-i theta(nc)), combine with chip signs; an (I, Q) pair per instruction ----
        const unsigned long long ph = tab.phi0 + (unsigned long long)(long long)nc * tab.dphi;
        const float ang = (float)(int)(ph >> 32) * 1.4629180792671596e-9f;  // 2*pi / 2^32, |ang| <= pi
        const float sn = __sinf(ang), cs = __cosf(ang);
        const float rr = cs * (1.0f / 32767.0f), ri = -sn * (1.0f / 32767.0f);
        const f2_t rotA = f2_pk(rr, ri), rotB = f2_pk(-ri, rr);
#define ROT(N) const f2_t N##p = f2_sfma((float)N##i, rotB, f2_mul(f2_pk((float)N##r, (float)N##r), rotA));
#if !FAST_NB
        ROT(SA) ROT(SB) ROT(SC)
#endif
        ROT(X) ROT(W1a) ROT(W1b) ROT(W2a) ROT(W2b)
#undef ROT
        const int cp_ = c == 0 ? 10229 : c - 1, cn_ = c == 10229 ? 0 : c + 1;
        const float cd = bit_of(bitsData, c) ? -1.f : 1.f, cdp = bit_of(bitsData, cp_) ? -1.f : 1.f,
                    cdn = bit_of(bitsData, cn_) ? -1.f : 1.f;
        const float cp = bit_of(bitsPilot, c) ? -1.f : 1.f, cpp = bit_of(bitsPilot, cp_) ? -1.f : 1.f,
                    cpn = bit_of(bitsPilot, cn_) ? -1.f : 1.f;
        const f2_t XEp = f2_sfma(-2.f, W1bp, f2_add(Xp, W1ap));
        const f2_t XLp = f2_sfma(2.f, W2ap, f2_sub(Xp, W2bp));
#if !FAST_NB
        const f2_t sAB = f2_add(SAp, SBp), sBC = f2_add(SBp, SCp);
        const f2_t SPp = f2_add(sAB, SCp), SEp = f2_sub(SCp, sAB), SLp = f2_sub(SAp, sBC);
#endif
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
#if !FAST_NB
        ACC2(2, EPL_P, f2_sfma(cp, SPp, a0))
        ACC2(2, EPL_E, f2_sfma(cp, SEp, f2_sfma(cpp - cp, W1ap, a0)))
        ACC2(2, EPL_L, f2_sfma(cp, SLp, f2_sfma(cp - cpn, W2bp, a0)))
#endif
#undef ACC2
#endif  // BDS_ABL & 1
    } else {
        // rare (~1e-6 of chips): keep the fast path's accumulators in registers
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

}  // namespace FAST_GEOM_NS
#ifndef FAST_GEOM_MULTI
using namespace FAST_GEOM_NS;
#endif
}  // namespace bds

// Chip-synchronous B2a tracking (sm_100a): one thread-block cluster (1, 2, 4 or 8 CTAs) per channel, epoch and loop
// closure inside the cluster.
//
// BDS-3_B2a/tracking.m integrates 1 ms (10 230 chips, 99 375 samples at 99.375 MHz) per loop update, so a channel
// closes its loops 1000 times per second of signal: the latency of one closure bounds the speed, not the arithmetic.
// The multi-CTA scheduler of the B1C kernel (slices of an epoch on many SMs, a closer CTA, three L2 round trips per
// closure) is the wrong shape for that; here a channel owns one SM:
//   * the epoch's samples (99 KB) are staged into shared memory with one TMA bulk copy into one of two buffers, issued
//     for epoch e+1 when the correlation of epoch e starts (the start of the next block is pos + blksize, whatever the
//     loops decide), so the copy is never waited for;
//   * 512 threads integrate the 1 023 ten-chip units of the epoch (two per thread) with the generated IDP.2A body
//     of gen_fast_b2a.py: 20 half-chip segments per unit, the chip signs applied by integer adds at the end of the
//     unit, one fp32 rotation per unit and replica;
//   * warp sums in Q8 fixed point (REDUX) -> warp 0: discriminators in parallel lanes, loop filters and the next NCO on
//     lane 0 (close_nco, the general kernel's arithmetic) -> the next table by five warps;
//   * a service warp (the 17th) issues the bulk copies and writes the trackResults planes of epoch e-1 (close_out /
//     close_cno) while the compute warps correlate epoch e: neither is on the closure path;
//   * with a cluster of CS CTAs per channel (latency mode: few channels per GPU) every CTA stages and integrates 1/CS of
//     the epoch's units, the CTAs exchange their twelve Q8 integer sums through distributed shared memory (one
//     st.shared::cluster per value and peer, one cluster barrier) and EVERY CTA closes the loops on the identical
//     integer totals - bit-identical NCO state in all of them, so nothing has to be sent back; CTA 0 alone writes
//     outputs and the channel state.
// Segment-edge decisions use the same fixed-point thresholds + guard band as the B1C body; units that come within
// the guard band of a decision, or that touch the ends of the block, are evaluated sample by sample with the float64
// expressions of the general kernel (tracking.m:262-296 incl. the two-ended colon), so chip lookups are the oracle's.
//
// Status: the AUTO choice for B2a (validated on a B200 in round 2).  The generated body and the chip-sign combination are verified on the CPU
// (tests/test_fast_body_emulation.py), the kernel on hardware (tests/test_gpu_b2a_unit.py).
#pragma once
#ifndef BDS_TRACK_FAST_COMMON
#include "bds_track_fast.cuh"
#endif

namespace bds {

#include "bds_track_fast_b2a_gen.inc"

constexpr int kB2aThreads = 512;                  // compute threads
constexpr int kB2aThreadsAll = kB2aThreads + 32;  // + the service warp (bulk copies, outputs of the previous epoch)
constexpr int kB2aMaxCluster = 8;
constexpr int kB2aUnits = 10230 / FASTB_CHIPS;   // thread units per epoch
constexpr int kB2aTileBytes = 99584;             // one block (<= 99 380 samples) + 16-byte alignment + the word reads of
                                                 // the last unit; a multiple of 128
static_assert(kB2aUnits * FASTB_CHIPS == 10230, "unit size must divide the code length");

struct __align__(16) FastbTab {
    int4 w[FASTB_NWORDS + 1];           // per 4-sample word: {wr01, wr23, wi01, wi23}, int16 pairs, Q15 exp(-i 2 pi r dphi)
    unsigned thr[24];                   // sorted thresholds (2^32 fixed point), thr[20..] = 0xffffffff
    unsigned mask[24];                  // decision masks by rank: bit k-1 set <=> sample R_k stays in the old segment
    unsigned char binStart[kFastBins + 16];
    double u0, sigma, S;                // 2*rem, 2*step, 1/sigma (half-chip units)
    unsigned long long dphi, phi0;      // carrier NCO, 2^-64 turns
    int valid;
    int pad[3];
};

inline bool fastb_supported(int mode, double fs, double fc, int codeLength, double d) {
    return mode == BDS_TRK_B2A && codeLength == 10230 && fs == FASTB_FS_HZ && fc == FASTB_FC_HZ && d == FASTB_D;
}

// ---- per-epoch table (one warp; tab and scratch (>= 64 words) in shared memory) ------------------------------
// carrier rotation entries t = first, first + stride, ... of the table (any group of threads)
__device__ inline void fastb_build_rot(FastbTab* tab, unsigned long long dphi, int first, int stride) {
    short* w = reinterpret_cast<short*>(tab->w);   // word i: [wr(4i..4i+3) | wi(4i..4i+3)] as int16
    for (int t = first; t < 4 * (FASTB_NWORDS + 1); t += stride) {
        unsigned long long ph = (unsigned long long)t * dphi;
        float sn, cs;
        sincospif((float)(int)(ph >> 32) * 4.656612873077392578125e-10f, &sn, &cs);
        w[(t >> 2) * 8 + (t & 3)] = (short)__float2int_rn(cs * 32767.0f);
        w[(t >> 2) * 8 + 4 + (t & 3)] = (short)__float2int_rn(-sn * 32767.0f);
    }
}
__device__ inline unsigned long long fastb_dphi(const EpochParams& np, double fs) {
    double r = np.carrFreq / fs;
    r -= floor(r);
    return __double2ull_rn(r * 18446744073709551616.0);
}

// the 128-bin rank index: bins [bin0, bin0 + 32) by one warp (entry kFastBins by the warp of bin0 == 0), from thresholds
// the warp evaluates for itself in its own scratch (>= 32 words) - same expressions as fastb_build_tab_warp
__device__ __forceinline__ void fastb_build_bins_warp(FastbTab* tab, const EpochParams& np, unsigned* scratch, int bin0) {
    const int lane = threadIdx.x & 31;
    const double S = 1.0 / (2.0 * np.step);
    if (lane < FASTB_NSEG) {
        double th = kFastbBeta[lane + 1] * S - (double)kFastbR[lane + 1];
        th = fmin(fmax(th, 0.0), 1.0);
        scratch[lane] = (unsigned)fmin(th * 4294967296.0, 4294967295.0) >> 25;
    }
    __syncwarp();
    auto put = [&](int t) {
        int cnt = 0;
#pragma unroll
        for (int j = 0; j < FASTB_NSEG; ++j) cnt += scratch[j] < (unsigned)t;
        tab->binStart[t] = (unsigned char)cnt;
    };
    put(bin0 + lane);
    if (bin0 == 0 && lane == 0) put(kFastBins);
    __syncwarp();
}

// reuseOrder (scratch >= 96 words, kept between calls): the thresholds move by ~1e-6 of a sample from one epoch to the
// next, so their ORDER almost never changes; when the previous build's order still sorts them (one comparison per
// lane) only the threshold values are refreshed - the rank sort and the decision masks, which depend on the order
// alone, are skipped.
__device__ void fastb_build_tab_warp(FastbTab* tab, const EpochParams& np, double fs, unsigned* scratch, bool withRotation = true,
                                     bool withBins = true, bool reuseOrder = false) {
    const int lane = threadIdx.x & 31;
    const double sigma = 2.0 * np.step, S = 1.0 / sigma;
    double r = np.carrFreq / fs;
    r -= floor(r);
    const unsigned long long dphi = __double2ull_rn(r * 18446744073709551616.0);
    double r0 = np.remCarr / 6.283185307179586476925286766559;
    r0 -= floor(r0);
    const unsigned long long phi0 = __double2ull_rn(r0 * 18446744073709551616.0);
    if (withRotation) fastb_build_rot(tab, dphi, lane, 32);
    unsigned* thr = scratch;        // [20] unsorted thresholds
    unsigned* pos = scratch + 32;   // [20] sorted position of threshold k-1
    int ok = 1;
    if (lane < FASTB_NSEG) {
        const int k = lane + 1;
        double th = kFastbBeta[k] * S - (double)kFastbR[k];  // theta_k / sigma
        ok = (th > 1e-6 && th < 1.0 - 1e-6);
        th = fmin(fmax(th, 0.0), 1.0);
        thr[lane] = (unsigned)fmin(th * 4294967296.0, 4294967295.0);
    }
    ok = __all_sync(0xffffffffu, ok);
    __syncwarp();
    int sameOrder = 0;
    if (reuseOrder) {   // inv[i] = threshold with sorted position i (previous build): still sorted, ties by index?
        const unsigned* inv = scratch + 64;
        sameOrder = 1;
        if (lane + 1 < FASTB_NSEG) {
            const unsigned a = inv[lane], b = inv[lane + 1];
            sameOrder = a < (unsigned)FASTB_NSEG && b < (unsigned)FASTB_NSEG && (thr[a] < thr[b] || (thr[a] == thr[b] && a < b));
        }
        sameOrder = __all_sync(0xffffffffu, sameOrder);
    }
    if (sameOrder) {
        if (lane < FASTB_NSEG) tab->thr[pos[lane]] = thr[lane];
        __syncwarp();
    } else {
        if (lane < FASTB_NSEG) {  // rank sort (ties broken by index)
            const unsigned v = thr[lane];
            int rank = 0;
            for (int j = 0; j < FASTB_NSEG; ++j) rank += (thr[j] < v) || (thr[j] == v && j < lane);
            tab->thr[rank] = v;
            pos[lane] = rank;
            if (reuseOrder) scratch[64 + rank] = (unsigned)lane;
        } else if (lane < 24) {
            tab->thr[lane] = 0xffffffffu;
        }
        __syncwarp();
        if (lane <= FASTB_NSEG) {
            // mask[j]: bit (k-1) set <=> Theta_k >= Psi <=> sorted position of k >= j (j = number of thresholds < Psi)
            unsigned m = 0;
            for (int k = 1; k <= FASTB_NSEG; ++k)
                if ((int)pos[k - 1] >= lane) m |= 1u << (k - 1);
            tab->mask[lane] = m;
        }
    }
    if (withBins) {
        for (int t = lane; t < kFastBins + 1; t += 32) {
            int cnt = 0;
            for (int j = 0; j < FASTB_NSEG; ++j) cnt += (thr[j] >> 25) < (unsigned)t;
            tab->binStart[t] = (unsigned char)cnt;
        }
    }
    // the rank refinement in the correlator does 4 steps: no bin may hold more than 4 thresholds, i.e. sorted
    // thresholds four places apart lie in different bins
    if (lane + 4 < FASTB_NSEG) ok &= (tab->thr[lane] >> 25) != (tab->thr[lane + 4] >> 25);
    ok = __all_sync(0xffffffffu, ok);
    if (lane == 0) {
        tab->u0 = 2.0 * np.rem;
        tab->sigma = sigma;
        tab->S = S;
        tab->dphi = dphi;
        tab->phi0 = phi0;
        tab->valid = ok;
    }
    __syncwarp();
}

// ---- exact per-sample evaluation (the general kernel's B2a arithmetic, bds_track.cu correlate_general) --------
__device__ inline void make_exact_ctx_b2a(const EpochParams& p, double d, double fs, ExactCtx& c) {
    c.n = p.blksize - 1;
    c.dd = p.step;
    const double base = __dadd_rn(__dmul_rn((double)c.n, p.step), p.rem);
    c.a[0] = __dadd_rn(p.rem, -d);
    c.a[1] = p.rem;
    c.a[2] = __dadd_rn(p.rem, d);
    c.stop[0] = __dadd_rn(base, -d);
    c.stop[1] = base;
    c.stop[2] = __dadd_rn(base, d);
    double r = p.carrFreq / fs;
    r -= floor(r);
    c.dphi = __double2ull_rn(r * 18446744073709551616.0);
    double r0 = p.remCarr / 6.283185307179586476925286766559;
    r0 -= floor(r0);
    c.phi0 = __double2ull_rn(r0 * 18446744073709551616.0);
}

// Accumulates block-relative samples k in [k0, k1] whose exact prompt chip index ceil(t) lies in [idxLo, idxHi].
__device__ __noinline__ void fastb_exact_range(const ExactCtx& c, const int8_t* xblk, const uint32_t* bitsData,
                                               const uint32_t* bitsPilot, int k0, int k1, int idxLo, int idxHi, float* acc) {
    for (int k = k0; k <= k1; ++k) {
        const double tP = colon_elem_f(c.a[1], c.dd, c.stop[1], c.n, k);
        const int ip = (int)ceil(tP);
        if (ip < idxLo || ip > idxHi) continue;
        const float xs = (float)xblk[k];
        const unsigned long long ph = c.phi0 + (unsigned long long)k * c.dphi;
        float sn, cs;
        sincospif((float)(int)(ph >> 32) * 4.656612873077392578125e-10f, &sn, &cs);
        const float qB = xs * cs, iB = xs * sn;   // exp(+i theta), I = imag, Q = real: tracking.m:309-314
#pragma unroll
        for (int o = 0; o < 3; ++o) {
            const double t = colon_elem_f(c.a[o], c.dd, c.stop[o], c.n, k);
            int cc = (int)ceil(t) - 1;            // padded table [code(end) code code(1)], tracking.m:262-296
            if (cc < 0) cc = 10229;
            if (cc >= 10230) cc = 0;
            const float sd = bit_of(bitsData, cc) ? -1.f : 1.f;
            const float sp = bit_of(bitsPilot, cc) ? -1.f : 1.f;
            acc[sum_idx(0, o, 0)] += sd * iB;
            acc[sum_idx(0, o, 1)] += sd * qB;
            acc[sum_idx(1, o, 0)] += sp * iB;
            acc[sum_idx(1, o, 1)] += sp * qB;
        }
    }
}

// code bits of chips 10u-1 .. 10u+10 (bit j+1 <-> chip 10u+j) from the code rotated by one chip (b2a_load_bits):
// ext bit n = chip n-1, ext bit 0 = the last chip, ext bit 10231 = the first one, so no unit needs a special case
__device__ __forceinline__ unsigned fastb_code12(const uint32_t* ext, int u) {
    const int start = FASTB_CHIPS * u;
    const int i0 = start >> 5;
    return __funnelshift_r(ext[i0], ext[min(i0 + 1, kPackedWordsDev - 1)], start & 31) & 0xfffu;
}

// ---- one unit of ten chips (one thread) ------------------------------------------------------------------------
// Integrates unit u of the epoch (tab, p) into acc[18] (families 0 = data, 1 = pilot; family 2 stays zero).  Returns
// true if the unit went through the exact per-sample path.
__device__ __forceinline__ bool fastb_unit(const FastbTab& tab, const EpochParams& p, const uint32_t* bitsData,
                                           const uint32_t* bitsPilot, const uint32_t* extData, const uint32_t* extPilot,
                                           const unsigned char* tile, long long tileBase, int tileBytes, long long B0,
                                           const int8_t* xblk, double dSpacing, double fs, int u, unsigned guard, float* acc) {
    const double q = ((double)(2 * FASTB_CHIPS * u) - tab.u0) * tab.S;  // sample position of the unit start
    const int nc = (int)floor(q) + 1;                                    // first sample of the unit
    const double psi = (double)nc - q;                                   // in (0,1] samples
    const unsigned Psi = (unsigned)fmin(psi * 4294967296.0, 4294967295.0);
    int j = tab.binStart[Psi >> 25];
#pragma unroll
    for (int it = 0; it < 4; ++it) j += (tab.thr[j] < Psi);
    const unsigned mk = tab.mask[j];
    const unsigned below = j > 0 ? Psi - tab.thr[j - 1] : Psi;
    const unsigned above = j < FASTB_NSEG ? tab.thr[j] - Psi : 0xffffffffu - Psi;
    bool exact = !tab.valid || below <= guard || above <= guard || Psi >= 0xffffffffu - guard;
    const int len = FASTB_RLAST + ((mk >> (FASTB_NSEG - 1)) & 1);        // bit 19 (k = 20): last sample still mine
    if (nc < 0 || nc + len > p.blksize) exact = true;
    const long long o = B0 + nc - tileBase;   // the unit's first sample inside the staged bytes
    if (o < 0 || o + 4 * (FASTB_NWORDS + 1) > (long long)tileBytes) exact = true;
    if (!exact) {
        const unsigned* raw = reinterpret_cast<const unsigned*>(tile) + (o >> 2);
        const unsigned sh = (unsigned)(o & 3) * 8u;
        const int4* wt = tab.w;
        FASTB_DECL_ACCS
#define FASTB_RAW(i) raw[i]
#define FASTB_FSH(lo, hi) __funnelshift_r(lo, hi, sh)
#define FASTB_WTAB(i) wt[i]
#define FASTB_DP_LO(a, b, c) __dp2a_lo((int)(a), (int)(b), (c))
#define FASTB_DP_HI(a, b, c) __dp2a_hi((int)(a), (int)(b), (c))
#define FASTB_SELU(k, v) sel_bit_u<((k)-1) & 31>(v, mk)
        FASTB_CHIP_BODY
        const unsigned dbits = fastb_code12(extData, u), pbits = fastb_code12(extPilot, u);
#define FASTB_CD(j) (1 - 2 * (int)((dbits >> ((j) + 1)) & 1u))
#define FASTB_CP(j) (1 - 2 * (int)((pbits >> ((j) + 1)) & 1u))
        FASTB_COMBINE
#undef FASTB_CD
#undef FASTB_CP
#undef FASTB_RAW
#undef FASTB_FSH
#undef FASTB_WTAB
#undef FASTB_DP_LO
#undef FASTB_DP_HI
#undef FASTB_SELU
        // ---- unit level: rotate by exp(-i theta(nc)); B2a mixes with exp(+i theta): Q = Re, I = -Im of the conjugate sum
        const unsigned long long ph = tab.phi0 + (unsigned long long)(long long)nc * tab.dphi;
        const float ang = (float)(int)(ph >> 32) * 1.4629180792671596e-9f;  // 2*pi / 2^32, |ang| <= pi
        const float sn = __sinf(ang), cs = __cosf(ang);
        const float rr = cs * (1.0f / 32767.0f), ri = -sn * (1.0f / 32767.0f);
#define ROTB(N, fam, epl)                                                                    \
    {                                                                                        \
        const float x_ = (float)N##r * rr - (float)N##i * ri, y_ = (float)N##r * ri + (float)N##i * rr; \
        acc[sum_idx(fam, epl, 0)] -= y_;                                                     \
        acc[sum_idx(fam, epl, 1)] += x_;                                                     \
    }
        ROTB(DE, 0, EPL_E) ROTB(DP, 0, EPL_P) ROTB(DL, 0, EPL_L)
        ROTB(PE, 1, EPL_E) ROTB(PP, 1, EPL_P) ROTB(PL, 1, EPL_L)
#undef ROTB
    } else {
        ExactCtx ex;
        make_exact_ctx_b2a(p, dSpacing, fs, ex);
        const double qe = ((double)(2 * FASTB_CHIPS * (u + 1)) - tab.u0) * tab.S;
        const int k0 = max(0, nc - 2), k1 = min(p.blksize - 1, (int)floor(qe) + 3);
        float tmp[kNSum];
#pragma unroll
        for (int i = 0; i < kNSum; ++i) tmp[i] = 0.f;
        fastb_exact_range(ex, xblk, bitsData, bitsPilot, k0, k1, FASTB_CHIPS * u + 1, FASTB_CHIPS * u + FASTB_CHIPS, tmp);
#pragma unroll
        for (int i = 0; i < kNSum; ++i) acc[i] += tmp[i];
    }
    return exact;
}

// ---- thread-block cluster primitives (a host emulation defines B2A_CLUSTER_SHIM and its own versions) -------------
#ifndef B2A_CLUSTER_SHIM
__device__ __forceinline__ unsigned b2a_cluster_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned b2a_cluster_size() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
// every thread of every CTA of the cluster arrives (release: its earlier shared::cluster stores are visible to whoever
// completes the wait) and waits (acquire).  Used once at the start (mbarrier initialisation) and once at the end.
__device__ __forceinline__ void b2a_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
// exchange barrier: an mbarrier in every CTA on which the lanes of the PEERS' warp 0 arrive after their stores into
// this CTA (release at cluster scope); only this CTA's warp 0 waits on it (acquire) - the other 500 threads of the
// CTA never take part in a cluster-wide barrier
__device__ __forceinline__ void b2a_xbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void b2a_xbar_arrive_peer(unsigned long long* bar, unsigned rank) {
    unsigned ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
// arm the local exchange barrier for one phase: one arrival (this one) + `bytes` of incoming bulk copies
__device__ __forceinline__ void b2a_xbar_expect(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bulk copy of `bytes` (a multiple of 16) from this CTA's shared memory into the same-named buffer of CTA `rank`,
// completing on THAT CTA's barrier (cp.async.bulk, shared::cta -> shared::cluster): one instruction per peer
__device__ __forceinline__ void b2a_bulk_to_peer(void* dstLocalName, const void* src, unsigned bytes, unsigned long long* barLocalName,
                                                 unsigned rank) {
    unsigned rd, rb;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rd) : "r"(smem_u32(dstLocalName)), "r"(rank));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(smem_u32(barLocalName)), "r"(rank));
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(rd),
                 "r"(smem_u32(src)), "r"(bytes), "r"(rb)
                 : "memory");
}
__device__ __forceinline__ void b2a_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// the three loop-closing warps (0..2) meet
__device__ __forceinline__ void b2a_closers_sync() { asm volatile("bar.sync 1, 96;" ::: "memory"); }
// ... and the ten warps (0..9) that close the loops / build the next epoch's table
__device__ __forceinline__ void b2a_builders_sync() { asm volatile("bar.sync 2, 320;" ::: "memory"); }
__device__ __forceinline__ void b2a_xbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "XWAIT_%=:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra XWAIT_%=;\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// v -> the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ void b2a_st_peer(long long* p, unsigned rank, long long v) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    unsigned ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
    asm volatile("st.shared::cluster.s64 [%0], %1;" ::"r"(ra), "l"(v) : "memory");
}
#endif

// ---- shared memory of the per-channel CTA ------------------------------------------------------------------------
struct __align__(128) B2aSmem {
    unsigned char tile[2][kB2aTileBytes];   // epoch e in tile[e & 1]: the next block is staged while this one is correlated
    FastbTab tab;
    uint32_t bits[2][kPackedWordsDev];  // packed primaries: data, pilot (bit k of word w = chip 32 w + k)
    uint32_t ext[2][kPackedWordsDev];   // the same rotated by one chip (fastb_code12)
    int res[kB2aThreads / 32][12];      // warp sums, Q8 fixed point
    __align__(16) long long part[2][kB2aMaxCluster][12];   // [epoch parity][cluster rank]: every CTA's Q8 sums, written by the CTAs themselves
    __align__(16) long long mine[2][12];     // this CTA's Q8 sums (source of the bulk copies to the peers)
    long long bcast[2];                 // lock-loss handling: CTA 0's {lockLost, lowLock} after a C/N0 interval
    unsigned long long xbar[2];         // exchange barriers by epoch parity: 96 bytes from every CTA of the cluster
    unsigned long long bbar;            // followers: CTA 0's bcast has arrived
    unsigned binScratch[4][32];         // thresholds of the warps that build the rank index
    double sums[kNSum];
    double v[12];                       // loop closure: {data, pilot} x {E, L, P} x {I, Q} as the discriminators take them
    double pre[8];                      // discriminator pieces evaluated by parallel lanes (close_nco)
    EpochParams p;                      // the epoch being correlated / about to be correlated
    EpochParams pDone;                  // the epoch whose loops were just closed (output phase)
    ChanState st;                       // channel state (thread 0 closes the loops on it)
    CloseAux aux;
    unsigned scratch[96];               // fastb_build_tab_warp: thresholds, sorted positions, inverse order (kept between epochs)
    EpochParams pNext;                  // the next epoch's NCO as far as the table needs it (step, rem, carrFreq, remCarr)
    unsigned long long full[2];         // mbarriers: the staged block has arrived
    long long tileBase[2];              // window byte offset of tile[b][0]
    int tileBytes[2];
    int run;
    int u0, u1;                         // this CTA integrates units [u0, u1) of every epoch
    int tileLo, tileSpan;               // ... which lie in bytes [tileLo, tileLo + tileSpan) of a block (nominal chip rate + margin)
    int eDone;                          // index of the epoch in pDone
    int outDone;                        // the outputs of epoch eDone were already written (lock-loss handling, interval end)
    double outv[kNFields];
};

// whole CTA (no barrier inside): the channel's code bits and their rotated copy
__device__ __forceinline__ void b2a_load_bits(const TrkDev& g, B2aSmem& sm, int c) {
    const uint32_t* src = g.codeBits + (size_t)c * 2 * kPackedWordsDev;
    for (int i = threadIdx.x; i < 2 * kPackedWordsDev; i += (int)blockDim.x) {
        const int f = i / kPackedWordsDev, k = i - f * kPackedWordsDev;
        const uint32_t* w = src + f * kPackedWordsDev;
        const uint32_t cur = __ldg(w + k);
        const uint32_t carry = k ? __ldg(w + k - 1) >> 31 : (__ldg(w + (10229 >> 5)) >> (10229 & 31)) & 1u;
        uint32_t e = (cur << 1) | carry;
        if (k == (10231 >> 5)) e |= (__ldg(w) & 1u) << (10231 & 31);
        sm.bits[f][k] = cur;
        sm.ext[f][k] = e;
    }
}

// one thread: stage this CTA's part [tileLo, tileLo + tileSpan) of the block that starts at absolute sample `pos`
// (whatever of it the window holds) into buffer b; returns the bytes requested (0: nothing to load, the barrier is not
// armed).  Units whose samples fall outside the staged bytes take the exact path from global memory.
__device__ __forceinline__ int b2a_issue_tile(const TrkDev& g, B2aSmem& sm, long long pos, int b) {
    const long long off = pos - g.winFirst;
    const long long base = (off + sm.tileLo) & ~15LL;
    const long long lim = g.winStage;                  // what a staged tile may cover (see TrkDev)
    long long bytes = lim - base;
    if (bytes > sm.tileSpan) bytes = sm.tileSpan;
    sm.tileBase[b] = base;
    if (off < 0 || bytes <= 0) {
        sm.tileBytes[b] = 0;
        return 0;
    }
    sm.tileBytes[b] = (int)bytes;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the tile was last read through the generic proxy
    tma_load_1d(sm.tile[b], g.x + base, (unsigned)bytes, &sm.full[b]);
    return (int)bytes;
}

// the correlator part of one epoch, whole CTA: per-warp Q8 sums -> sm.res
__device__ __forceinline__ void b2a_correlate(const TrkDev& g, B2aSmem& sm, const EpochParams& p, unsigned guard,
                                              unsigned& nFast, unsigned& nExact, int b) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long B0 = p.pos - g.winFirst;
    float acc[kNSum];
#pragma unroll
    for (int i = 0; i < kNSum; ++i) acc[i] = 0.f;
    if (tid == 0 && sm.u0 == 0 && p.rem == 0.0) {
        // remCodePhase == 0: the t = 0 sample reads the padded table's first entry, the previous period's last chip
        ExactCtx ex;
        make_exact_ctx_b2a(p, g.d, g.fs, ex);
        fastb_exact_range(ex, g.x + B0, sm.bits[0], sm.bits[1], 0, 0, -100, 0, acc);
    }
    for (int u = sm.u0 + tid; u < sm.u1; u += kB2aThreads) {
        const bool ex = fastb_unit(sm.tab, p, sm.bits[0], sm.bits[1], sm.ext[0], sm.ext[1], sm.tile[b], sm.tileBase[b],
                                   sm.tileBytes[b], B0, g.x + B0, g.d, g.fs, u, guard, acc);
        nFast += !ex;
        nExact += ex;
    }
    __syncwarp();   // lanes leave the unit loop after one or two units
#pragma unroll
    for (int i = 0; i < 12; ++i) {   // warp sums in Q8 fixed point (exact integer adds from here on)
        const int s = __reduce_add_sync(0xffffffffu, __float2int_rn(acc[i] * 256.f));
        if (lane == 0) sm.res[warp][i] = s;
    }
}

// warp 0 after the CTA barrier: this CTA's 12 Q8 sums (exact integers)
__device__ __forceinline__ long long b2a_cta_sum(const B2aSmem& sm, int lane) {
    long long t = 0;
    if (lane < 12) {
#pragma unroll
        for (int w = 0; w < kB2aThreads / 32; ++w) t += sm.res[w][lane];
    }
    return t;
}
// warp 0: 12 sums as doubles from the Q8 totals of the nParts contributions in part[] (the pilot family is zero without
// a pilot, like the general kernel).  Integer adds: the total does not depend on the order or the cluster size.
__device__ __forceinline__ void b2a_collect(const TrkDev& g, B2aSmem& sm, const long long (*part)[12], int nParts) {
    const int lane = threadIdx.x & 31;
    if (lane < kNSum) {
        double v = 0.0;
        if (lane < 6 || (lane < 12 && g.hasPilot)) {
            long long t = 0;
            for (int r = 0; r < nParts; ++r) t += part[r][lane];
            v = (double)t * (1.0 / 256.0);
        }
        sm.sums[lane] = v;
    }
    __syncwarp();
}

// exact fmod(x, y) for 0 <= x, 0 < y, x/y < 2^52 (one fma; libdevice fmod iterates ~20 times here)
__device__ __forceinline__ double b2a_fmod_pos(double x, double y) {
    if (!(x >= 0.0)) return fmod(x, y);
    const double n = floor(x / y);
    double r = fma(-n, y, x);
    if (r < 0.0) r += y;
    if (r >= y) r -= y;
    return r;
}

// one of the 12 sums as a double from the Q8 totals in part[] (zero for the pilot family without a pilot)
__device__ __forceinline__ double b2a_total(const TrkDev& g, const long long (*part)[12], int nParts, int idx) {
    if (!(idx < 6 || g.hasPilot)) return 0.0;
    long long t = 0;
    for (int r = 0; r < nParts; ++r) t += part[r][idx];
    return (double)t * (1.0 / 256.0);
}

// The discriminator pieces of tracking.m:337-377 (the same expressions as close_nco evaluates when it gets no `pre`):
// square roots, arctangents, divisions and the exact fmod of the carrier phase are long dependent fp64 chains - what
// makes a one-thread loop closure slow.  They are independent of one another, so three warps evaluate them side by
// side (a warp's lanes share one instruction stream: inside one warp the chains would run one after the other):
//   warp 0  |E|, |L| of data and pilot -> the two DLL discriminators        pre[2], pre[3]
//   warp 1  atan(Q_P / I_P) / 2 pi of data and (rotated) pilot              pre[0], pre[1]
//   warp 2  rem(trigarg(blksize + 1), 2 pi), the next carrier phase         pre[4]
__device__ __forceinline__ void b2a_disc_dll(const TrkDev& g, B2aSmem& sm, const long long (*part)[12], int n) {
    const int lane = threadIdx.x & 31;
    const int l4 = lane & 3;   // 0: data E, 1: data L, 2: pilot E, 3: pilot L
    const int fam = l4 >> 1, epl = (l4 & 1) ? EPL_L : EPL_E;
    const double A = b2a_total(g, part, n, sum_idx(fam, epl, 0)), B = b2a_total(g, part, n, sum_idx(fam, epl, 1));
    const double mag = sqrt(A * A + B * B);
    const double magN = __shfl_down_sync(0xffffffffu, mag, 1);
    const double disc = (mag - magN) / (mag + magN);
    if (lane == 0) sm.pre[2] = disc;
    if (lane == 2) sm.pre[3] = disc;
}
__device__ __forceinline__ void b2a_disc_pll(const TrkDev& g, B2aSmem& sm, const long long (*part)[12], int n) {
    const int lane = threadIdx.x & 31;
    const int fam = lane & 1;   // 0: data prompt, 1: pilot prompt rotated by exp(-i pi/2) (tracking.m:345)
    const double iP = b2a_total(g, part, n, sum_idx(fam, EPL_P, 0)), qP = b2a_total(g, part, n, sum_idx(fam, EPL_P, 1));
    const double cr = 6.123233995736766e-17;  // cos(pi/2) in double
    const double A = fam ? iP * cr + qP : iP, B = fam ? qP * cr - iP : qP;
    const double ang = atan(B / A) / 6.283185307179586476925286766559;
    if (lane < 2) sm.pre[lane] = ang;
}
__device__ __forceinline__ void b2a_disc_carr(const TrkDev& g, B2aSmem& sm, const EpochParams& p) {
    const double trig = ((p.carrFreq * 2.0 * 3.14159265358979323846) * ((double)p.blksize / g.fs)) + p.remCarr;
    const double fm = b2a_fmod_pos(trig, 6.283185307179586476925286766559);   // tracking.m:305 (trig >= 0)
    if ((threadIdx.x & 31) == 0) sm.pre[4] = fm;
}

// this CTA's share of an epoch: units [u0, u1) and the bytes of a block they can touch at any plausible code rate
__device__ __forceinline__ void b2a_plan_share(const TrkDev& g, B2aSmem& sm, int c, unsigned rk, unsigned CS) {
    const int U = (kB2aUnits + (int)CS - 1) / (int)CS;
    sm.u0 = min((int)rk * U, kB2aUnits);
    sm.u1 = min(sm.u0 + U, kB2aUnits);
    const double spu = (double)FASTB_CHIPS * g.fs / g.cc[c].chCodeFreq;   // samples per unit at the nominal chip rate
    const int margin = 96;   // samples: the code phase at the block start (< 1 sample) plus any plausible code Doppler over 1 ms
    const int lo = rk == 0 ? 0 : max(0, (int)(sm.u0 * spu) - margin);
    const int hi = rk + 1 == CS ? kB2aTileBytes : min(kB2aTileBytes, (int)(sm.u1 * spu) + margin + 4 * (FASTB_NWORDS + 1) + 32);
    sm.tileLo = lo;
    sm.tileSpan = (hi - (lo & ~15) + 15) & ~15;   // the copy starts at a 16-byte boundary at or below lo
}

// the service warp of CTA 0: trackResults planes (and, at the end of a C/N0 interval, C/N0 + lock detector) of the epoch
// whose loops were closed last
__device__ __forceinline__ void b2a_write_outputs(const TrkDev& g, B2aSmem& sm, int c, double* out, int cap) {
    const int lane = threadIdx.x & 31;
    const int e = sm.eDone;
    if (lane == 0) close_out(g, sm.sums, sm.pDone, sm.aux, sm.outv);
    __syncwarp();
    for (int f = lane; f < kNFields; f += 32)
        if (field_written(g, f)) out[(size_t)f * cap + e] = sm.outv[f];
    if (g.cnoInterval > 0 && (e + 1) % g.cnoInterval == 0) {
        __threadfence();
        __syncwarp();
        if (lane == 0) close_cno(g, c, e, sm.st);   // touches cnoPrev / the lock counters only
    }
    __syncwarp();
}

// Closed loop: grid = active channels x cluster size, one cluster per channel.  Every launch runs each channel for up to
// g.maxEpochs epochs from its device-side state, never past epoch index g.epochLimit, as far as the resident window
// allows (a short read stops the channel exactly like tracking.m:246-251).  Per epoch:
//   compute warps: correlate this CTA's units | service warp: stage the next block, write the previous epoch's outputs
//   CTA barrier | warp 0: Q8 sums of the CTA -> every CTA of the cluster | cluster barrier
//   warp 0 (every CTA, identical arithmetic): totals, discriminators in parallel lanes, loop filters + next NCO on lane 0
//   CTA barrier | the next epoch's table (warp 0: thresholds, warps 1-4: carrier rotation) | CTA barrier.
__global__ void __launch_bounds__(kB2aThreadsAll, 1) trk_b2a_unit_kernel(TrkDev g) {
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    B2aSmem& sm = *reinterpret_cast<B2aSmem*>(dyn_smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned CS = b2a_cluster_size(), rk = b2a_cluster_rank();
    const bool lead = rk == 0;                       // writes the channel's outputs and state
    const bool service = warp == kB2aThreads / 32;   // bulk copies + outputs, off the closure path
    const int c = g.act[blockIdx.x / CS];
    if (!g.cc[c].active) return;                     // uniform over the cluster
    b2a_load_bits(g, sm, c);
    int done = 0, pending[2] = {0, 0};   // service lane 0: a bulk copy into tile[b] is in flight
    double* const out = g.out + (size_t)c * kNFields * g.capacity;
    const int cap = g.capacity;
    const unsigned guard = g.pad ? (1u << 24) : kFastGuard;   // g.pad: test hook, widens the guard band
    unsigned nFast = 0, nExact = 0;
    unsigned phase[2] = {0, 0};
    int buf = 0, par = 0;
    if (tid == 0) {
        mbar_init(&sm.full[0], 1);
        mbar_init(&sm.full[1], 1);
        b2a_xbar_init(&sm.xbar[0], 1);   // one local arrival (expect_tx) + 96 bytes from every CTA of the cluster per phase
        b2a_xbar_init(&sm.xbar[1], 1);
        b2a_xbar_init(&sm.bbar, 1);
        sm.tileBytes[0] = sm.tileBytes[1] = 0;
        for (int i = 64; i < 96; ++i) sm.scratch[i] = 0xffffffffu;   // no previous threshold order
        b2a_plan_share(g, sm, c, rk, CS);
        sm.st = g.st[c];
        EpochParams np;
        const bool okp = next_params(g, sm.st, np), lim = sm.st.epoch < g.epochLimit;
        if (lead && !okp && lim && sm.st.lockLost == 0) out[(size_t)F_ABS * cap + sm.st.epoch] = (double)sm.st.pos;   // tracking.m:228 precedes the failed read
        sm.run = okp && lim && g.maxEpochs > 0;
        sm.eDone = -1;
        sm.outDone = 0;
        if (sm.run) sm.p = np;
    }
    __syncthreads();
    if (CS > 1) b2a_cluster_sync();   // every CTA's exchange barriers are initialised before anybody arrives on them
    int eCur = sm.st.epoch;   // index of the epoch about to be correlated (every thread keeps its own copy)
    unsigned xph[2] = {0, 0}, bph = 0;
    if (service && lane == 0 && sm.run) pending[0] = b2a_issue_tile(g, sm, sm.p.pos, 0);
    if (warp == 0 && sm.run) fastb_build_tab_warp(&sm.tab, sm.p, g.fs, sm.scratch, true, true, true);
    __syncthreads();
#ifdef BDS_FW_DEV
    long long tW = 0, tC = 0, tL = 0, tT = 0, tX = 0, tD = 0, tN = 0, t0_ = 0;
#define B2A_T(acc) { const long long t1_ = clock64(); acc += t1_ - t0_; t0_ = t1_; }
    t0_ = clock64();
#else
#define B2A_T(acc)
#endif
    while (sm.run) {
        const EpochParams p = sm.p;
        if (service) {
            // the next block starts at pos + blksize whatever the loops decide: stage it now, under this epoch's correlation
            // (the other buffer was last read before the barrier that ended the previous iteration)
            if (lane == 0) pending[buf ^ 1] = b2a_issue_tile(g, sm, p.pos + p.blksize, buf ^ 1);
            if (sm.tileBytes[buf] > 0) {   // consumed by the compute warps before the barrier below
                phase[buf] ^= 1;
                pending[buf] = 0;
            }
            if (lead && sm.eDone >= 0 && !sm.outDone) b2a_write_outputs(g, sm, c, out, cap);   // epoch e-1, under epoch e
        } else {
            if (sm.tileBytes[buf] > 0) {
                mbar_wait(&sm.full[buf], phase[buf]);
                phase[buf] ^= 1;
            }
            B2A_T(tW)
            b2a_correlate(g, sm, p, guard, nFast, nExact, buf);
        }
        __syncthreads();   // every warp is done with the tile and the table; warp sums are visible; epoch e-1 is written out
        B2A_T(tC)
        // a C/N0 interval ends with this epoch and lock-loss handling is on: the lock detector decides whether there is
        // a next epoch, so CTA 0 writes this epoch's outputs and the interval's C/N0 first and tells the others
        const bool lockStep = g.lockPLD > 0.0 && g.cnoInterval > 0 && (eCur + 1) % g.cnoInterval == 0;
        if (warp < 3) {
            if (warp == 0) {
                // this CTA's Q8 sums -> slot [rk] of every CTA of the cluster (its own included): one bulk copy through
                // distributed shared memory per peer, completing on the peer's exchange barrier
                const long long t = b2a_cta_sum(sm, lane);
                if (CS > 1) {
                    if (lane < 12) sm.mine[par][lane] = t;
                    b2a_fence_async_smem();   // the sums were written through the generic proxy, the copies read them through the async one
                    __syncwarp();
                    if (lane == 0) b2a_xbar_expect(&sm.xbar[par], 96u * CS);
                    if ((unsigned)lane < CS) b2a_bulk_to_peer(sm.part[par][rk], sm.mine[par], 96u, &sm.xbar[par], (unsigned)lane);
                } else if (lane < 12) {
                    sm.part[par][0][lane] = t;
                }
            }
            if (CS > 1) {
                b2a_xbar_wait(&sm.xbar[par], xph[par]);   // everybody's sums have arrived
                xph[par] ^= 1;
            } else {
                b2a_closers_sync();
            }
            B2A_T(tX)
            if (warp == 0) {
                b2a_collect(g, sm, sm.part[par], (int)CS);
                b2a_disc_dll(g, sm, sm.part[par], (int)CS);
            } else if (warp == 1) {
                b2a_disc_pll(g, sm, sm.part[par], (int)CS);
            } else {
                b2a_disc_carr(g, sm, p);
            }
            b2a_closers_sync();
        }
        if (warp == 0) {
            B2A_T(tD)
            if (lane == 0) {
                close_nco(g, sm.sums, p, g.cc[c].chCodeFreq, sm.st, sm.aux, sm.pre);
                // what the next table needs of the next NCO, with next_params' own expressions: the builder warps start
                // on it now, while this thread works out blksize, the window check and whether there is a next epoch
                sm.pNext.rem = sm.st.remCodePhase;
                sm.pNext.step = sm.st.codeFreq / g.fs;
                sm.pNext.carrFreq = sm.st.carrFreq;
                sm.pNext.remCarr = sm.st.remCarrPhase;
            }
        }
        if (warp <= 9) b2a_builders_sync();   // sm.pNext is visible to the builder warps
        if (warp == 0) {
            B2A_T(tN)
            if (lane == 0) {
                sm.pDone = p;
                const int e = sm.eDone = sm.st.epoch;
                sm.st.epoch += 1;
                ++done;
                sm.outDone = 0;
                if (lockStep && lead) {
                    close_out(g, sm.sums, p, sm.aux, sm.outv);
                    for (int f = 0; f < kNFields; ++f)
                        if (field_written(g, f)) out[(size_t)f * cap + e] = sm.outv[f];
                    __threadfence();
                    close_cno(g, c, e, sm.st);
                    for (unsigned r = 1; r < CS; ++r) {
                        b2a_st_peer(&sm.bcast[0], r, sm.st.lockLost);
                        b2a_st_peer(&sm.bcast[1], r, (long long)sm.st.lowLock);
                        b2a_xbar_arrive_peer(&sm.bbar, r);
                    }
                }
                if (lockStep) sm.outDone = 1;
                if (lockStep && !lead) {
                    b2a_xbar_wait(&sm.bbar, bph);
                    sm.st.lockLost = sm.bcast[0];
                    sm.st.lowLock = (int)sm.bcast[1];
                }
            }
            if (lockStep) bph ^= 1;
        }
        if (warp == 0 && lane == 0) {
            EpochParams np;
            const bool okp = next_params(g, sm.st, np), lim = sm.st.epoch < g.epochLimit;
            if (lead && !okp && lim && sm.st.lockLost == 0) out[(size_t)F_ABS * cap + sm.st.epoch] = (double)sm.st.pos;
            const int run = okp && lim && done < g.maxEpochs;
            if (run) sm.p = np;
            sm.run = run;
        } else if (warp >= 1 && warp <= 9) {
            // the next epoch's table from sm.pNext, spread over nine warps, while lane 0 of warp 0 finishes the NCO
            // (thresholds / order / masks: warp 1; carrier rotation: warps 2-5; rank index: warps 6-9)
            if (warp == 1) fastb_build_tab_warp(&sm.tab, sm.pNext, g.fs, sm.scratch, false, false, true);
            else if (warp <= 5) fastb_build_rot(&sm.tab, fastb_dphi(sm.pNext, g.fs), tid - 64, 128);
            else fastb_build_bins_warp(&sm.tab, sm.pNext, sm.binScratch[warp - 6], 32 * (warp - 6));
        }
        __syncthreads();   // the next epoch's NCO (sm.p), its table, sm.run and the closure's by-products are visible
        B2A_T(tL)
        buf ^= 1;
        par ^= 1;
        ++eCur;
    }
    // the last epoch's outputs (nothing overwrites the closure's by-products any more)
    if (service && lead && sm.eDone >= 0 && !sm.outDone) b2a_write_outputs(g, sm, c, out, cap);
    if (service && lane == 0) {
        for (int b = 0; b < 2; ++b)
            if (pending[b]) mbar_wait(&sm.full[b], phase[b]);   // never leave with a bulk copy in flight
    }
    __syncthreads();   // close_cno of the last interval has updated sm.st
#ifdef BDS_FW_DEV
    if (tid == 0 && lead && g.counters) {
        atomicAdd(g.counters + 4, (unsigned long long)tW);
        atomicAdd(g.counters + 5, (unsigned long long)tC);
        atomicAdd(g.counters + 6, (unsigned long long)tL);
        atomicAdd(g.counters + 7, (unsigned long long)tT);
        atomicAdd(g.counters + 8, (unsigned long long)done);
        atomicAdd(g.counters + 9, (unsigned long long)tX);
        atomicAdd(g.counters + 10, (unsigned long long)tD);
        atomicAdd(g.counters + 11, (unsigned long long)tN);
    }
#endif
#undef B2A_T
    if (CS > 1) b2a_cluster_sync();   // nobody leaves while a peer could still store into its shared memory
    if (tid == 0 && lead) g.st[c] = sm.st;
    __syncwarp();
    if (g.counters && !service) {
        const unsigned tf = __reduce_add_sync(0xffffffffu, nFast), te = __reduce_add_sync(0xffffffffu, nExact);
        if (lane == 0) {
            if (tf) atomicAdd(g.counters + 0, (unsigned long long)tf);
            if (te) atomicAdd(g.counters + 1, (unsigned long long)te);
        }
    }
}

// Open loop (teacher forced): one CTA per channel-epoch, sums[ce][18] written directly.
__global__ void __launch_bounds__(kB2aThreads, 1) trk_b2a_unit_open_kernel(TrkDev g, const EpochParams* params, int nEpochs,
                                                                         double* sums) {
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    B2aSmem& sm = *reinterpret_cast<B2aSmem*>(dyn_smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ce = blockIdx.x, c = ce / nEpochs;
    b2a_load_bits(g, sm, c);
    if (tid == 0) {
        mbar_init(&sm.full[0], 1);
        b2a_plan_share(g, sm, c, 0, 1);
        sm.p = params[ce];
        b2a_issue_tile(g, sm, sm.p.pos, 0);
    }
    __syncthreads();
    if (warp == 0) fastb_build_tab_warp(&sm.tab, sm.p, g.fs, sm.scratch);
    __syncthreads();
    const EpochParams p = sm.p;
    if (sm.tileBytes[0] > 0) mbar_wait(&sm.full[0], 0);
    unsigned nFast = 0, nExact = 0;
    b2a_correlate(g, sm, p, g.pad ? (1u << 24) : kFastGuard, nFast, nExact, 0);
    __syncthreads();
    if (warp == 0) {
        const long long t = b2a_cta_sum(sm, lane);
        if (lane < 12) sm.part[0][0][lane] = t;
        __syncwarp();
        b2a_collect(g, sm, sm.part[0], 1);
        if (lane < kNSum) sums[(size_t)ce * kNSum + lane] = sm.sums[lane];
    }
    __syncwarp();
    if (g.counters) {
        const unsigned tf = __reduce_add_sync(0xffffffffu, nFast), te = __reduce_add_sync(0xffffffffu, nExact);
        if (lane == 0) {
            if (tf) atomicAdd(g.counters + 0, (unsigned long long)tf);
            if (te) atomicAdd(g.counters + 1, (unsigned long long)te);
        }
    }
}

}  // namespace bds

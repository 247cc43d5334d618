// Warp-specialised persistent kernel for the chip-synchronous B1C wide-band correlator.
// (included from bds_track.cu after close_epoch / next_params)
//
// One CTA of 16 warps per SM:
//   warp 0      producer : walks this CTA's work items, waits until the item's channel-epoch has
//                          been published (acquire load), then stages everything a pass needs —
//                          the IF tile, the per-epoch tables, the packed codes and the NCO params
//                          — into one of kFwStages shared-memory stages with TMA bulk copies that
//                          complete on the stage's mbarrier.  It runs ahead of the compute warps,
//                          so global-memory latency is off their critical path.
//   warp 1      closer   : receives the per-warp sums of a finished slice, stores the slice
//                          partial, bumps the channel's arrival counter and — if it was the last
//                          slice — reduces all partials in a fixed order, closes the PLL/DLL in
//                          fp64 (close_epoch), builds the next epoch's tables and publishes them
//                          with a release store.
//   warps 2..15 compute  : one chip per thread and pass (fast_chip); warp sums via REDUX.
#pragma once

namespace bds {

constexpr int kFwThreads = 512;
constexpr int kFwCompute = 14;                 // compute warps
constexpr int kFwChips = kFwCompute * 32;      // chips per pass
constexpr int kFwStages = 3;
constexpr int kFwTile = 44032;                 // >= 448 chips * 97.2 samples + margins, multiple of 128
constexpr int kFwBitsBytes = 2 * kPackedWordsDev * 4;

struct FwUnit {
    int c, e, sl, seq;     // channel (or -1: terminate), epoch, slice, task sequence number in this CTA
    int c0, cEnd;          // chip range of this pass
    int first, last;       // first / last pass of the task
    int ce, pad0;          // open loop: channel-epoch index
    long long tileBase;    // window byte offset of tile[0]
    long long B0;          // window byte offset of the block start
};

struct __align__(128) FwStage {
    FastTab tab;
    uint32_t bits[2][kPackedWordsDev];
    EpochParams p;
    FwUnit u;
    __align__(128) unsigned char tile[kFwTile + 256];
};

struct __align__(128) FwSmem {
    unsigned long long full[kFwStages], empty[kFwStages], resFull[2], resEmpty[2];
    int res[2][kFwCompute][kNSum];
    int resTask[2][4];
    double sums[kNSum];
    unsigned scratch[128];
    EpochParams np;
    FwStage st[kFwStages];
};

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ bool mbar_test(unsigned long long* bar, unsigned phase) {
    unsigned ok;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
    return ok != 0;
}

__host__ __device__ inline int fw_chips_per_slice(int S) {
    int cps = (10230 + S - 1) / S;
    return ((cps + kFwChips - 1) / kFwChips) * kFwChips;
}

// ---- closure by one warp ----------------------------------------------------------------------
__device__ void fw_closure(const TrkDev& g, FwSmem& sm, int c, int e) {
    const int lane = threadIdx.x & 31;
    if (lane < kNSum) {  // fixed summation order over the S slices (deterministic)
        const double* part = g.partial + (size_t)c * g.S * kNSum;
        double a = 0;
        for (int s = 0; s < g.S; ++s) a += __ldcg(part + (size_t)s * kNSum + lane);
        sm.sums[lane] = a;
    }
    __syncwarp();
    int ok = 0;
    if (lane == 0) {
        int npOk;
        close_epoch(g, c, e, sm.sums, sm.np, npOk);
        ok = npOk;
    }
    ok = __shfl_sync(0xffffffffu, ok, 0);
    __syncwarp();
    if (ok) {
        fast_build_tab_warp(g.fastTab + (size_t)c * 2 + ((e + 1) & 1), sm.np, g.fs, sm.scratch);
        if (lane == 0) store_cg(g.params + c * 2 + ((e + 1) & 1), sm.np);
    }
    __threadfence();
    __syncwarp();
    if (lane == 0) {
        __threadfence();
        if (ok) st_release(g.ready + c, e + 1);
        else st_release(g.stop + c, e + 1);
    }
}

// ---- the kernel ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFwThreads, 1) trk_fw_kernel(TrkDev g) {
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    FwSmem& sm = *reinterpret_cast<FwSmem*>(dyn_smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kFwStages; ++s) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], kFwCompute);
        }
        for (int r = 0; r < 2; ++r) {
            mbar_init(&sm.resFull[r], kFwCompute);
            mbar_init(&sm.resEmpty[r], 1);
        }
    }
    __syncthreads();
    const bool openLoop = g.olParams != nullptr;
    const long long perRound = openLoop ? (long long)g.S : (long long)g.nAct * g.S;
    const long long total = openLoop ? (long long)g.olCount * g.S : perRound * g.maxEpochs;
    const int cps = fw_chips_per_slice(g.S);

    if (warp == 0) {
        // ================================ producer ================================
        if (lane != 0) return;
        unsigned u = 0;
        int seq = 0;
        for (long long t = blockIdx.x; t < total; t += gridDim.x) {
            int c, e, sl, ce = 0;
            const EpochParams* gp;
            const FastTab* gt;
            if (openLoop) {
                ce = (int)(t / g.S);
                sl = (int)(t - (long long)ce * g.S);
                c = ce / g.olEpochs;
                e = ce - c * g.olEpochs;
                gp = g.olParams + ce;
                gt = g.fastTab + ce;
            } else {
                const int i = (int)(t / perRound);
                const int idx = (int)(t - (long long)i * perRound);
                c = g.act[idx / g.S];
                sl = idx % g.S;
                e = g.cc[c].pad + i;
                bool go = false;
                while (true) {
                    if (ld_acquire(g.stop + c) <= e) break;
                    if (ld_acquire(g.ready + c) >= e) {
                        go = true;
                        break;
                    }
                    __nanosleep(40);
                }
                if (!go) continue;
                gp = g.params + c * 2 + (e & 1);
                gt = g.fastTab + (size_t)c * 2 + (e & 1);
            }
            const int cLo = sl * cps, cHi = min(10230, cLo + cps);  // host guarantees S = ceil(10230 / cps): never empty
            const EpochParams p = load_cg(gp);
            const double u0 = __ldcg(&gt->u0), Ss = __ldcg(&gt->S);
            const long long B0 = p.pos - g.winFirst;
            int c0 = cLo;
            do {
                const int cEnd = min(c0 + kFwChips, cHi);
                const int stage = u % kFwStages;
                mbar_wait(&sm.empty[stage], ((u / kFwStages) & 1) ^ 1);
                FwStage& st = sm.st[stage];
                long long na = 0, nb = 0;
                if (cEnd > c0) {
                    double qa = ((double)(12 * c0) - u0) * Ss, qb = ((double)(12 * cEnd) - u0) * Ss;
                    na = (long long)floor(qa) - 2;
                    nb = (long long)floor(qb) + 4;
                    if (na < 0) na = 0;
                    if (nb > p.blksize) nb = p.blksize;
                    if (nb < na) nb = na;
                }
                const long long gA = (B0 + na) & ~15LL;
                long long gE = (B0 + nb + 15) & ~15LL;
                if (gE - gA > kFwTile) gE = gA + kFwTile;
                const unsigned bytes = (unsigned)(gE - gA);
                FwUnit d;
                d.c = c; d.e = e; d.sl = sl; d.seq = seq;
                d.c0 = c0; d.cEnd = cEnd;
                d.first = (c0 == cLo); d.last = (cEnd >= cHi);
                d.ce = ce; d.pad0 = 0;
                d.tileBase = gA; d.B0 = B0;
                st.u = d;
                mbar_expect_tx(&sm.full[stage], bytes + (unsigned)sizeof(FastTab) + kFwBitsBytes + (unsigned)sizeof(EpochParams));
                if (bytes) tma_bulk(st.tile, g.x + gA, bytes, &sm.full[stage]);
                tma_bulk(&st.tab, gt, (unsigned)sizeof(FastTab), &sm.full[stage]);
                tma_bulk(st.bits, g.codeBits + (size_t)c * 2 * kPackedWordsDev, kFwBitsBytes, &sm.full[stage]);
                tma_bulk(&st.p, gp, (unsigned)sizeof(EpochParams), &sm.full[stage]);
                ++u;
                c0 = cEnd;
            } while (c0 < cHi);
            ++seq;
        }
        // terminate
        const int stage = u % kFwStages;
        mbar_wait(&sm.empty[stage], ((u / kFwStages) & 1) ^ 1);
        sm.st[stage].u.c = -1;
        sm.st[stage].u.seq = seq;
        mbar_arrive(&sm.full[stage]);
    } else if (warp == 1) {
        // ================================ closer ================================
        for (int k = 0;; ++k) {
            const int rs = k & 1;
            mbar_wait(&sm.resFull[rs], (k >> 1) & 1);
            const int c = sm.resTask[rs][0], e = sm.resTask[rs][1], sl = sm.resTask[rs][2], ce = sm.resTask[rs][3];
            double v = 0;
            if (lane < kNSum && c >= 0) {
                long long a = 0;
#pragma unroll
                for (int w = 0; w < kFwCompute; ++w) a += sm.res[rs][w][lane];
                v = (double)a * (1.0 / 256.0);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.resEmpty[rs]);
            if (c < 0) break;
            if (openLoop) {
                if (lane < kNSum) g.partial[((size_t)ce * g.S + sl) * kNSum + lane] = v;
                continue;
            }
            if (lane < kNSum) g.partial[((size_t)c * g.S + sl) * kNSum + lane] = v;
            __syncwarp();
            int last = 0;
            if (lane == 0) {
                __threadfence();
                int old = atomicAdd(g.count + c, 1);
                last = (old == g.S - 1);
                if (last) {
                    g.count[c] = 0;
                    __threadfence();
                }
            }
            last = __shfl_sync(0xffffffffu, last, 0);
            if (last) fw_closure(g, sm, c, e);
        }
    } else {
        // ================================ compute ================================
        const int cw = warp - 2;
        float acc[kNSum];
#pragma unroll
        for (int i = 0; i < kNSum; ++i) acc[i] = 0.f;
        const unsigned guard = g.pad ? (1u << 24) : kFastGuard;  // g.pad: test hook, widens the guard band
        for (unsigned u = 0;; ++u) {
            const int stage = u % kFwStages;
            mbar_wait(&sm.full[stage], (u / kFwStages) & 1);
            const FwStage& st = sm.st[stage];
            const FwUnit d = st.u;
            if (d.c < 0) {  // terminate: forward to the closer through the result channel
                const int rs = d.seq & 1;
                if (lane == 0) {
                    mbar_wait(&sm.resEmpty[rs], ((d.seq >> 1) & 1) ^ 1);
                    if (cw == 0) sm.resTask[rs][0] = -1;
                    mbar_arrive(&sm.resFull[rs]);
                }
                break;
            }
            const int c = d.c0 + cw * 32 + lane;
            const bool active = c < d.cEnd;
            bool exact = false;
            if (d.first && d.sl == 0 && cw == 0 && lane == 0 && st.p.rem == 0.0) {
                // the t = 0 sample takes the previous period's last chip (SURVEY quirk i)
                ExactCtx ex;
                make_exact_ctx(st.p, g.d, g.fs, ex);
                float tmp[kNSum];
#pragma unroll
                for (int i = 0; i < kNSum; ++i) tmp[i] = 0.f;
                fast_exact_range(ex, g.x + d.B0, st.bits[0], st.bits[1], 0, 0, -100, 0, tmp);
#pragma unroll
                for (int i = 0; i < kNSum; ++i) acc[i] += tmp[i];
            }
            if (active)
                exact = fast_chip(st.tab, st.p, st.bits[0], st.bits[1], st.tile, d.tileBase, d.B0, g.x + d.B0, g.d, g.fs,
                                  c, guard, acc);
            if (g.counters) {
                unsigned bf = __ballot_sync(0xffffffffu, active && !exact), be = __ballot_sync(0xffffffffu, active && exact);
                if (lane == 0) {
                    if (bf) atomicAdd(g.counters + 0, (unsigned long long)__popc(bf));
                    if (be) atomicAdd(g.counters + 1, (unsigned long long)__popc(be));
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty[stage]);
            if (d.last) {
                const int rs = d.seq & 1;
                int mine = 0;
#pragma unroll
                for (int i = 0; i < kNSum; ++i) {
                    int s = __reduce_add_sync(0xffffffffu, __float2int_rn(acc[i] * 256.f));
                    if (lane == i) mine = s;
                    acc[i] = 0.f;
                }
                mbar_wait(&sm.resEmpty[rs], ((d.seq >> 1) & 1) ^ 1);
                if (lane < kNSum) sm.res[rs][cw][lane] = mine;
                if (cw == 0 && lane == 0) {
                    sm.resTask[rs][0] = d.c;
                    sm.resTask[rs][1] = d.e;
                    sm.resTask[rs][2] = d.sl;
                    sm.resTask[rs][3] = d.ce;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.resFull[rs]);
            }
        }
    }
}

// Builds the per-epoch tables for an array of params (open loop) — one warp per entry.
__global__ void fw_tab_kernel(const EpochParams* params, int n, double fs, FastTab* tabs) {
    __shared__ unsigned scratch[4][128];
    const int w = threadIdx.x >> 5;
    const int i = blockIdx.x * 4 + w;
    if (i >= n) return;
    fast_build_tab_warp(tabs + i, params[i], fs, scratch[w]);
}

// First params + tables of every channel for the current window (run start): one warp per channel.
__global__ void fw_prepare_kernel(TrkDev g) {
    __shared__ unsigned scratch[4][128];
    __shared__ EpochParams nps[4];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * 4 + w;
    if (c >= g.nCh) return;
    if (!g.cc[c].active) {
        if (lane == 0) {
            g.count[c] = 0;
            g.stop[c] = 0;
            g.ready[c] = -1;
        }
        return;
    }
    int ok = 0, e = 0;
    if (lane == 0) {
        g.count[c] = 0;
        ChanState st = g.st[c];
        g.cc[c].pad = st.epoch;
        e = st.epoch;
        ok = next_params(g, st, nps[w]) && st.epoch < g.capacity;
        if (!ok && st.epoch < g.capacity)
            g.out[((size_t)c * kNFields + F_ABS) * g.capacity + st.epoch] = (double)st.pos;
        g.ready[c] = st.epoch - 1;
        g.stop[c] = INT_MAX;
    }
    ok = __shfl_sync(0xffffffffu, ok, 0);
    e = __shfl_sync(0xffffffffu, e, 0);
    __syncwarp();
    if (ok) {
        fast_build_tab_warp(g.fastTab + (size_t)c * 2 + (e & 1), nps[w], g.fs, scratch[w]);
        if (lane == 0) store_cg(g.params + c * 2 + (e & 1), nps[w]);
    }
    __threadfence();
    __syncwarp();
    if (lane == 0) {
        if (ok) g.ready[c] = e;
        else g.stop[c] = e;
    }
}

}  // namespace bds

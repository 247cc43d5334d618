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

#ifndef BDS_FW_COMPUTE_WARPS
#define BDS_FW_COMPUTE_WARPS 14
#endif
constexpr int kFwCompute = BDS_FW_COMPUTE_WARPS;   // compute warps
constexpr int kFwThreads = (kFwCompute + 2) * 32;
constexpr int kFwChips = kFwCompute * 32;      // chips per pass
constexpr int kFwStages = 3;   // compiled-in maximum; g.stages (2..3) are used
constexpr int kFwTile = ((kFwChips * 98 + 512 + 127) / 128) * 128;   // chips * 97.2 samples + margins
constexpr int kFwBitsBytes = 2 * kPackedWordsDev * 4;

struct FwUnit {
    int c, e, sl, seq;     // channel (or -1: terminate), epoch, slice, task sequence number in this CTA
    int c0, cEnd;          // chip range of this pass
    int first, last;       // first / last pass of the task
    int ce, ticket;        // open loop: channel-epoch index; queue ticket (developer tracing)
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
    FwStage st[kFwStages];
};

// scratch of one loop-closing warp (closer CTAs overlay an array of these on the dynamic smem)
struct __align__(16) FwCloseScratch {
    double sums[kNSum];
    double pre[8];
    unsigned scratch[128];
    EpochParams np;
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

// ---- ready-task queue -----------------------------------------------------------------------------
// payload: channel (7 bits) | slice (6 bits) << 7 | epoch (19 bits) << 13 ; 0xffffffff = terminate
__device__ __forceinline__ unsigned fw_payload(int c, int sl, int e) { return (unsigned)c | ((unsigned)sl << 7) | ((unsigned)e << 13); }
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// push the S slices of (c, e); caller has fenced its prior writes (params, tables)
__device__ void fw_push_slices(const TrkDev& g, int c, int e, int lane) {
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(g.qctl + 1, (unsigned)g.S);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (int s = lane; s < g.S; s += 32)
        st_release_u64(g.queue + ((base + s) & g.qMask), ((unsigned long long)(base + s + 1) << 32) | fw_payload(c, s, e));
}
__device__ void fw_push_terminate(const TrkDev& g, int n, int lane) {
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(g.qctl + 1, (unsigned)n);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (int s = lane; s < n; s += 32)
        st_release_u64(g.queue + ((base + s) & g.qMask), ((unsigned long long)(base + s + 1) << 32) | 0xffffffffull);
}
// a channel has no further epoch in this launch: the last one to end shuts the grid down
__device__ void fw_channel_done(const TrkDev& g, int lane, int nCtas) {
    unsigned left = 0;
    if (lane == 0) left = atomicSub(g.qctl + 2, 1u) - 1u;
    left = __shfl_sync(0xffffffffu, left, 0);
    if (left == 0) fw_push_terminate(g, nCtas, lane);
}

__device__ __forceinline__ unsigned long long gtimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// ---- closure by one warp ----------------------------------------------------------------------
// returns true if another epoch of the channel was published
__device__ bool fw_closure(const TrkDev& g, FwCloseScratch& sm, int c, int e) {
    const int lane = threadIdx.x & 31;
    const unsigned long long tIn = gtimer_ns();
    const long long tcIn = clock64();
    if (lane == 0 && g.counters && g.pubTime) {  // developer timing: publish -> all slices done
        unsigned long long t0 = g.pubTime[c];
        if (t0) atomicAdd(g.counters + 12, tIn - t0);
    }
    if (lane < kNSum) {  // fixed summation order over the S slices (deterministic)
        const double* part = g.partial + (size_t)c * g.S * kNSum;
        double a = 0;
        for (int s = 0; s < g.S; ++s) a += __ldcg(part + (size_t)s * kNSum + lane);
        sm.sums[lane] = a;
    }
    __syncwarp();
    const long long tc0 = clock64();
    const bool par = (g.tune & 2) != 0;
    if (par) {   // expensive scalar pieces in parallel lanes, then the sequential filter update on lane 0
        const double v = close_pre_wb(g, c, e, sm.sums, lane);
        if (lane < 5) sm.pre[lane] = v;
    }
    __syncwarp();
    int ok = 0;
    if (lane == 0) {
        int npOk;
        close_epoch(g, c, e, sm.sums, sm.np, npOk, par ? sm.pre : nullptr);
        ok = npOk;
    }
    ok = __shfl_sync(0xffffffffu, ok, 0);
    __syncwarp();
    const long long tc1 = clock64();
    const bool more = ok && (e + 1 - g.cc[c].pad) < g.maxEpochs;  // another epoch of this channel in this launch?
    if (more) {
        fast_build_tab_warp(g.fastTab + (size_t)c * 2 + ((e + 1) & 1), sm.np, g.fs, sm.scratch);
        if (lane == 0) store_cg(g.params + c * 2 + ((e + 1) & 1), sm.np);
    }
    const long long tc2 = clock64();
    __threadfence();
    __syncwarp();
    if (lane == 0) {
        if (ok) st_release(g.ready + c, e + 1);
        else st_release(g.stop + c, e + 1);
    }
    if (lane == 0 && g.counters) {
        atomicAdd(g.counters + 14, (unsigned long long)(tc0 - tcIn));
        atomicAdd(g.counters + 15, (unsigned long long)(tc1 - tc0));
        atomicAdd(g.counters + 16, (unsigned long long)(tc2 - tc1));
        atomicAdd(g.counters + 17, (unsigned long long)(clock64() - tc2));
    }
    if (lane == 0 && g.pubTime) {
        const unsigned long long tOut = gtimer_ns();
        g.pubTime[c] = tOut;
        if (g.counters) {
            atomicAdd(g.counters + 13, tOut - tIn);
            atomicAdd(g.counters + 11, 1ull);
        }
    }
    if (more) fw_push_slices(g, c, e + 1, lane);
    else fw_channel_done(g, lane, g.nCompute);
    return more;
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
    const unsigned nst = (unsigned)g.stages;

    if (!openLoop && (int)blockIdx.x >= g.nCompute) {
        // ================================ closer CTA ================================
        // Every warp owns the channels c with (index in the active list) % (closer warps) == its id and
        // polls their slice-arrival counters; when all S slices of an epoch have arrived it closes the
        // loops (fp64), builds the next epoch's tables, publishes and queues the next slices.
        FwCloseScratch* cs = reinterpret_cast<FwCloseScratch*>(dyn_smem) + warp;
        const int nCw = ((int)gridDim.x - g.nCompute) * (kFwThreads / 32);
        const int me = ((int)blockIdx.x - g.nCompute) * (kFwThreads / 32) + warp;
        int chan[8], ep[8], n = 0;
        for (int i = me; i < g.nAct && n < 8; i += nCw) {
            const int c = g.act[i];
            // channels that could not start were already retired by the prepare kernel (stop <= first epoch)
            if (g.stop[c] > g.cc[c].pad) {
                chan[n] = c;
                ep[n] = g.cc[c].pad;
                ++n;
            }
        }
        while (n > 0) {
            bool fired = false;
            for (int i = 0; i < n; ++i) {
                int v = 0;
                if (lane == 0) v = ld_acquire(g.count + chan[i]);
                v = __shfl_sync(0xffffffffu, v, 0);
                if (v != g.S) continue;
                fired = true;
                if (lane == 0) g.count[chan[i]] = 0;
                __syncwarp();
                const bool more = fw_closure(g, *cs, chan[i], ep[i]);
                if (more) {
                    ++ep[i];
                } else {
                    chan[i] = chan[n - 1];
                    ep[i] = ep[n - 1];
                    --n;
                    --i;
                }
            }
            if (!fired) __nanosleep(100);
        }
        return;
    }
    if (warp == 0) {
        // ================================ producer ================================
        if (lane != 0) return;
        unsigned u = 0;
        int seq = 0;
        long long tQueue = 0, tEmpty = 0, tStart = clock64();
        unsigned curTicket = 0;
        for (long long t = blockIdx.x;; t += gridDim.x) {
            int c, e, sl, ce = 0;
            const EpochParams* gp;
            const FastTab* gt;
            if (openLoop) {
                if (t >= total) break;
                ce = (int)(t / g.S);
                sl = (int)(t - (long long)ce * g.S);
                c = ce / g.olEpochs;
                e = ce - c * g.olEpochs;
                gp = g.olParams + ce;
                gt = g.fastTab + ce;
            } else {
                // take work only when a stage is free for it (tasks are scarce while channels sit in loop
                // closure, so a CTA must not hoard them), then pop the next ready (channel, epoch, slice)
                if (g.ahead >= 0 && (int)u > g.ahead) {  // unit u-1-ahead released => at most `ahead` units pending
                    const unsigned v = u - 1u - (unsigned)g.ahead;
                    long long t1 = clock64();
                    mbar_wait(&sm.empty[v % nst], (v / nst) & 1);
                    tEmpty += clock64() - t1;
                }
                long long t0 = clock64();
                const unsigned ticket = atomicAdd(g.qctl + 0, 1u);
                unsigned long long ent;
                while (((ent = ld_acquire_u64(g.queue + (ticket & g.qMask))) >> 32) != (unsigned long long)ticket + 1ull)
                    __nanosleep(64);
                tQueue += clock64() - t0;
                curTicket = ticket;
                if (g.trace && ticket < g.traceCap) {
                    g.trace[(size_t)ticket * 8 + 0] = ((unsigned long long)blockIdx.x << 32) | (unsigned)ent;
                    g.trace[(size_t)ticket * 8 + 1] = gtimer_ns();
                }
                const unsigned pl = (unsigned)ent;
                if (pl == 0xffffffffu) break;
                c = (int)(pl & 127u);
                sl = (int)((pl >> 7) & 63u);
                e = (int)(pl >> 13);
                gp = g.params + c * 2 + (e & 1);
                gt = g.fastTab + (size_t)c * 2 + (e & 1);
            }
            const int cLo = sl * cps, cHi = min(10230, cLo + cps);  // host guarantees S = ceil(10230 / cps): never empty
            const EpochParams p = load_cg(gp);
            double u0 = 12.0 * p.rem, Ss = 1.0 / (12.0 * p.step);  // == tab.u0, tab.S (same expressions)
            if (g.tune & 4) {
                u0 = __ldcg(&gt->u0);
                Ss = __ldcg(&gt->S);
            }
            const long long B0 = p.pos - g.winFirst;
            int c0 = cLo;
            do {
                const int cEnd = min(c0 + kFwChips, cHi);
                const int stage = u % nst;
                long long t1 = clock64();
                mbar_wait(&sm.empty[stage], ((u / nst) & 1) ^ 1);
                tEmpty += clock64() - t1;
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
                d.ce = ce; d.ticket = (int)curTicket;
                d.tileBase = gA; d.B0 = B0;
                st.u = d;
                mbar_expect_tx(&sm.full[stage], bytes + (unsigned)sizeof(FastTab) + kFwBitsBytes + (unsigned)sizeof(EpochParams));
                if (bytes) tma_bulk(st.tile, g.x + gA, bytes, &sm.full[stage]);
                tma_bulk(&st.tab, gt, (unsigned)sizeof(FastTab), &sm.full[stage]);
                tma_bulk(st.bits, g.codeBits + (size_t)c * 2 * kPackedWordsDev, kFwBitsBytes, &sm.full[stage]);
                tma_bulk(&st.p, gp, (unsigned)sizeof(EpochParams), &sm.full[stage]);
                if (g.trace && d.first && curTicket < g.traceCap && !openLoop) g.trace[(size_t)curTicket * 8 + 2] = gtimer_ns();
                ++u;
                c0 = cEnd;
            } while (c0 < cHi);
            ++seq;
        }
        // terminate
        const int stage = u % nst;
        mbar_wait(&sm.empty[stage], ((u / nst) & 1) ^ 1);
        sm.st[stage].u.c = -1;
        sm.st[stage].u.seq = seq;
        mbar_arrive(&sm.full[stage]);
        if (g.counters) {
            atomicAdd(g.counters + 4, (unsigned long long)tQueue);
            atomicAdd(g.counters + 5, (unsigned long long)tEmpty);
            atomicAdd(g.counters + 6, (unsigned long long)(clock64() - tStart));
        }
    } else if (warp == 1) {
        // ================================ epilogue warp ================================
        long long tClose = 0, tEpi = 0;
        int nClose = 0;
        for (int k = 0;; ++k) {
            const int rs = k & 1;
            if (lane == 0)
                while (!mbar_test(&sm.resFull[rs], (k >> 1) & 1)) __nanosleep(100);
            long long t0 = clock64();
            __syncwarp();
            const int c = sm.resTask[rs][0], sl = sm.resTask[rs][2], ce = sm.resTask[rs][3];
            double v = 0;
            if (lane < kNSum && c >= 0) {
                long long a = 0;
#pragma unroll
                for (int w = 0; w < kFwCompute; ++w) a += sm.res[rs][w][lane];
                v = (double)a * (1.0 / 256.0);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.resEmpty[rs]);
            if (c < 0) {
                if (lane == 0 && g.counters) {
                    atomicAdd(g.counters + 9, (unsigned long long)tClose);
                    atomicAdd(g.counters + 10, (unsigned long long)tEpi);
                    atomicAdd(g.counters + 11, (unsigned long long)nClose);
                }
                break;
            }
            if (openLoop) {
                if (lane < kNSum) g.partial[((size_t)ce * g.S + sl) * kNSum + lane] = v;
                continue;
            }
            if (lane < kNSum) {
                g.partial[((size_t)c * g.S + sl) * kNSum + lane] = v;
                __threadfence();
            }
            __syncwarp();
            if (lane == 0) {   // count the slice; the channel's closer warp (closer CTA) polls this counter
                asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(g.count + c) : "memory");
                tEpi += clock64() - t0;
                if (g.trace && (unsigned)ce < g.traceCap) g.trace[(size_t)ce * 8 + 5] = gtimer_ns();
            }
        }
    } else {
        // ================================ compute ================================
        const int cw = warp - 2;
        float acc[kNSum];
#pragma unroll
        for (int i = 0; i < kNSum; ++i) acc[i] = 0.f;
        const unsigned guard = g.pad ? (1u << 24) : kFastGuard;  // g.pad: test hook, widens the guard band
        long long tFull = 0, tRes = 0;
        for (unsigned u = 0;; ++u) {
            const int stage = u % nst;
            long long t0 = clock64();
            mbar_wait(&sm.full[stage], (u / nst) & 1);
            tFull += clock64() - t0;
            const FwStage& st = sm.st[stage];
            const FwUnit d = st.u;
            if (d.c < 0) {  // terminate: forward to the closer through the result channel
                const int rs = d.seq & 1;
                if (lane == 0) {
                    mbar_wait(&sm.resEmpty[rs], ((d.seq >> 1) & 1) ^ 1);
                    if (cw == 0) sm.resTask[rs][0] = -1;
                    mbar_arrive(&sm.resFull[rs]);
                    if (cw == 0 && g.counters) {
                        atomicAdd(g.counters + 7, (unsigned long long)tFull);
                        atomicAdd(g.counters + 8, (unsigned long long)tRes);
                    }
                }
                break;
            }
            const int c = d.c0 + cw * 32 + lane;
            const bool active = c < d.cEnd;
            bool exact = false;
            const bool tr = g.trace && cw == 0 && lane == 0 && (unsigned)d.ticket < g.traceCap && !openLoop;
            if (tr && d.first) g.trace[(size_t)d.ticket * 8 + 3] = gtimer_ns();
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
            if (tr && d.last) g.trace[(size_t)d.ticket * 8 + 4] = gtimer_ns();
            if (d.last) {
                const int rs = d.seq & 1;
                int mine = 0;
#pragma unroll
                for (int i = 0; i < kNSum; ++i) {
                    int s = __reduce_add_sync(0xffffffffu, __float2int_rn(acc[i] * 256.f));
                    if (lane == i) mine = s;
                    acc[i] = 0.f;
                }
                long long t2 = clock64();
                mbar_wait(&sm.resEmpty[rs], ((d.seq >> 1) & 1) ^ 1);
                tRes += clock64() - t2;
                if (lane < kNSum) sm.res[rs][cw][lane] = mine;
                if (cw == 0 && lane == 0) {
                    sm.resTask[rs][0] = d.c;
                    sm.resTask[rs][1] = d.e;
                    sm.resTask[rs][2] = d.sl;
                    sm.resTask[rs][3] = openLoop ? d.ce : d.ticket;
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

// First params + tables of every channel for the current window (run start).  One CTA; warps take the
// channels in turn, publish their first epoch and queue its slices; if nothing can run the grid is
// told to terminate right away.
__global__ void __launch_bounds__(1024) fw_prepare_kernel(TrkDev g, int nCtas) {
    __shared__ unsigned scratch[32][128];
    __shared__ EpochParams nps[32];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    if (threadIdx.x == 0) {
        g.qctl[0] = 0;
        g.qctl[1] = 0;
        g.qctl[2] = (unsigned)g.nCh + 1u;   // every channel + this kernel hold a reference
    }
    __syncthreads();
    for (int c = w; c < g.nCh; c += nw) {
        int ok = 0, e = 0;
        if (lane == 0) {
            g.count[c] = 0;
            if (g.cc[c].active) {
                ChanState st = g.st[c];
                g.cc[c].pad = st.epoch;
                e = st.epoch;
                ok = next_params(g, st, nps[w]) && st.epoch < g.capacity && g.maxEpochs > 0;
                if (!ok && st.epoch < g.capacity)
                    g.out[((size_t)c * kNFields + F_ABS) * g.capacity + st.epoch] = (double)st.pos;
                g.ready[c] = ok ? e : e - 1;
                g.stop[c] = ok ? INT_MAX : e;
            } else {
                g.stop[c] = 0;
                g.ready[c] = -1;
            }
        }
        ok = __shfl_sync(0xffffffffu, ok, 0);
        e = __shfl_sync(0xffffffffu, e, 0);
        __syncwarp();
        if (ok) {
            fast_build_tab_warp(g.fastTab + (size_t)c * 2 + (e & 1), nps[w], g.fs, scratch[w]);
            if (lane == 0) store_cg(g.params + c * 2 + (e & 1), nps[w]);
            __threadfence();
            __syncwarp();
            fw_push_slices(g, c, e, lane);
        } else {
            fw_channel_done(g, lane, nCtas);
        }
    }
    __syncthreads();
    if (w == 0) fw_channel_done(g, lane, nCtas);   // drop the kernel's own reference
}

}  // namespace bds

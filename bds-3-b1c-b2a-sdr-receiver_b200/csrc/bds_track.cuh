// Device-side data layout of a tracking session (see DESIGN.md "HBM layout").
#pragma once
#include "bds_common.cuh"

namespace bds {

constexpr int kNSum = 18;       // {data,p11,p61} x {E,P,L} x {I,Q}
constexpr int kNFieldsLoop = 21;  // trackResults planes written per epoch
constexpr int kNFields = kNFieldsLoop + kNSum;
constexpr int kNCno = 5;        // DataCNo, DataPLD, PilotCNo, PilotPLD, TotalCNo
constexpr int kPackedWordsDev = 320;
constexpr int kTrkThreads = 256;

enum Field {
    F_ABS = 0, F_CODEFREQ, F_CARRFREQ, F_I_P, F_I_E, F_I_L, F_Q_E, F_Q_P, F_Q_L,
    F_PI_P, F_PI_E, F_PI_L, F_PQ_E, F_PQ_P, F_PQ_L, F_DLL, F_DLLF, F_PLL, F_PLLF,
    F_REMCODE, F_REMCARR, F_RAW0
};

// index into the 18 raw sums
__host__ __device__ constexpr int sum_idx(int fam, int epl, int iq) { return fam * 6 + epl * 2 + iq; }
enum { EPL_E = 0, EPL_P = 1, EPL_L = 2 };

struct ChanConst {
    int prn;
    int active;
    int status;
    int pad;
    double chCodeFreq;    // channel.codeFreq
    double acquiredFreq;  // channel.acquiredFreq
    long long startPos;   // skipNumberOfBytes + codePhase - 1
};

struct ChanState {
    double codeFreq, remCodePhase, carrFreq, carrFreqBasis, remCarrPhase;
    double oldCodeNco, oldCodeError, d2CarrError, dCarrError;
    double cnoPrev[3];
    long long pos;        // absolute sample index of the next block
    long long samples;    // samples consumed so far
    int epoch;            // epochs completed
    int lowLock;          // consecutive C/N0 intervals with the lock detector below cfg.lockLossPLD
    long long lockLost;   // 0, or the number of epochs completed when the channel was dropped for loss of lock
};
static_assert(sizeof(ChanState) % 16 == 0, "ChanState must be a 16-byte multiple");

struct EpochParams {
    long long pos;   // absolute 0-based sample index of the block start
    int blksize;
    int pad;
    double rem;      // remCodePhase at block start [chips]
    double step;     // codePhaseStep [chips/sample]
    double carrFreq; // [Hz]
    double remCarr;  // [rad]
};
static_assert(sizeof(EpochParams) % 16 == 0, "EpochParams must be a 16-byte multiple");

struct TrkDev {
    const int8_t* x;      // resident IF window
    long long winFirst;   // absolute index of x[0]
    long long winLen;
    long long winStage;   // bytes of the window a TMA tile may cover: winLen rounded up to 16 when the buffer has slack past
                          // winLen (session-owned records), rounded DOWN for caller-owned device records (no over-read;
                          // the last partial 16 bytes are then read per sample by the exact path)
    int mode, hasPilot, hasP61, nCh;
    int S, nAct, maxEpochs, capacity;
    int cnoCap, cnoInterval, kernelKind, pad;
    double fs, L, d, PDI, tau1, tau2, pf1, pf2, pf3, factor;
    double tau2over1, PDIoverTau1;   // tau2/tau1 and PDI/tau1 (loop invariant)
    double lockPLD;                  // > 0: drop a channel whose lock detector stays below this (cfg.lockLossPLD)
    int lockIntervals;               // ... for this many consecutive C/N0 intervals
    int iq;                          // 1: x holds interleaved I/Q int8 pairs (fileType 2); window quantities stay in samples
    const uint32_t* codeBits;  // [nCh][3][320] packed primaries: data, pilot, (unused)
    ChanConst* cc;
    ChanState* st;
    EpochParams* params;       // [nCh][2]
    int* ready;                // [nCh] params valid up to this epoch index
    int* stop;                 // [nCh] first epoch that cannot run (INT_MAX while running)
    int* count;                // [nCh] slice arrival counter
    double* partial;           // [nCh][S][18]
    double* acc;               // [nCh][18] slice sums of the epoch in flight (warp-specialised kernel, RED.ADD.F64)
    double* out;               // [nCh][kNFields][capacity]
    double* cno;               // [nCh][kNCno][cnoCap]
    const int* act;            // [nAct] indices of the active channels
    struct FastTab* fastTab;   // [nCh][2] per-epoch tables of the chip-synchronous kernel (or null)
    unsigned long long* counters;  // [4] diagnostics: fast chips, exact-path chips, general-kernel slices
    const EpochParams* olParams;   // open loop (teacher forced): params per channel-epoch, else null
    int olEpochs, olCount;         // open loop: epochs per channel, number of channel-epochs
    unsigned long long* queue;     // ready-task ring of the warp-specialised kernel ((ticket+1)<<32 | payload)
    unsigned* qctl;                // [0] head, [1] tail, [2] channels still running in this launch
    unsigned qMask;                // ring size - 1
    unsigned long long* pubTime;   // [nCh] developer timing: globaltimer of the last publication
    unsigned long long* trace;     // developer tracing: [traceCap][8] timestamps per queue ticket
    unsigned traceCap;
    int nCompute;                  // CTAs [0,nCompute) correlate, the rest close loops
    int ahead;                     // pop a new work item only when <= ahead passes are still pending (-1: no limit)
    int stages, tune;              // pipeline stages (compile-time kFwStages); tune: unused developer bits
    int epochLimit;                // no channel runs an epoch with index >= epochLimit (<= capacity)
};

}  // namespace bds

/*
 * bdsgpu.h — C ABI of libbdsgpu.so: the B200 (sm_100a) acquisition / tracking
 * correlator that replaces the MATLAB hot path of
 * lyf8118/BDS-3-B1C-B2a-SDR-receiver.
 *
 * The reference has no FFI: the seam is three MATLAB function signatures called
 * from postProcessing (paths relative to /root/reference/BDS3_B1C_B2a):
 *   acqResults = acquisition(longSignal, settings)
 *        BDS-3_B1C/postProcessing.m:105-111, BDS-3_B2a/postProcessing.m:100
 *   [trackResults, channel] = WB_tracking / NB_tracking / tracking(fid, channel, settings)
 *        BDS-3_B1C/postProcessing.m:137-143, BDS-3_B2a/postProcessing.m:123
 * Every entry point below names the reference lines it replaces.  A MEX gateway
 * (matlab/bds_mex.c) and a ctypes binding (bds3_b200/_lib.py) sit on top.
 *
 * Conventions: plain pointers and sizes only; caller allocates every output;
 * callee never retains caller pointers past return (except the IF record passed
 * to bds_track_open with BDS_LOC_DEVICE, which must outlive the handle).
 * Return 0 = OK, <0 = error (text via bds_last_error()).  There is no CPU
 * fallback: every compute entry point fails with BDS_ERR_NO_DEVICE when no
 * sm_100 device is usable.
 */
#ifndef BDSGPU_H
#define BDSGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BDS_ABI_VERSION 3

/* error codes */
#define BDS_OK 0
#define BDS_ERR_ARG (-1)
#define BDS_ERR_NO_DEVICE (-2)
#define BDS_ERR_CUDA (-3)
#define BDS_ERR_NOMEM (-4)
#define BDS_ERR_IO (-5)
#define BDS_ERR_UNSUPPORTED (-6)

/* signals */
#define BDS_SIG_B1C 1
#define BDS_SIG_B2A 2

/* tracking modes: which reference function is mirrored */
#define BDS_TRK_B1C_WB 1 /* BDS-3_B1C/WB_tracking.m  (data + QMBOC pilot, 18 sums) */
#define BDS_TRK_B1C_NB 2 /* BDS-3_B1C/NB_tracking.m  (data + BOC(1,1) pilot, 12 sums) */
#define BDS_TRK_B2A 3    /* BDS-3_B2a/tracking.m     (data + pilot, 12 sums) */

/* code components for bds_gen_code / bds_make_code_table */
#define BDS_CODE_B1C_DATA_PRIMARY 1  /* 10230 chips   generateDataBOC11.m:61-82  */
#define BDS_CODE_B1C_PILOT_PRIMARY 2 /* 10230 chips   generatePilotBOC11.m:62-83 */
#define BDS_CODE_B1C_DATA_BOC11 3    /* 20460         generateDataBOC11.m:85-90  */
#define BDS_CODE_B1C_PILOT_BOC11 4   /* 20460         generatePilotBOC11.m:86-94 */
#define BDS_CODE_B1C_PILOT_BOC61 5   /* 122760        generatePilotBOC61.m:89-96 */
#define BDS_CODE_B2A_DATA 6          /* 10230         generateB2aDataCode.m:104-138 */
#define BDS_CODE_B2A_PILOT 7         /* 10230         generateB2aPilotCode.m:104-138 */

/* where a sample / output buffer lives */
#define BDS_LOC_HOST 0
#define BDS_LOC_DEVICE 1

/* kernel selection for tracking (BDS_KERNEL_AUTO picks the chip-synchronous
 * fast path when the configuration allows it, else the general kernel).
 * Configurations with a chip-synchronous kernel - real samples (fileType 1), codeLength 10230 and
 *   B1C WB (pilotTRKflag 2) / NB with pilot: codeFreqBasis 1.023e6, dllCorrelatorSpacing 0.06,
 *       samplingFreq 99.375e6 or 53e6 (B1C/initSettings.m:57), up to 1023 channels per session;
 *   B2a: codeFreqBasis 10.23e6, dllCorrelatorSpacing 0.5, samplingFreq 99.375e6 (B2a/initSettings.m:64).
 * Everything else (other rates / spacings, I/Q records, data-only modes) runs on the general kernel: same
 * results, roughly 30x slower per sample.  BDS_KERNEL_FAST fails with BDS_ERR_UNSUPPORTED instead of falling
 * back; bds_track_counters() tells which kernel ran. */
#define BDS_KERNEL_AUTO 0
#define BDS_KERNEL_GENERAL 1
#define BDS_KERNEL_FAST 2

/* ---- library lifetime ------------------------------------------------------ */
int bds_abi_version(void);
/* Select the CUDA device (ordinal) this process drives.  One process per GPU. */
int bds_init(int device);
void bds_shutdown(void);
const char* bds_last_error(void);
/* number of kernels of this library launched since bds_init (for bench.py's gpu_launches) */
long long bds_launch_count(void);
/* 1 if a CUDA device of compute capability 10.x is usable */
int bds_device_ok(void);

/* ---- code generation (host, integer) ---------------------- */
/* out[n] of +1/-1.  n must equal the component length.  prn in 1..63. */
int bds_gen_code(int component, int prn, int8_t* out, int n);
/* Sampled code table: makeDataTable.m:45-68 / makePilotTable.m:45-69 (B1C, BOC11
 * components) and makeB2aDataTable.m:42-67 / makeB2aPilotTable.m:42-68 (B2a).
 * n_samples must equal round(fs/(codeFreqBasis/codeLength)). */
int bds_make_code_table(int component, int prn, double fs, double codeFreqBasis, int codeLength,
                        int8_t* out, int n_samples);

/* ---- acquisition -------------------------------- */
typedef struct bds_acq_cfg {
    double samplingFreq;  /* settings.samplingFreq */
    double IF;            /* settings.IF */
    double codeFreqBasis; /* settings.codeFreqBasis */
    int32_t codeLength;   /* settings.codeLength (10230) */
    double acqSearchBand; /* settings.acqSearchBand [Hz] */
    double acqStep;       /* settings.acqStep [Hz] */
    double acqThreshold;  /* settings.acqThreshold */
    int32_t acqCohT;      /* B1C: settings.acqCohT [ms]; ignored for B2a */
    int32_t pilotACQflag; /* B1C: settings.pilotACQflag; ignored for B2a */
    int32_t fineNoncoh;   /* B2a: settings.fineNoncoh [ms]; ignored for B1C */
    int32_t fileType;     /* settings.fileType: 0/1 = real samples, 2 = interleaved I/Q pairs, longSignal = I + 1i*Q
                           * (postProcessing.m:94-99); x then holds 2*n int8 values for n samples */
    /* resampling pre-conditioner (acquisition.m:56-123, B2a acquisition.m:56-124): with resamplingflag == 1 and
     * samplingFreq > resamplingThreshold the record is band-pass filtered around the IF (fir1(700) + filtfilt), the search
     * runs at a band-pass sampling rate on every index-th sample and codePhase / carrFreq are mapped back (:321-338) */
    double resamplingThreshold; /* settings.resamplingThreshold [Hz] */
    int32_t resamplingflag;     /* settings.resamplingflag */
    int32_t tune;               /* 0.  Test / developer hook, results unchanged: bit 0 = run the inverse passes on the generic
                                 * kernels for every transform shape, bit 1 = other row tiling of the specialised row pass,
                                 * bit 2 = inverse passes on one stream instead of two, bit 3 = on four */
} bds_acq_cfg;

/* Replaces acquisition(longSignal, settings):
 *   B1C: BDS-3_B1C/acquisition.m:129-338   B2a: BDS-3_B2a/acquisition.m:130-365
 * x: n int8 IF samples (host or device per x_loc; real, or n interleaved I/Q pairs = 2n bytes with cfg.fileType 2).
 * prn[n_prn] = acqSatelliteList.
 * carrFreq/codePhase/peakMetric: caller-allocated [max_prn] doubles, indexed PRN-1,
 * zero-filled by the callee, 0 = not found (acquisition.m:161-165).
 * prn_lo/prn_hi shard the list for multi-GPU: only list entries with index in
 * [prn_lo, prn_hi) are searched (pass 0, n_prn for all).
 * dbg (optional, may be NULL): [max_prn*4] doubles: coarse bin index (0-based), coarse
 * code phase (1-based), peak size, normaliser (sigPower or second peak). */
int bds_acquire(int signal, const int8_t* x, size_t n, int x_loc, const bds_acq_cfg* cfg,
                const int32_t* prn, int n_prn, int prn_lo, int prn_hi, double* carrFreq,
                double* codePhase, double* peakMetric, int max_prn, double* dbg);

/* ---- tracking ------------------------------------------- */
typedef struct bds_trk_cfg {
    double samplingFreq;         /* settings.samplingFreq */
    double codeFreqBasis;        /* settings.codeFreqBasis */
    int32_t codeLength;          /* settings.codeLength */
    double dllCorrelatorSpacing; /* settings.dllCorrelatorSpacing [chips] */
    double intTime;              /* settings.intTime (PDIcode) */
    int32_t pilotTRKflag;        /* settings.pilotTRKflag */
    int32_t CNoInterval;         /* settings.CNoInterval */
    /* loop constants, computed by the caller exactly as the reference does
     * (Common/calcLoopCoef.m:41-45, calcLoopCoefCarr.m:41-56,
     *  BDS-3_B1C/include/CalcWeighingFactor.m:43-81) */
    double tau1code, tau2code;
    double pf3, pf2, pf1;
    double wbFactor;             /* WB only */
    int32_t kernel;              /* BDS_KERNEL_* */
    int32_t reserved;            /* 0; test hooks: bit 0 = widen the fast kernel's exact-path guard band, bit 1 = narrow band on the
                                  * wide-band chip body instead of the narrow-band one */
    /* Tuning and diagnostics of the chip-synchronous kernel.  All 0 = library defaults.  The library never reads the
     * caller's environment: everything that changes its behaviour is in this struct. */
    int32_t fwPassesPerTask;     /* passes of 512 chips per queued (channel, epoch, slice) task, 1..8 */
    int32_t fwPrefetch;          /* tasks a CTA may stage ahead of its compute warps, plus one (1 = none) */
    int32_t debug;               /* BDS_DBG_* bits */
    int32_t traceTickets;        /* BDS_DBG_TRACE: number of queue tickets to record (bds_track_dump_trace) */
    /* Lock-loss status and early channel drop.  An EXTENSION, off by default (0): the reference
     * copies channel.status unconditionally (WB_tracking.m:485-488).  With lockLossPLD > 0 a channel whose lock
     * detector (Calc_CNo_PLD.m:70-73; the pilot's when a pilot is tracked, else the data component's) stays below
     * lockLossPLD for lockLossIntervals consecutive C/N0 intervals runs no further epoch; bds_trk_out.lockLostEpoch
     * reports the number of epochs it completed. */
    double lockLossPLD;
    int32_t lockLossIntervals;   /* >= 1 */
    int32_t fwMaxCtas;           /* 0 = one CTA per SM; else the chip-synchronous kernel's persistent grid is limited to this
                                  * many CTAs, leaving SMs to kernels on other streams (a second session, NCCL) */
    int32_t fileType;            /* settings.fileType: 0/1 = real int8 samples; 2 = interleaved I/Q int8 pairs,
                                  * rawSignal = I + 1i*Q (WB_tracking.m:155-159,270-274; B2a tracking.m:143-147,241-244).
                                  * Every sample count, offset and absoluteSample of this API stays in SAMPLES (the reference's
                                  * ftell/dataAdaptCoeff); buffers and files hold 2 bytes per sample.  I/Q records run on
                                  * the general kernel (BDS_KERNEL_FAST is refused). */
    int32_t b2aClusterSize;      /* B2a chip-synchronous kernel: CTAs (one thread-block cluster) per channel: 0 = the largest of
                                  * 8 / 4 / 2 / 1 for which all channels are resident at once, else 1, 2, 4 or 8 */
} bds_trk_cfg;
#define BDS_DBG_TIMING 1 /* per-stage cycle counters of the tracking kernel, printed by bds_track_counters */
#define BDS_DBG_TRACE 2  /* per-ticket timestamps */

typedef struct bds_channel {
    int32_t PRN;         /* 0 = unused channel (WB_tracking.m:165) */
    int32_t status;      /* 'T' or '-' */
    double acquiredFreq; /* channel.acquiredFreq */
    double codePhase;    /* channel.codePhase: 1-based sample lag (acquisition.m:307) */
    double codeFreq;     /* channel.codeFreq */
} bds_channel;

/* Output planes, each [n_ch * n_epochs] doubles, row-major by channel.  NULL planes
 * are skipped.  Initial values are written by the callee exactly as the reference
 * preallocates them (0 or Inf; WB_tracking.m:53-112). */
typedef struct bds_trk_out {
    double *absoluteSample, *codeFreq, *carrFreq;
    double *I_P, *I_E, *I_L, *Q_E, *Q_P, *Q_L;
    double *Pilot_I_P, *Pilot_I_E, *Pilot_I_L, *Pilot_Q_E, *Pilot_Q_P, *Pilot_Q_L;
    double *dllDiscr, *dllDiscrFilt, *pllDiscr, *pllDiscrFilt;
    double *remCodePhase, *remCarrPhase;
    /* [n_ch * floor(n_epochs/CNoInterval)] */
    double *DataCNo, *DataPLD, *PilotCNo, *PilotPLD, *TotalCNo;
    /* optional raw correlator sums [n_ch * n_epochs * 18]:
     * {d,p11,p61} x {E,P,L} x {I,Q}; index = fam*6 + {E,P,L}*2 + {I,Q} */
    double* raw;
    /* [n_ch] number of epochs completed per channel (short input stops a channel
     * like WB_tracking.m:279-283) */
    int32_t* epochsDone;
    /* [n_ch] optional: 0, or the number of epochs completed when the channel was dropped for loss of lock
     * (cfg.lockLossPLD > 0 only) */
    int32_t* lockLostEpoch;
} bds_trk_out;

typedef struct bds_trk bds_trk;

/* Open a tracking session on the IF record x[n] (int8; n samples = n bytes, or 2n bytes of I/Q pairs with cfg.fileType 2).  With BDS_LOC_HOST the
 * record is copied to the device; with BDS_LOC_DEVICE it is used in place (16-byte aligned; all n samples
 * are usable and nothing is read past x[n-1], so an epoch may end on the record's last sample).
 * skipNumberOfBytes is settings.skipNumberOfBytes; channel blocks start at
 * skip + codePhase - 1 (WB_tracking.m:174-176). */
int bds_track_open(int mode, const bds_trk_cfg* cfg, const int8_t* x, size_t n, int x_loc,
                   long long skipNumberOfBytes, const bds_channel* ch, int n_ch, bds_trk** out);
/* Same, reading the record from a file (the reference's `fid`): schar, fileType 1 or 2 (cfg.fileType); max_samples
 * (0 = whole file) and skipNumberOfBytes count samples, as the reference's dataAdaptCoeff*(skip + codePhase - 1) does. */
int bds_track_open_file(int mode, const bds_trk_cfg* cfg, const char* path, long long skipNumberOfBytes,
                        long long max_samples, const bds_channel* ch, int n_ch, bds_trk** out);
/* Replace the resident IF window: the session keeps loop state; x[n] holds samples
 * [first_sample, first_sample+n) of the record (streaming chunks for the e2e path). */
int bds_track_feed(bds_trk* h, const int8_t* x, size_t n, int x_loc, long long first_sample);
/* Run up to n_epochs more epochs per channel (closed loop, persistent kernel).
 * Results are appended at epoch offset = epochs already run.  out planes are host
 * pointers with row stride out_stride (>= total epochs).  Returns 0 or error. */
int bds_track_run(bds_trk* h, int n_epochs, const bds_trk_out* out, int out_stride);
/* Device-resident variant: launches only; results stay on the device until
 * bds_track_fetch.  Used by bench.py for the HBM-resident timing and by the
 * multi-GPU path (NCCL gather of the packed device block). */
int bds_track_run_async(bds_trk* h, int n_epochs);
/* Host-record variant of bds_track_run_async for the end-to-end path (the reference streams the file
 * with one fread per epoch, WB_tracking.m:265): x[n] (host; pinned memory gives full PCIe rate) is
 * copied to the device in chunk_bytes pieces (0 = 128 MiB) on a copy stream while the persistent
 * kernel tracks what has already arrived; n_epochs = total epochs per channel.  Replaces the
 * session's resident record.  x must stay valid until bds_track_sync / bds_track_fetch returns. */
int bds_track_run_streamed(bds_trk* h, const int8_t* x, size_t n, size_t chunk_bytes, int n_epochs);
/* Device-record variant for a record that is still arriving on the device (multi-GPU: every rank uploads 1/N of
 * each chunk and the ranks all-gather it over NVLink): one asynchronous launch over the first n_avail samples of
 * x_dev (valid and device-synchronised by the caller); channels stop at the end of the data or at epoch index
 * epoch_limit and are resumed by the next call with a larger n_avail.  Nothing is read past x_dev[n_avail-1]. */
int bds_track_run_window(bds_trk* h, const int8_t* x_dev, size_t n_avail, int epoch_limit);
int bds_track_sync(bds_trk* h);
int bds_track_fetch(bds_trk* h, const bds_trk_out* out, int out_stride);
/* packed device result block: [n_ch][n_fields=21+18][capacity] doubles */
int bds_track_device_block(bds_trk* h, void** dev_ptr, size_t* bytes, int* n_fields, int* capacity);
/* total IF samples consumed so far (sum over channels of blksize), epochs run */
int bds_track_stats(bds_trk* h, long long* channel_samples, int* epochs_run, float* last_kernel_ms);
/* diagnostics: out4 = {chips integrated by the chip-synchronous body, chips re-evaluated by its exact
 * per-sample path, slices run by the general kernel, 0}.  h == NULL: counters of the last
 * bds_track_correlate_open_loop call (out4[3] = device microseconds of its correlator kernel). */
int bds_track_counters(bds_trk* h, long long* out4);
/* developer tracing (cfg.debug & BDS_DBG_TRACE, cfg.traceTickets): dump per-work-item timestamps */
int bds_track_dump_trace(bds_trk* h, const char* path);
/* reset loop state to the initial channel state (re-run the same record) */
int bds_track_reset(bds_trk* h);
void bds_track_close(bds_trk* h);

/* Open-loop ("teacher-forced") correlator: SURVEY §7 step 4.  nco[n_ch*n_epochs*6] =
 * {block start (0-based sample), blksize, remCodePhase, codePhaseStep, carrFreq,
 * remCarrPhase}; sums[n_ch*n_epochs*18] as bds_trk_out.raw.  Replaces exactly
 * WB_tracking.m:289-372 / NB_tracking.m:271-343 / B2a tracking.m:260-331. */
int bds_track_correlate_open_loop(int mode, const bds_trk_cfg* cfg, const int8_t* x, size_t n, int x_loc,
                                  const int32_t* prn, int n_ch, int n_epochs, const double* nco,
                                  double* sums);

/* ---- frame synchronisation on the device ------------ */
/* The PRN's 1800-chip B1C pilot secondary code as +1/-1 (generate2ndCode.m:44-84). */
int bds_secondary_code(int prn, int8_t* out1800);
/* The correlation that starts the reference's nav decoding, on the prompt plane the tracker produced:
 *   BDS_SIG_B1C  BCNAV1decoding.m:66-91: bits = sign(prompt) (caller passes Pilot_I_P for pilotTRKflag 2, else
 *                Pilot_Q_P), XcorrResult = xcorr(bits, Secondary(prn)) for lags >= 0, index = find(abs(.) >= 1799.5)
 *   BDS_SIG_B2A  BCNAV2decoding.m:69-97: bits = sign(I_P), pattern = kron(preamble_bits, secondCode),
 *                index = find(abs(.) > 115); prn is ignored
 * prompt[n]: host or device doubles (loc).  xcorr (optional, host): [n] correlation values for lags 0..n-1 (integers).
 * index (host): up to index_cap 1-based lags in ascending order, *n_index = how many the record holds. */
int bds_frame_sync(int signal, int prn, const double* prompt, int n, int loc, double* xcorr, int32_t* index,
                   int index_cap, int32_t* n_index);

/* ---- synthetic IF (bench / tests; the reference has no generator) ---------- */
typedef struct bds_sat {
    int32_t PRN;
    int32_t reserved;
    double doppler;    /* Hz at IF */
    double codeDelay;  /* samples: first code-period start, in [0, samplesPerCode) */
    double carrPhase;  /* rad at sample 0 */
    double amplitude;  /* LSB */
} bds_sat;
/* Writes n int8 samples starting at absolute sample index first_sample into out
 * (device or host).  Signal model: SURVEY §8(d).  Deterministic in (seed, sample). */
int bds_synth_if(int signal, double fs, double IF, double carrFreqBasis, double codeFreqBasis,
                 const bds_sat* sats, int n_sats, double noise_sigma, uint64_t seed,
                 long long first_sample, size_t n, int8_t* out, int out_loc);

/* raw device helpers so a host language without a CUDA binding can hold buffers */
int bds_dev_alloc(void** p, size_t bytes);
int bds_dev_free(void* p);
int bds_host_alloc_pinned(void** p, size_t bytes);
int bds_host_free_pinned(void* p);
int bds_memcpy_h2d(void* dst, const void* src, size_t bytes);
int bds_memcpy_d2h(void* dst, const void* src, size_t bytes);
int bds_dev_sync(void);

#ifdef __cplusplus
}
#endif
#endif /* BDSGPU_H */

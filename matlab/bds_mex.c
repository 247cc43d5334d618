/*
 * bds_mex.c — MEX gateway between the reference's MATLAB call surface and libbdsgpu.so.
 *
 *   mex -R2018a -I../include bds_mex.c -L../bds-3-b1c-b2a-sdr-receiver_b200 -lbdsgpu
 *
 * A thin shim over the C ABI of include/bdsgpu.h: it only marshals mxArrays.  It cannot be
 * compiled in the build image (no mex.h / MATLAB there); every code path below the mxArray
 * marshalling is exercised through the ctypes binding (bds3_b200/_lib.py) by tests/.
 *
 * Commands (first argument is a string):
 *   [carrFreq, codePhase, peakMetric] = bds_mex('acquire', signal, int8(longSignal), acqCfg, prnList)
 *        replaces acquisition(longSignal, settings)
 *        BDS-3_B1C/postProcessing.m:105-111, BDS-3_B2a/postProcessing.m:100
 *        a complex int8 longSignal (settings.fileType == 2, postProcessing.m:96-99) is passed as its interleaved I, Q
 *        pairs with cfg.fileType = 2
 *   [planes, cno, epochsDone] = bds_mex('track', mode, fileName, skipNumberOfBytes, trkCfg, channelMatrix, nEpochs)
 *        replaces {WB_,NB_,}tracking(fid, channel, settings)
 *        BDS-3_B1C/postProcessing.m:137-143, BDS-3_B2a/postProcessing.m:123
 *        planes: [nEpochs x 21 x nCh] doubles in bds_trk_out plane order, cno: [nCno x 5 x nCh]
 *   code = bds_mex('gencode', component, prn)          (generate*.m replacements)
 * acqCfg / trkCfg are double row vectors in the field order of bds_acq_cfg / bds_trk_cfg;
 * channelMatrix is [nCh x 5] = PRN, status (char code), acquiredFreq, codePhase, codeFreq.
 */
#ifdef MATLAB_MEX_FILE
#include <string.h>

#include "bdsgpu.h"
#include "mex.h"

static int g_ready = 0;

static void at_exit(void) {
    bds_shutdown();
    g_ready = 0;
}

static void ensure_init(void) {
    if (g_ready) return;
    if (bds_init(0) != BDS_OK) mexErrMsgIdAndTxt("bds:init", "%s", bds_last_error());
    mexLock(); /* keep the CUDA context for the MATLAB session */
    mexAtExit(at_exit);
    g_ready = 1;
}

static void check(int rc) {
    if (rc != BDS_OK) mexErrMsgIdAndTxt("bds:error", "libbdsgpu error %d: %s", rc, bds_last_error());
}

static void do_acquire(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    if (nrhs != 5 || !mxIsInt8(prhs[2])) mexErrMsgIdAndTxt("bds:args", "acquire: (signal, int8 samples, cfg, prnList)");
    const int signal = (int)mxGetScalar(prhs[1]);
    const double* c = mxGetDoubles(prhs[3]);
    bds_acq_cfg cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.samplingFreq = c[0]; cfg.IF = c[1]; cfg.codeFreqBasis = c[2]; cfg.codeLength = (int32_t)c[3];
    cfg.acqSearchBand = c[4]; cfg.acqStep = c[5]; cfg.acqThreshold = c[6]; cfg.acqCohT = (int32_t)c[7];
    cfg.pilotACQflag = (int32_t)c[8]; cfg.fineNoncoh = (int32_t)c[9];
    if (mxGetNumberOfElements(prhs[3]) > 11) { /* resampling pre-conditioner, acquisition.m:56-123 */
        cfg.resamplingThreshold = c[10];
        cfg.resamplingflag = (int32_t)c[11];
    }
    const mwSize nprn = mxGetNumberOfElements(prhs[4]);
    const double* pl = mxGetDoubles(prhs[4]);
    int32_t prn[64];
    int maxprn = 0;
    if (nprn > 63) mexErrMsgIdAndTxt("bds:args", "at most 63 PRNs");
    for (mwSize i = 0; i < nprn; ++i) {
        prn[i] = (int32_t)pl[i];
        if (prn[i] > maxprn) maxprn = prn[i];
    }
    for (int k = 0; k < 3; ++k) plhs[k] = mxCreateDoubleMatrix(1, maxprn, mxREAL); /* MATLAB owns the outputs */
    /* interleaved complex API (-R2018a): a complex int8 array is stored as I0, Q0, I1, Q1, ... = the fileType-2 file layout */
    const int iq = mxIsComplex(prhs[2]) ? 1 : 0;
    cfg.fileType = iq ? 2 : 1;
    const int8_t* xs = iq ? (const int8_t*)mxGetComplexInt8s(prhs[2]) : (const int8_t*)mxGetInt8s(prhs[2]);
    check(bds_acquire(signal, xs, mxGetNumberOfElements(prhs[2]), BDS_LOC_HOST, &cfg,
                      prn, (int)nprn, 0, (int)nprn, mxGetDoubles(plhs[0]), mxGetDoubles(plhs[1]),
                      mxGetDoubles(plhs[2]), maxprn, NULL));
    (void)nlhs;
}

static void do_track(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    if (nrhs != 7) mexErrMsgIdAndTxt("bds:args", "track: (mode, fileName, skip, cfg, channels, nEpochs)");
    const int mode = (int)mxGetScalar(prhs[1]);
    char path[4096];
    if (mxGetString(prhs[2], path, sizeof path)) mexErrMsgIdAndTxt("bds:args", "fileName too long");
    const long long skip = (long long)mxGetScalar(prhs[3]);
    const double* c = mxGetDoubles(prhs[4]);
    bds_trk_cfg cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.samplingFreq = c[0]; cfg.codeFreqBasis = c[1]; cfg.codeLength = (int32_t)c[2];
    cfg.dllCorrelatorSpacing = c[3]; cfg.intTime = c[4]; cfg.pilotTRKflag = (int32_t)c[5];
    cfg.CNoInterval = (int32_t)c[6]; cfg.tau1code = c[7]; cfg.tau2code = c[8];
    cfg.pf3 = c[9]; cfg.pf2 = c[10]; cfg.pf1 = c[11]; cfg.wbFactor = c[12]; cfg.kernel = BDS_KERNEL_AUTO;
    cfg.fileType = mxGetNumberOfElements(prhs[4]) > 13 ? (int32_t)c[13] : 1; /* settings.fileType (WB_tracking.m:155-159) */
    const mwSize nch = mxGetM(prhs[5]);
    const double* cm = mxGetDoubles(prhs[5]); /* column major [nCh x 5] */
    bds_channel* ch = (bds_channel*)mxCalloc(nch, sizeof(bds_channel));
    for (mwSize i = 0; i < nch; ++i) {
        ch[i].PRN = (int32_t)cm[i];
        ch[i].status = (int32_t)cm[i + nch];
        ch[i].acquiredFreq = cm[i + 2 * nch];
        ch[i].codePhase = cm[i + 3 * nch];
        ch[i].codeFreq = cm[i + 4 * nch];
    }
    const int nE = (int)mxGetScalar(prhs[6]);
    const int nC = cfg.CNoInterval > 0 ? nE / cfg.CNoInterval : 0;
    const mwSize d1[3] = {(mwSize)nE, 21, nch}, d2[3] = {(mwSize)nC, 5, nch};
    plhs[0] = mxCreateNumericArray(3, d1, mxDOUBLE_CLASS, mxREAL);
    plhs[1] = mxCreateNumericArray(3, d2, mxDOUBLE_CLASS, mxREAL);
    plhs[2] = mxCreateNumericMatrix(1, nch, mxINT32_CLASS, mxREAL);
    /* the C ABI wants one plane per field, [nCh][stride]: fetch into scratch, then interleave */
    double* scratch = (double*)mxMalloc(sizeof(double) * (size_t)nch * nE * 21);
    double* cscr = (double*)mxCalloc((size_t)nch * (nC > 0 ? nC : 1) * 5, sizeof(double));
    bds_trk_out o;
    memset(&o, 0, sizeof o);
    double** planes = (double**)&o; /* the 21 per-epoch planes are the first 21 pointers of bds_trk_out */
    for (int f = 0; f < 21; ++f) planes[f] = scratch + (size_t)f * nch * nE;
    if (nC > 0) {
        o.DataCNo = cscr; o.DataPLD = cscr + (size_t)nch * nC; o.PilotCNo = cscr + (size_t)2 * nch * nC;
        o.PilotPLD = cscr + (size_t)3 * nch * nC; o.TotalCNo = cscr + (size_t)4 * nch * nC;
    }
    o.epochsDone = (int32_t*)mxGetData(plhs[2]);
    bds_trk* h = NULL;
    check(bds_track_open_file(mode, &cfg, path, skip, 0, ch, (int)nch, &h));
    int rc = bds_track_run(h, nE, &o, nE);
    bds_track_close(h);
    check(rc);
    double* out = mxGetDoubles(plhs[0]);
    for (mwSize c2 = 0; c2 < nch; ++c2)
        for (int f = 0; f < 21; ++f)
            memcpy(out + ((size_t)c2 * 21 + f) * nE, scratch + ((size_t)f * nch + c2) * nE, sizeof(double) * nE);
    double* co = mxGetDoubles(plhs[1]);
    for (mwSize c2 = 0; c2 < nch && nC > 0; ++c2)
        for (int f = 0; f < 5; ++f)
            memcpy(co + ((size_t)c2 * 5 + f) * nC, cscr + ((size_t)f * nch + c2) * nC, sizeof(double) * nC);
    mxFree(scratch);
    mxFree(cscr);
    mxFree(ch);
    (void)nlhs;
}

static void do_gencode(mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    static const int len[8] = {0, 10230, 10230, 20460, 20460, 122760, 10230, 10230};
    if (nrhs != 3) mexErrMsgIdAndTxt("bds:args", "gencode: (component, prn)");
    const int comp = (int)mxGetScalar(prhs[1]);
    if (comp < 1 || comp > 7) mexErrMsgIdAndTxt("bds:args", "component 1..7");
    mxArray* tmp = mxCreateNumericMatrix(1, len[comp], mxINT8_CLASS, mxREAL);
    check(bds_gen_code(comp, (int)mxGetScalar(prhs[2]), (int8_t*)mxGetInt8s(tmp), len[comp]));
    plhs[0] = mxCreateDoubleMatrix(1, len[comp], mxREAL);
    const int8_t* s = (const int8_t*)mxGetInt8s(tmp);
    double* d = mxGetDoubles(plhs[0]);
    for (int i = 0; i < len[comp]; ++i) d[i] = (double)s[i];
    mxDestroyArray(tmp);
}

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    char cmd[32];
    if (nrhs < 1 || mxGetString(prhs[0], cmd, sizeof cmd)) mexErrMsgIdAndTxt("bds:args", "first argument: command string");
    if (!strcmp(cmd, "gencode")) {
        do_gencode(plhs, nrhs, prhs);
        return;
    }
    ensure_init();
    if (!strcmp(cmd, "acquire")) do_acquire(nlhs, plhs, nrhs, prhs);
    else if (!strcmp(cmd, "track")) do_track(nlhs, plhs, nrhs, prhs);
    else mexErrMsgIdAndTxt("bds:args", "unknown command %s", cmd);
}
#else
/* Compiled outside MATLAB (e.g. `gcc -fsyntax-only`): nothing to build. */
typedef int bds_mex_requires_matlab;
#endif

/*
 * C restatement of the tracking correlator of lyf8118/BDS-3-B1C-B2a-SDR-receiver.
 *
 * TEST INFRASTRUCTURE ONLY — the independent second implementation that pins the numpy
 * oracle (oracle/bds_oracle.py) and the timed CPU baseline of bench.py.  Never linked
 * into or called from the shipped product path.  "Parity unpinned by reference goldens":
 * the reference ships no test vectors and MATLAB is not available in this image.
 *
 * Follows, line by line (paths under /root/reference/BDS3_B1C_B2a):
 *   BDS-3_B1C/WB_tracking.m:289-372   (mode 1, 18 sums)
 *   BDS-3_B1C/NB_tracking.m:271-343   (mode 2, 12 sums)
 *   BDS-3_B2a/tracking.m:260-331      (mode 3, 12 sums)
 * float64 throughout, plain (unfused) multiply/add for every value feeding a ceil().
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off (see oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#define TWO_PI 6.283185307179586476925286766559

/* MATLAB a:d:b element k of n+1 (two-ended construction, see bds_oracle.colon) */
static inline double colon_elem(double a, double d, double c, long n, long k) {
    long h = n / 2;
    if ((n % 2) == 0 && k == h) return (a + c) / 2;
    if (k <= h) return a + (double)k * d;
    return c - (double)(n - k) * d;
}

/*
 * mode: 1 WB, 2 NB, 3 B2a.  data/pilot/pilot61: 1-padded replicas [code(end) code code(1)] as
 * +-1 int8 (pilot / pilot61 may be NULL).  out[18]: index fam*6 + {E,P,L}*2 + {I,Q}.
 */
void orc_correlate_epoch(int mode, const int8_t* raw, long blksize, const int8_t* data, const int8_t* pilot,
                         const int8_t* pilot61, double remCodePhase, double codePhaseStep, double carrFreq,
                         double remCarrPhase, double earlyLateSpc, double fs, double codeLength, double* out,
                         double* remCodeNext, double* remCarrNext) {
    const int b1c = mode != 3;
    const double mul = b1c ? 2.0 : 1.0;
    const long n = blksize - 1;
    const double base = (double)n * codePhaseStep + remCodePhase;
    const double dd = codePhaseStep * mul;
    double a[3], c[3];
    a[0] = (remCodePhase - earlyLateSpc) * mul;
    a[1] = remCodePhase * mul;
    a[2] = (remCodePhase + earlyLateSpc) * mul;
    c[0] = (base - earlyLateSpc) * mul;
    c[1] = base * mul;
    c[2] = (base + earlyLateSpc) * mul;
    for (int i = 0; i < 18; ++i) out[i] = 0.0;
    const double w = carrFreq * 2.0 * 3.14159265358979323846;
    for (long k = 0; k < blksize; ++k) {
        double trig = (w * ((double)k / fs)) + remCarrPhase;
        double cs = cos(trig), sn = sin(trig);
        double x = (double)raw[k];
        double iB, qB;
        if (b1c) { /* exp(-i*trig): I = real, Q = imag */
            iB = cs * x;
            qB = -sn * x;
        } else {   /* exp(+i*trig): Q = real, I = imag */
            qB = cs * x;
            iB = sn * x;
        }
        for (int o = 0; o < 3; ++o) {
            double t = colon_elem(a[o], dd, c[o], n, k);
            long idx = (long)ceil(t);
            double sd = data[idx];
            out[0 * 6 + o * 2 + 0] += sd * iB;
            out[0 * 6 + o * 2 + 1] += sd * qB;
            if (pilot) {
                double sp = pilot[idx];
                out[1 * 6 + o * 2 + 0] += sp * iB;
                out[1 * 6 + o * 2 + 1] += sp * qB;
            }
            if (pilot61) {
                double s6 = pilot61[(long)ceil(t * 6)];
                out[2 * 6 + o * 2 + 0] += s6 * iB;
                out[2 * 6 + o * 2 + 1] += s6 * qB;
            }
        }
    }
    /* remCodePhase = tcode(blksize)[/2] + codePhaseStep - codeLength ; tcode(blksize) == stop expr */
    *remCodeNext = (b1c ? c[1] / 2 : c[1]) + codePhaseStep - codeLength;
    double trigEnd = (w * ((double)blksize / fs)) + remCarrPhase;
    *remCarrNext = fmod(trigEnd, TWO_PI);
}

#!/usr/bin/env python
"""Generates bds_track_fast_gen.inc: the straight-line per-chip body of the chip-synchronous
B1C wide-band correlator for one (samplingFreq, codeFreqBasis, dllCorrelatorSpacing).

One thread integrates one primary-code chip.  In units of BOC(6,1) sub-chips (1/12 chip) the
prompt, late and early replicas change sign at offsets j, j+(1-delta) and j+delta inside every
sub-chip (delta = 12 * dllCorrelatorSpacing, 0.5 < delta < 1), so a chip splits into 36
segments in which all nine replicas are constant.  For a fixed sampling rate the k-th segment
boundary falls at sample R_k = floor(beta_k * S) or one later, depending only on the sub-sample
phase Psi of the thread's first sample: sample R_k still belongs to the old segment iff
Psi <= Theta_k = frac(beta_k * S) ("jitter" sample).  The 36 thresholds cut (0, 1] into 37
intervals ("ranks"); their ORDER is a property of the nominal geometry (the thresholds move by
< 1e-3 sample over +-15 kHz of Doppler while neighbours are >= 8e-3 apart), so everything that
depends on the rank only is tabulated here at generation time:

    kFastMask[rank]         bit k-1 set <=> the jitter sample of boundary k is old.  From it the body forms, with one
                            SEL between two immediates per boundary, the PREFIX byte mask P_k of the samples of word
                            (R_k >> 2) that lie before boundary k (bytes 0 .. (R_k & 3) - 1, plus byte (R_k & 3) when
                            the jitter sample is old); every segment piece of a word is X & P, X & ~P or
                            X & ~P_k & P_k+1: one LOP3.  (A table of the 36 prefix masks per rank was measured too:
                            nine conflicting LDS.128 per chip made shared memory the bottleneck, 86 % of its peak.)
    kFastRankLo[bin]        the rank of the first nominal threshold that can lie in or above the 1/512-wide bin of Psi
                            (at most one threshold lies within a bin and its neighbours: one compare gives the rank)
    kFastThrNom[s], kFastPos[k-1]   nominal sorted thresholds (2^32 fixed point) / sorted position of threshold k

Arithmetic of the body: the int8 samples are never unpacked.  Each re-aligned 32-bit word of
four samples is AND-masked to the bytes that belong to a segment and fed to IDP.2A
(dp2a: two int16 carrier values x two int8 samples, int32 accumulate), once for the real and once
for the imaginary carrier table, straight into the accumulator of that segment's class:
    {E,O}{1,2}{A,B,C}  - even/odd sub-chip, first/second half chip, segment class A/B/C
    A1,B1,A7,B7,B6,C6,B12,C12 - the eight segments that also form the BOC(1,1) early/late windows
from which X = H2 - H1, SA, SB, SC, W1a, W1b, W2a, W2b follow by additions at the end of the chip.

Narrow band (`nb` as a fourth argument): NB_tracking.m correlates the data and the BOC(1,1) pilot only, whose E/P/L replicas
change sign at 0, d, 1/2-d, 1/2, 1/2+d, 1-d of a chip: SIX segments instead of 36.  The body then accumulates the six segment
sums N0..N5 directly (most words are unmasked) and the combination reduces to X = (N3+N4+N5) - (N0+N1+N2), W1a = N0, W2a = N2,
W1b = N3, W2b = N5 - the five basis sums the data / BOC(1,1) tail of fast_chip uses; SA, SB, SC do not exist.

Usage: python gen_fast_wb.py [fs_hz fc_hz d] > bds_track_fast_gen.inc
       python gen_fast_wb.py 53000000 1023000 3/50 > bds_track_fast_gen_53.inc   (B1C/initSettings.m:57)
       python gen_fast_wb.py 99375000 1023000 3/50 nb > bds_track_fast_gen_nb.inc
       python gen_fast_wb.py 53000000 1023000 3/50 nb > bds_track_fast_gen_53_nb.inc
"""
from __future__ import annotations

import sys
from fractions import Fraction as F

RANK_BINS = 512     # default; main() doubles it while the nominal thresholds are closer than three bins


def geometry(fs, fc, d, nb=False):
    S = fs / (12 * fc)                 # samples per sub-chip
    delta = 12 * d
    assert F(1, 2) < delta < 1, "generator assumes 0.5 < 12*d < 1"
    beta = []
    if nb:                             # BOC(1,1) replicas only: 0, d, 1/2-d, 1/2, 1/2+d, 1-d, 1 of a chip, in sub-chips
        beta = [F(0), delta, 6 - delta, F(6), 6 + delta, 12 - delta]
    else:
        for j in range(12):
            beta += [F(j), j + (1 - delta), j + delta]
    beta.append(F(12))                 # nseg + 1 boundaries, beta[0] = chip start, beta[nseg] = chip end
    nseg = len(beta) - 1
    R = [int(b * S // 1) for b in beta]
    assert all(R[k + 1] - R[k] >= 1 for k in range(nseg)), "segments shorter than one sample"
    theta = [b * S - r for b, r in zip(beta, R)]        # exact fractions, theta[k] in [0, 1)
    return S, beta, R, theta


def rank_tables(R, theta):
    """sorted thresholds, positions, rank -> decision bits, prefix masks, rank lower bounds"""
    nseg = len(R) - 1
    order = sorted(range(1, nseg + 1), key=lambda k: (theta[k], k))   # boundary numbers by ascending threshold
    pos = [0] * nseg
    for s, k in enumerate(order):
        pos[k - 1] = s
    thr_nom = [int(theta[k] * (1 << 32)) for k in order]
    gaps = [b - a for a, b in zip(thr_nom, thr_nom[1:])]
    w = (1 << 32) // RANK_BINS
    assert min(gaps) > 3 * w, "nominal thresholds too close for the one-compare rank search"
    assert thr_nom[0] > 2 * w and thr_nom[-1] < (1 << 32) - 2 * w, "a threshold too close to the chip edge decision"
    masks = []
    for j in range(nseg + 1):          # j = number of thresholds < Psi
        # bit k-1: Theta_k >= Psi, i.e. the jitter sample R_k still belongs to segment k-1
        masks.append(sum(1 << (k - 1) for k in range(1, nseg + 1) if pos[k - 1] >= j))
    rank_lo = []
    for b in range(RANK_BINS):
        lo = (b - 1) * w
        rank_lo.append(sum(1 for t in thr_nom if t < lo))
    return order, pos, thr_nom, masks, rank_lo


def emit_body(R, acc_name):
    """straight-line per-word code with one prefix mask per boundary (FAST_PSEL(k, lo, hi), k = 1..nseg)"""
    nseg = len(R) - 1
    nwords = (R[nseg] + 1 + 3) // 4
    lines = []
    e = lines.append
    loaded = set()

    def pfx(k):                        # P_k = (jitter sample old) ? bytes 0..b : bytes 0..b-1, b = R_k & 3
        if k not in loaded:
            loaded.add(k)
            b = R[k] & 3
            lo = (1 << (8 * b)) - 1
            hi = (lo | (0xFF << (8 * b))) & 0xFFFFFFFF
            e("const unsigned P%d = FAST_PSEL(%d, 0x%08xu, 0x%08xu);" % (k, k, lo, hi))
        return "P%d" % k

    for i in range(nwords):
        body = []
        for k in range(nseg):
            lo_s = R[k] if k >= 1 else 0                         # first sample that can belong to segment k
            hi_s = R[k + 1]                                      # last sample that can belong to it
            pot = 0
            for s in range(lo_s, hi_s + 1):
                if s >> 2 == i:
                    pot |= 0xFF << (8 * (s & 3))
            if pot == 0:
                continue
            lb = k if (k >= 1 and R[k] >> 2 == i) else None       # boundary k (start of the segment) inside this word
            ub = k + 1 if R[k + 1] >> 2 == i else None            # boundary k+1 (its end) inside this word
            if lb is None and ub is None:
                expr = "X"
            elif lb is None:
                expr = "X & %s" % pfx(ub)
            elif ub is None:
                expr = "X & ~%s" % pfx(lb)
            else:
                expr = "X & ~%s & %s" % (pfx(lb), pfx(ub))
            n = acc_name(k)
            s_ = "{ const unsigned M = %s; " % expr
            if pot & 0x0000FFFF:
                s_ += "%sr = FAST_DP_LO(T.x, M, %sr); %si = FAST_DP_LO(T.z, M, %si); " % (n, n, n, n)
            if pot & 0xFFFF0000:
                s_ += "%sr = FAST_DP_HI(T.y, M, %sr); %si = FAST_DP_HI(T.w, M, %si); " % (n, n, n, n)
            s_ += "}"
            body.append(s_)
        e("{ const unsigned X = FAST_FSH(FAST_RAW(%d), FAST_RAW(%d)); const int4 T = FAST_WTAB(%d);" % (i, i + 1, i))
        for b in body:
            e(b)
        e("}")
    return lines, nwords


def main():
    fs = F(sys.argv[1]) if len(sys.argv) > 1 else F(99375000)
    fc = F(sys.argv[2]) if len(sys.argv) > 2 else F(1023000)
    d = F(sys.argv[3]) if len(sys.argv) > 3 else F(6, 100)
    nb = len(sys.argv) > 4 and sys.argv[4] == "nb"
    global RANK_BINS
    S, beta, R, theta = geometry(fs, fc, d, nb)
    nseg = len(R) - 1
    while True:       # 99.375 MHz: 512 bins; the reference's shipped 53 MHz: 1024 (closest thresholds 0.0046 sample apart)
        try:
            order, pos, thr_nom, masks, rank_lo = rank_tables(R, theta)
            break
        except AssertionError:
            if RANK_BINS >= 4096:
                raise
            RANK_BINS *= 2
    nsamp = R[nseg] + 1                # samples 0..R[nseg] (the last one is the end-boundary jitter sample)
    out = []
    w = out.append
    w("// GENERATED by gen_fast_wb.py — do not edit.  fs=%s Hz, fc=%s Hz, d=%s%s" % (fs, fc, float(d), ", narrow band (BOC(1,1) replicas only)" if nb else ""))
    w("#define FAST_NSEG %d   /* segments of a chip on which every replica is constant */" % nseg)
    w("#define FAST_NB %d     /* 1: narrow-band body (data + BOC(1,1) pilot), no BOC(6,1) class sums */" % (1 if nb else 0))
    w("#define FAST_FS_HZ %.1f" % float(fs))
    w("#define FAST_FC_HZ %.1f" % float(fc))
    w("#define FAST_D %.17g" % float(d))
    w("#define FAST_NSAMP %d" % nsamp)
    w("#define FAST_NWORDS %d" % ((nsamp + 3) // 4))
    w("#define FAST_RLAST %d" % R[nseg])
    w("#define FAST_RANK_BINS %d" % RANK_BINS)
    w("#define FAST_RANK_BITS %d" % (RANK_BINS.bit_length() - 1))
    w("#define FAST_POS_LAST %d   /* sorted position of the chip-end threshold (the last boundary) */" % pos[nseg - 1])
    w("#define FAST_SAMPLES_PER_CHIP %.17g" % float(12 * S))
    w("FAST_CONST int kFastR[%d] = {%s};" % (nseg + 1, ", ".join(map(str, R))))
    w("FAST_CONST double kFastBeta[%d] = {%s};" % (nseg + 1, ", ".join("%.17g" % float(b) for b in beta)))
    w("// nominal frac(beta_k * S): %s" % " ".join("%.3f" % float(t) for t in theta[1:]))
    w("FAST_CONST unsigned kFastThrNom[%d] = {%s};" % (nseg, ", ".join("0x%08xu" % t for t in thr_nom)))
    w("FAST_CONST unsigned char kFastPos[%d] = {%s};" % (nseg, ", ".join(map(str, pos))))
    w("FAST_CONST unsigned char kFastRankLo[%d] = {%s};" % (RANK_BINS, ", ".join(map(str, rank_lo))))
    w("FAST_CONST unsigned long long kFastMask[%d] = {%s};" % (nseg + 1, ", ".join("0x%010xull" % m for m in masks)))

    special = {(1, 0): "A1", (1, 1): "B1", (7, 0): "A7", (7, 1): "B7", (6, 1): "B6", (6, 2): "C6", (12, 1): "B12",
               (12, 2): "C12"}

    def acc_name(k):
        if nb:
            return "N%d" % k
        j, cls = k // 3 + 1, k % 3
        if (j, cls) in special:
            return special[(j, cls)]
        return "%s%d%s" % ("E" if j % 2 == 0 else "O", 1 if j <= 6 else 2, "ABC"[cls])

    names = []
    for k in range(nseg):
        n = acc_name(k)
        if n not in names:
            names.append(n)
    w("#define FAST_DECL_ACCS int " + ", ".join("%sr = 0, %si = 0" % (n, n) for n in names) + ";")

    body, nwords = emit_body(R, acc_name)
    w("#define FAST_CHIP_BODY \\")
    for ln in body:
        w("    " + ln + " \\")
    w("    /* end */")

    # ---- chip-end combination into the eight basis sums, as chains of three-input adds (IADD3)
    #   X  = H2 - H1 (second minus first half chip), SA/SB/SC = even minus odd sub-chips per segment class,
    #   W1a = A1+B1, W1b = A7+B7 (windows after the two BOC(1,1) edges), W2a = B6+C6, W2b = B12+C12 (before them)
    comb = []
    c = comb.append

    def chain(name, comp, terms):
        """name = sum of signed terms, three inputs per statement"""
        terms = [(sg, t if t.startswith("W") else t) for sg, t in terms]
        cur, k, n = None, 0, 0
        while k < len(terms):
            take = terms[k:k + (3 if cur is None else 2)]
            k += len(take)
            expr = (cur or "")
            for sg, t in take:
                v = t + comp
                expr += (" %s %s" % (sg, v)) if expr else (v if sg == "+" else "-" + v)
            last = k >= len(terms)
            dst = name + comp if last else "u%s%s%d" % (name, comp, n)
            c("const int %s = %s;" % (dst, expr))
            cur, n = dst, n + 1

    P, M = "+", "-"
    if nb:
        for comp in ("r", "i"):
            c("const int W1a%s = N0%s, W2a%s = N2%s, W1b%s = N3%s, W2b%s = N5%s;" % ((comp,) * 8))
            chain("X", comp, [(P, "N3"), (P, "N4"), (P, "N5"), (M, "N0"), (M, "N1"), (M, "N2")])
    for comp in (() if nb else ("r", "i")):
        c("const int W1a%s = A1%s + B1%s, W1b%s = A7%s + B7%s, W2a%s = B6%s + C6%s, W2b%s = B12%s + C12%s;" % ((comp,) * 12))
        chain("X", comp, [(P, "E2A"), (P, "O2A"), (P, "E2B"), (P, "O2B"), (P, "E2C"), (P, "O2C"), (P, "W1b"), (P, "W2b"),
                          (M, "E1A"), (M, "O1A"), (M, "E1B"), (M, "O1B"), (M, "E1C"), (M, "O1C"), (M, "W1a"), (M, "W2a")])
        chain("SA", comp, [(P, "E1A"), (P, "E2A"), (M, "O1A"), (M, "O2A"), (M, "A1"), (M, "A7")])
        chain("SB", comp, [(P, "E1B"), (P, "E2B"), (M, "O1B"), (M, "O2B"), (P, "B6"), (P, "B12"), (M, "B1"), (M, "B7")])
        chain("SC", comp, [(P, "E1C"), (P, "E2C"), (M, "O1C"), (M, "O2C"), (P, "C6"), (P, "C12")])
    w("#define FAST_COMBINE \\")
    for ln in comb:
        w("    " + ln + " \\")
    w("    /* end */")
    # the file may be included once per geometry in one translation unit: drop the previous geometry's macros first
    names = []
    for ln in out:
        if ln.startswith("#define "):
            names.append(ln.split()[1].split("(")[0])
    out[1:1] = ["#undef %s" % n for n in names]
    sys.stdout.write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()

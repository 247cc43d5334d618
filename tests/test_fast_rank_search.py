"""fast_chip's per-chip bookkeeping (csrc/bds_track_fast.cuh) restated in Python for the nominal B1C geometry: the
one-compare rank search over the generated tables, and the chips' sample ranges tiling an epoch's block."""
import re
import os

import numpy as np

INC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bds-3-b1c-b2a-sdr-receiver_b200", "csrc",
                   "bds_track_fast_gen.inc")
TEXT = open(INC).read()
R = [int(v) for v in re.search(r"kFastR\[37\] = \{(.*?)\}", TEXT).group(1).split(",")]
BETA = [float(v) for v in re.search(r"kFastBeta\[37\] = \{(.*?)\}", TEXT).group(1).split(",")]


RANK_LO = [int(v) for v in re.search(r"kFastRankLo\[\d+\] = \{(.*?)\}", TEXT, flags=re.S).group(1).split(",")]
POS_LAST = int(re.search(r"#define FAST_POS_LAST (\d+)", TEXT).group(1))


def tables(code_freq):
    """sorted thresholds of an epoch (fast_build_tab_lane)"""
    S = 1.0 / (12.0 * code_freq / 99.375e6)
    thr = []
    for k in range(1, 37):
        th = BETA[k] * S - R[k]
        assert 1e-6 < th < 1 - 1e-6
        thr.append(int(min(th * 4294967296.0, 4294967295.0)))
    return sorted(thr) + [0xFFFFFFFF] * 4


def search(srt, Psi, guard):
    """fast_chip: rank = rankLo[bin] + (thr[rankLo[bin]] < Psi); near = within the guard band of that threshold or of 0 / 1"""
    jlo = RANK_LO[Psi >> 23]
    t = srt[jlo]
    j = jlo + (t < Psi)
    near = ((t - Psi + guard) & 0xFFFFFFFF) <= 2 * guard or Psi <= guard or Psi >= 0xFFFFFFFF - guard
    return j, near


def test_b1c_chip_bookkeeping_tiles_the_block():
    """fast_chip's per-chip bookkeeping (first sample nc, sub-sample phase psi, rank -> jitter mask, length) restated in
    Python for whole epochs: the chips' sample ranges [nc, nc + len) tile the block of blksize samples exactly once
    (apart from chips the kernel sends to the exact path), and the mask picked through the sorted thresholds equals the
    mask from the definition (boundary sample R_k stays in the old segment iff it lies before beta_k S - psi)."""
    import math
    fs, L = 99.375e6, 10230
    for rem, code_freq in ((0.0, 1.023e6), (0.0093, 1.023e6 - 2.9), (0.0007, 1.023e6 + 3.1)):
        step = code_freq / fs
        blk = int(math.ceil((L - rem) / step))
        srt = tables(code_freq)
        pos = {v: i for i, v in enumerate(srt[:36])}
        S = 1.0 / (12.0 * step)
        thr_of_k = [int(min((BETA[k] * S - R[k]) * 4294967296.0, 4294967295.0)) for k in range(1, 37)]
        masks = [sum(1 << (k - 1) for k in range(1, 37) if pos[thr_of_k[k - 1]] >= j) for j in range(37)]
        cover = np.zeros(blk + 200, dtype=np.int64)
        n_exact = 0
        for c in range(L):
            q = (12.0 * c - 12.0 * rem) * S
            nc = int(math.floor(q)) + 1
            psi = nc - q
            Psi = int(min(psi * 4294967296.0, 4294967295.0))
            j, near = search(srt, Psi, 16)
            assert j == sum(v < Psi for v in srt[:36])
            mk = masks[j]
            ln = R[36] + (j <= POS_LAST)
            assert (j <= POS_LAST) == bool((mk >> 35) & 1)
            if near or nc < 0 or nc + ln > blk:
                n_exact += 1
                continue
            want = sum(1 << (k - 1) for k in range(1, 37) if R[k] < BETA[k] * S - psi)
            assert mk == want, (c, hex(mk), hex(want))
            cover[nc:nc + ln] += 1
        # rem = 0 at exactly the nominal code rate puts (12 c + beta_k) S on an integer for ~1 % of the chips (exact ties,
        # all sent through the exact path); any real code-rate offset removes them
        assert n_exact <= (200 if rem == 0.0 else 3), n_exact
        holes = np.flatnonzero(cover[:blk] != 1)
        assert cover[blk:].sum() == 0 and holes.size <= 98 * n_exact, (holes[:5], n_exact)
        assert np.all(cover[:blk] <= 1)

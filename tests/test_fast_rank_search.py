"""The two rank searches of fast_chip (csrc/bds_track_fast.cuh) pick the same jitter mask.

Default build: bin start + four dependent refinement steps + distances to the neighbouring thresholds.
-DBDS_FAST_BINREC=1: one 8-byte record per pair of bins (the at most two thresholds inside) + a guard band at the pair's
edges.  Restated here in Python for the nominal B1C geometry under Doppler: the rank j must be identical for every
sub-sample phase Psi, and every Psi the default sends to the exact path must go there in the record variant too."""
import re
import os

import numpy as np

INC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bds-3-b1c-b2a-sdr-receiver_b200", "csrc",
                   "bds_track_fast_gen.inc")
TEXT = open(INC).read()
R = [int(v) for v in re.search(r"kFastR\[37\] = \{(.*?)\}", TEXT).group(1).split(",")]
BETA = [float(v) for v in re.search(r"kFastBeta\[37\] = \{(.*?)\}", TEXT).group(1).split(",")]


def tables(code_freq):
    S = 1.0 / (12.0 * code_freq / 99.375e6)
    thr = []
    for k in range(1, 37):
        th = BETA[k] * S - R[k]
        assert 1e-6 < th < 1 - 1e-6
        thr.append(int(min(th * 4294967296.0, 4294967295.0)))
    srt = sorted(thr) + [0xFFFFFFFF] * 4
    bins = [sum((v >> 25) < t for v in thr) for t in range(129)]
    assert all(sum((v >> 25) == t for v in thr) <= 4 for t in range(129))
    rec = []
    for b in range(64):
        s0, s1 = bins[2 * b], bins[2 * b + 2]
        assert s1 - s0 <= 2
        rec.append((srt[s0] if s0 < s1 else 0xFFFFFFFF, srt[s0 + 1] if s0 + 1 < s1 else 0xFFFFFFFF))
    return srt, bins, rec


def search_default(srt, bins, Psi, guard):
    j = bins[Psi >> 25]
    for _ in range(4):
        j += srt[j] < Psi
    below = Psi - srt[j - 1] if j > 0 else Psi
    above = srt[j] - Psi if j < 36 else 0xFFFFFFFF - Psi
    return j, below <= guard or above <= guard or Psi >= 0xFFFFFFFF - guard


def search_record(bins, rec, Psi, guard):
    u32 = lambda v: v & 0xFFFFFFFF
    rx, ry = rec[Psi >> 26]
    j = bins[(Psi >> 26) * 2] + (rx < Psi) + (ry < Psi)
    eg, low = min(guard, 4096), Psi & 0x3FFFFFF
    exact = (u32(rx - Psi + guard) <= 2 * guard or u32(ry - Psi + guard) <= 2 * guard or low <= eg
             or low >= 0x3FFFFFF - eg or Psi >= 0xFFFFFFFF - guard)
    return j, exact


def test_record_search_equals_refinement_search():
    rng = np.random.default_rng(11)
    for code_freq in (1.023e6, 1.023e6 - 3.2, 1.023e6 + 3.3, 1.023e6 + 0.017):
        srt, bins, rec = tables(code_freq)
        probes = [int(v) for v in rng.integers(0, 1 << 32, size=4000)]
        for t in srt[:36]:                                     # around every threshold and every bin edge
            probes += [max(0, min(0xFFFFFFFF, t + dlt)) for dlt in (-4097, -17, -16, -1, 0, 1, 16, 17, 4097)]
        for b in range(1, 128):
            probes += [(b << 25) + dlt for dlt in (-17, -1, 0, 1, 17)]
        probes += [0, 1, 15, 16, 17, 0xFFFFFFFF, 0xFFFFFFFF - 16, 0xFFFFFFFF - 17]
        for guard in (16, 1 << 24):
            n_def = n_rec = 0
            for Psi in probes:
                jd, ed = search_default(srt, bins, Psi, guard)
                jr, er = search_record(bins, rec, Psi, guard)
                assert jd == jr, (code_freq, hex(Psi))
                if guard <= 4096:                              # the wide guard band is a test hook: any subset will do
                    assert er or not ed, (code_freq, hex(Psi), guard)
                n_def += ed
                n_rec += er
            if guard == 16:
                assert n_rec <= n_def + 5 * 127 + 8          # extra exact chips only at the 127 bin edges probed


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
        srt, bins, rec = tables(code_freq)
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
            j, near = search_default(srt, bins, Psi, 16)
            mk = masks[j]
            ln = R[36] + ((mk >> 35) & 1)
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

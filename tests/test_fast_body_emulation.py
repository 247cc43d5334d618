"""CPU emulation of the generated chip body of the chip-synchronous correlator.

`csrc/bds_track_fast_gen.inc` (written by `csrc/gen_fast_wb.py`) is straight-line C of ~900 statements that only
runs on the GPU.  Here the same text is executed statement by statement in Python (the macros FAST_RAW / FAST_FSH /
FAST_WTAB / FAST_DP_LO / FAST_DP_HI / FAST_SELU get Python definitions with the semantics of the CUDA intrinsics) and
checked against the definition it implements:
  1. every one of the 20 segment-class accumulators equals the per-sample sum over the samples of its segments, for
     random samples, carrier tables, alignments and jitter masks (integer-exact);
  2. the nine basis sums + the chip-sign combination of `fast_chip` reproduce the nine replicas of
     WB_tracking.m:289-372 (data / pilot BOC(1,1), pilot BOC(6,1) x E/P/L) evaluated sample by sample.
A change of the generator that breaks the arithmetic fails here, on the CPU, before any GPU time is spent."""
import os
import re

import numpy as np
import pytest

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bds-3-b1c-b2a-sdr-receiver_b200", "csrc")
INC = os.path.join(CSRC, "bds_track_fast_gen.inc")
INC_B2A = os.path.join(CSRC, "bds_track_fast_b2a_gen.inc")   # gen_fast_b2a.py: ten B2a chips per thread


def _macro(text, name):
    """body of `#define name ...` with line continuations joined"""
    m = re.search(r"#define %s\b(.*?)(?<!\\)\n" % name, text, flags=re.S)
    return m.group(1).replace("\\\n", "\n")


def _s32(v):
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v & 0x80000000 else v


def _s16(v):
    v &= 0xFFFF
    return v - (1 << 16) if v & 0x8000 else v


def _s8(v):
    v &= 0xFF
    return v - 256 if v & 0x80 else v


def dp2a(lo, a, b, c):
    """__dp2a_lo / __dp2a_hi (signed): c + a.lo16 * b.byte[0|2] + a.hi16 * b.byte[1|3]"""
    sh = 0 if lo else 16
    return _s32(c + _s16(a) * _s8(b >> sh) + _s16(a >> 16) * _s8(b >> (sh + 8)))


class _Row:
    def __init__(self, v):
        self.x, self.y, self.z, self.w = v


class Gen:
    def __init__(self, inc=INC, prefix="FAST", tab="kFast"):
        text = open(inc).read()
        self.prefix = prefix
        self.R = [int(v) for v in re.search(tab + r"R\[\d+\] = \{(.*?)\}", text).group(1).split(",")]
        self.beta = [float(v) for v in re.search(tab + r"Beta\[\d+\] = \{(.*?)\}", text).group(1).split(",")]
        self.nb = len(self.R) - 1                                     # boundaries = segments per thread unit
        self.nwords = int(re.search(r"#define %s_NWORDS (\d+)" % prefix, text).group(1))
        self.names = re.findall(r"(\w+)r = 0", _macro(text, prefix + "_DECL_ACCS"))
        self.body = self._to_python(_macro(text, prefix + "_CHIP_BODY"))
        self.combine = self._to_python(_macro(text, prefix + "_COMBINE"))

    @staticmethod
    def _to_python(c):
        c = re.sub(r"/\*.*?\*/", "", c)
        c = re.sub(r"\bconst (unsigned|uint4|int4|int)\b", "", c)
        c = re.sub(r"0x([0-9a-fA-F]+)u", r"0x\1", c)
        c = c.replace("{", "\n").replace("}", "\n").replace(";", "\n")
        for f, i in (("x", 0), ("y", 1), ("z", 2), ("w", 3)):
            c = c.replace("T." + f, "T[%d]" % i)
        out = []
        for ln in c.split("\n"):
            ln = ln.strip()
            if not ln:
                continue
            if "," in ln and "=" in ln and "(" not in ln.split("=")[0] and ln.count("=") > 1:
                # "tO1Ar = a + b, tO1Br = c + d" -> separate statements
                out += [p.strip() for p in re.split(r",\s*(?=\w+\s*=)", ln)]
            else:
                out.append(ln)
        return "\n".join(out)

    def _prefix(self, k, mk):
        b = self.R[k] & 3
        return (((1 << (8 * b)) - 1) | ((0xFF << (8 * b)) if (mk >> (k - 1)) & 1 else 0)) & 0xFFFFFFFF

    def run(self, raw_words, table, sh, mk, extra=None):
        """one chip: raw_words = 32-bit words starting at the aligned word of the first sample, sh = 8 * (offset & 3),
        table[i] = (wr01, wr23, wi01, wi23) packed int16 pairs, mk = jitter mask (bit k-1: boundary k)"""
        env = {n + c: 0 for n in self.names for c in "ri"}
        p = self.prefix
        env.update({
            p + "_RAW": lambda i: raw_words[i],
            p + "_FSH": lambda lo, hi: ((lo | (hi << 32)) >> sh) & 0xFFFFFFFF,
            p + "_WTAB": lambda i: table[i],
            p + "_DP_LO": lambda a, b, c: dp2a(True, a, b, c),
            p + "_DP_HI": lambda a, b, c: dp2a(False, a, b, c),
            p + "_SELU": lambda k, v: v if (mk >> (k - 1)) & 1 else 0,
            # B1C body: prefix byte mask of boundary k from its bit of the decision mask
            p + "_PSEL": lambda k, lo, hi: hi if (mk >> (k - 1)) & 1 else lo})
        env.update(extra or {})
        exec(self.body, env)
        exec(self.combine, env)
        return env


GEN = Gen()


def _class_of_segment(k):
    """accumulator name of segment k (gen_fast_wb.py acc_name)"""
    special = {(1, 0): "A1", (1, 1): "B1", (7, 0): "A7", (7, 1): "B7", (6, 1): "B6", (6, 2): "C6", (12, 1): "B12",
               (12, 2): "C12"}
    j, cls = k // 3 + 1, k % 3
    return special.get((j, cls), "%s%d%s" % ("E" if j % 2 == 0 else "O", 1 if j <= 6 else 2, "ABC"[cls]))


def _random_chip(rng, off, gen=None):
    gen = gen or GEN
    x = rng.integers(-127, 128, size=4 * (gen.nwords + 2)).astype(np.int64)      # int8 samples from the aligned word on
    wr = rng.integers(-32767, 32768, size=4 * (gen.nwords + 1)).astype(np.int64)
    wi = rng.integers(-32767, 32768, size=4 * (gen.nwords + 1)).astype(np.int64)
    b = (x & 0xFF).astype(np.uint64)
    words = [int(b[4 * i] | (b[4 * i + 1] << np.uint64(8)) | (b[4 * i + 2] << np.uint64(16)) | (b[4 * i + 3] << np.uint64(24)))
             for i in range(gen.nwords + 2)]
    pk = lambda a, c: int((int(a) & 0xFFFF) | ((int(c) & 0xFFFF) << 16))
    table = [(pk(wr[4 * i], wr[4 * i + 1]), pk(wr[4 * i + 2], wr[4 * i + 3]), pk(wi[4 * i], wi[4 * i + 1]),
              pk(wi[4 * i + 2], wi[4 * i + 3])) for i in range(gen.nwords + 1)]
    xs = x[off:]                                                                # sample r of the chip = x[off + r]
    return words, table, xs, wr, wi


@pytest.mark.parametrize("seed", range(6))
def test_generated_body_equals_per_sample_segment_sums(seed):
    rng = np.random.default_rng(seed)
    R = GEN.R
    for _ in range(12):
        off = int(rng.integers(0, 4))
        mk = int(rng.integers(0, 1 << 36))
        words, table, xs, wr, wi = _random_chip(rng, off)
        env = GEN.run(words, table, 8 * off, mk)
        want = {n + c: 0 for n in GEN.names for c in "ri"}
        last = R[36] + ((mk >> 35) & 1)                      # sample R36 is mine only if bit 35 is set
        for r in range(last):
            k = 0                                            # segment of sample r: boundary sample R_k stays in the OLD
            for kk in range(1, 36):                          # segment (k-1) iff mask bit k-1 is set
                if r > R[kk] or (r == R[kk] and not (mk >> (kk - 1)) & 1):
                    k = kk
            name = _class_of_segment(k)
            want[name + "r"] += int(xs[r]) * int(wr[r])
            want[name + "i"] += int(xs[r]) * int(wi[r])
        for n in want:
            assert env[n] == want[n], (n, off, hex(mk))


def test_basis_sums_and_chip_signs_reproduce_the_nine_replicas():
    """fast_chip's combination (bds_track_fast.cuh) of the nine basis sums with the chip signs == the per-sample
    definition of the nine +-1 replicas for a chip whose neighbours carry arbitrary signs."""
    rng = np.random.default_rng(42)
    S = 99.375e6 / (12 * 1.023e6)                            # samples per BOC(6,1) sub-chip
    delta = 12 * 0.06                                        # early/late spacing in sub-chips
    for _ in range(40):
        psi = float(rng.uniform(0.02, 0.98))                 # first sample of the chip is psi samples after the chip start
        off = int(rng.integers(0, 4))
        words, table, xs, wr, wi = _random_chip(rng, off)
        # jitter mask from the definition: boundary k at sample position beta_k * S - psi (relative to sample 0);
        # sample R_k belongs to the old segment iff it lies before the boundary
        mk = 0
        for k in range(1, 37):
            if GEN.R[k] < GEN.beta[k] * S - psi:
                mk |= 1 << (k - 1)
        env = GEN.run(words, table, 8 * off, mk)
        z = lambda r: complex(int(xs[r]) * int(wr[r]), int(xs[r]) * int(wi[r]))
        cd, cdp, cdn, cp, cpp, cpn = (int(v) for v in rng.choice([-1, 1], size=6))
        n = GEN.R[36] + ((mk >> 35) & 1)
        want = {}
        for fam, (c0, cprev, cnext) in (("d", (cd, cdp, cdn)), ("p", (cp, cpp, cpn))):
            # WB_tracking.m:289-317: the "early" replica reads the code at remCodePhase - d, the "late" one at + d
            for name, shift in (("E", -delta), ("P", 0.0), ("L", +delta)):
                acc11 = acc61 = 0
                for r in range(n):
                    v = (r + psi) / S + shift                # sub-chip coordinate seen by this replica
                    chip = c0 if 0 <= v < 12 else (cprev if v < 0 else cnext)
                    vv = v % 12
                    s11 = -1 if vv < 6 else 1                # chip c -> [-c, +c]   (generateDataBOC11.m:85-90)
                    s61 = 1 if int(vv) % 2 else -1           # chip c -> (-1)^ii c, ii = 1..12 (generatePilotBOC61.m:89-96)
                    acc11 += chip * s11 * z(r)
                    acc61 += chip * s61 * z(r)
                want[(fam, name)] = acc11
                if fam == "p":
                    want[("p61", name)] = acc61
        g = lambda nm: complex(env[nm + "r"], env[nm + "i"])
        SA, SB, SC, X = g("SA"), g("SB"), g("SC"), g("X")
        W1a, W1b, W2a, W2b = g("W1a"), g("W1b"), g("W2a"), g("W2b")
        XE = X + W1a - 2 * W1b
        XL = X + 2 * W2a - W2b
        got = {("d", "P"): cd * X, ("d", "E"): cd * XE + cdp * W1a, ("d", "L"): cd * XL - cdn * W2b,
               ("p", "P"): cp * X, ("p", "E"): cp * XE + cpp * W1a, ("p", "L"): cp * XL - cpn * W2b,
               ("p61", "P"): cp * (SA + SB + SC), ("p61", "E"): cp * (SC - SA - SB) + (cpp - cp) * W1a,
               ("p61", "L"): cp * (SA - SB - SC) + (cp - cpn) * W2b}
        for key in got:
            assert got[key] == want[key], (key, psi, off)


def test_generated_rank_tables_follow_from_the_geometry():
    """kFastMask / kFastThrNom / kFastPos / kFastRankLo of the generated file against their definitions, and the
    (lo, hi) constants of every FAST_PSEL against the prefix-mask definition"""
    text = open(INC).read()
    arr = lambda name: [int(v.strip().rstrip("ul"), 0) for v in re.search(name + r"\[\d+\] = \{(.*?)\}", text, flags=re.S).group(1).split(",")]
    thr_nom, pos, rank_lo = arr("kFastThrNom"), arr("kFastPos"), arr("kFastRankLo")
    S = 99.375e6 / (12 * 1.023e6)
    theta = [GEN.beta[k] * S - GEN.R[k] for k in range(37)]
    order = sorted(range(1, 37), key=lambda k: theta[k])
    assert [pos[k - 1] for k in order] == list(range(36))
    assert all(abs(thr_nom[s] - theta[k] * 2 ** 32) < 64 for s, k in enumerate(order))
    bins = len(rank_lo)
    w = (1 << 32) // bins
    assert all(rank_lo[b] == sum(t < (b - 1) * w for t in thr_nom) for b in range(bins))
    assert min(b - a for a, b in zip(thr_nom, thr_nom[1:])) > 3 * w        # one threshold per three-bin window
    masks = arr("kFastMask")
    assert masks == [sum(1 << (k - 1) for k in range(1, 37) if pos[k - 1] >= j) for j in range(37)]   # old <=> threshold k not below Psi
    sel = {int(k): (int(lo, 16), int(hi, 16)) for k, lo, hi in re.findall(r"FAST_PSEL\((\d+), 0x([0-9a-f]+)u, 0x([0-9a-f]+)u\)", text)}
    assert sorted(sel) == list(range(1, 37))
    for k, (lo, hi) in sel.items():
        assert lo == GEN._prefix(k, 0) and hi == GEN._prefix(k, 1 << (k - 1))
    assert int(re.search(r"#define FAST_POS_LAST (\d+)", text).group(1)) == pos[35]


# ---- B2a: ten chips per thread, 20 half-chip segments (gen_fast_b2a.py; not wired into a kernel yet) ------------
GENB = Gen(INC_B2A, "FASTB", "kFastb")


def test_b2a_generator_output_is_current():
    """the committed .inc files are what the generators write"""
    import subprocess
    import sys
    for gen, inc in (("gen_fast_wb.py", INC), ("gen_fast_b2a.py", INC_B2A)):
        out = subprocess.run([sys.executable, os.path.join(CSRC, gen)], capture_output=True, text=True, check=True).stdout
        assert out == open(inc).read(), gen


@pytest.mark.parametrize("seed", range(4))
def test_b2a_body_equals_per_sample_segment_sums(seed):
    rng = np.random.default_rng(100 + seed)
    R, nb = GENB.R, GENB.nb
    assert nb == 20 and GENB.names == ["H%d" % k for k in range(20)]
    for _ in range(12):
        off = int(rng.integers(0, 4))
        mk = int(rng.integers(0, 1 << nb))
        words, table, xs, wr, wi = _random_chip(rng, off, GENB)
        env = GENB.run(words, table, 8 * off, mk, {"FASTB_CD": lambda j: 1, "FASTB_CP": lambda j: 1})
        want = {n + c: 0 for n in GENB.names for c in "ri"}
        last = R[nb] + ((mk >> (nb - 1)) & 1)
        for r in range(last):
            k = 0
            for kk in range(1, nb):
                if r > R[kk] or (r == R[kk] and not (mk >> (kk - 1)) & 1):
                    k = kk
            want["H%dr" % k] += int(xs[r]) * int(wr[r])
            want["H%di" % k] += int(xs[r]) * int(wi[r])
        for n in want:
            assert env[n] == want[n], (n, off, hex(mk))


def test_b2a_combination_reproduces_the_six_replicas():
    """FASTB_COMBINE with the chip signs == early / prompt / late of the data and the pilot code evaluated sample by
    sample with the reference's indexing (tracking.m:262-296: code(ceil(t -+ d) + 1) on the table extended by one chip
    on either side, d = 0.5)."""
    import math
    rng = np.random.default_rng(7)
    S = 99.375e6 / (2 * 10.23e6)                              # samples per half chip
    nb = GENB.nb
    for _ in range(40):
        psi = float(rng.uniform(0.02, 0.98))
        off = int(rng.integers(0, 4))
        words, table, xs, wr, wi = _random_chip(rng, off, GENB)
        mk = 0
        for k in range(1, nb + 1):
            if GENB.R[k] < GENB.beta[k] * S - psi:
                mk |= 1 << (k - 1)
        cd = {j: int(v) for j, v in zip(range(-1, 11), rng.choice([-1, 1], size=12))}
        cp = {j: int(v) for j, v in zip(range(-1, 11), rng.choice([-1, 1], size=12))}
        env = GENB.run(words, table, 8 * off, mk, {"FASTB_CD": lambda j: cd[j], "FASTB_CP": lambda j: cp[j]})
        n = GENB.R[nb] + ((mk >> (nb - 1)) & 1)
        for fam, code in (("D", cd), ("P", cp)):
            for name, shift in (("E", -0.5), ("P", 0.0), ("L", 0.5)):
                acc = 0
                for r in range(n):
                    t = (r + psi) / (2 * S) + shift           # code phase in chips from the start of the unit
                    acc += code[math.ceil(t) - 1] * complex(int(xs[r]) * int(wr[r]), int(xs[r]) * int(wi[r]))
                got = complex(env[fam + name + "r"], env[fam + name + "i"])
                assert got == acc, (fam, name, psi, off)

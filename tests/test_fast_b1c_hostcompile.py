"""fast_chip (csrc/bds_track_fast.cuh) compiled for the host, in its default form and in the opt-in variants
(-DBDS_FAST_F32X2=1, -DBDS_FAST_BINREC=1), run over a whole B1C epoch against the float64 oracle.

Same method as tests/test_fast_b2a_hostcompile.py: the header is taken as it is, the functions that only exist on the
GPU (inline PTX, warp-cooperative table build, TMA helpers) are cut out and replaced by host definitions, the rest
(fast_chip with the generated body, the chip-sign combination, the exact path, the accumulator helpers) is compiled with
g++.  The per-epoch table (thresholds, rank masks, bins, per-bin records) is built in Python the way
fast_build_tab_warp builds it.  18 sums must equal WB_tracking.m:289-380 as restated by the oracle within 1e-4 - for every
build flag combination, so a variant is known to be arithmetically right before it is ever given GPU time."""
import ctypes as C
import math
import os
import re
import subprocess

import numpy as np
import pytest

import bds_oracle as O
from test_fast_b2a_hostcompile import SHIM, EpochParams, _block, _pack_bits
from test_fast_rank_search import BETA, R

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bds-3-b1c-b2a-sdr-receiver_b200", "csrc")
FS, L = 99.375e6, 10230
TWO32, TWO64 = 1 << 32, 1 << 64

GPU_ONLY = [r"__device__ void fast_build_tab_warp", r"__device__ __forceinline__ unsigned smem_u32",
            r"__device__ __forceinline__ void mbar_init", r"__device__ __forceinline__ void tma_load_1d",
            r"__device__ __forceinline__ void mbar_wait", r"template <int B>\s*__device__ __forceinline__ int sext_byte",
            r"template <int BIT>\s*__device__ __forceinline__ unsigned sel_bit_u",
            r"template <int BIT>\s*__device__ __forceinline__ int sel_bit\b",
            r"__device__ __forceinline__ f2_t f2_pk", r"__device__ __forceinline__ void f2_unpk",
            r"__device__ __forceinline__ f2_t f2_fma", r"__device__ __forceinline__ f2_t f2_mul",
            r"__device__ __forceinline__ f2_t f2_add", r"__device__ __forceinline__ f2_t f2_sub"]

F2_SHIM = r"""
typedef unsigned long long f2_t;
static inline f2_t f2_pk(float a, float b) { uint32_t x, y; std::memcpy(&x, &a, 4); std::memcpy(&y, &b, 4); return ((f2_t)y << 32) | x; }
static inline void f2_unpk(f2_t v, float& a, float& b) { uint32_t x = (uint32_t)v, y = (uint32_t)(v >> 32); std::memcpy(&a, &x, 4); std::memcpy(&b, &y, 4); }
#define F2_OP(name, expr) static inline f2_t name { float ax, ay, bx, by, cx = 0, cy = 0; f2_unpk(a, ax, ay); f2_unpk(b, bx, by); expr }
static inline f2_t f2_fma(f2_t a, f2_t b, f2_t c) { float ax, ay, bx, by, cx, cy; f2_unpk(a, ax, ay); f2_unpk(b, bx, by); f2_unpk(c, cx, cy); return f2_pk(std::fmaf(ax, bx, cx), std::fmaf(ay, by, cy)); }
static inline f2_t f2_mul(f2_t a, f2_t b) { float ax, ay, bx, by; f2_unpk(a, ax, ay); f2_unpk(b, bx, by); return f2_pk(ax * bx, ay * by); }
static inline f2_t f2_add(f2_t a, f2_t b) { float ax, ay, bx, by; f2_unpk(a, ax, ay); f2_unpk(b, bx, by); return f2_pk(ax + bx, ay + by); }
static inline f2_t f2_sub(f2_t a, f2_t b) { float ax, ay, bx, by; f2_unpk(a, ax, ay); f2_unpk(b, bx, by); return f2_pk(ax - bx, ay - by); }
struct uint2 { unsigned x, y; };
static inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r; r.x = x; r.y = y; return r; }
#define BDS_TRK_B1C_WB 0
#define BDS_TRK_B1C_NB 1
"""

DRIVER = r"""
extern "C" int run_epoch(const FastTab* tab, const EpochParams* p, const uint32_t* bits, const unsigned char* tile,
                         long long tileBase, int tileBytes, long long B0, const int8_t* xblk, unsigned guard, double* sums18) {
    long long q8[kNSum];
    for (int i = 0; i < kNSum; ++i) q8[i] = 0;
    int nExact = 0;
    if (p->rem == 0.0) {   // trk_fw_kernel: first pass of an epoch with remCodePhase == 0
        float a0[kNSum];
        for (int i = 0; i < kNSum; ++i) a0[i] = 0.f;
        ExactCtx ex;
        make_exact_ctx(*p, FAST_D, FAST_FS_HZ, ex);
        fast_exact_range(ex, xblk, bits, bits + kPackedWordsDev, 0, 0, -100, 0, a0);
        for (int i = 0; i < kNSum; ++i) q8[i] += __float2int_rn(a0[i] * 256.f);
    }
    for (int c = 0; c < 10230; ++c) {
        fast_acc_t acc[kFastAccN];
        fast_acc_zero(acc);
        nExact += fast_chip(*tab, *p, bits, bits + kPackedWordsDev, tile, tileBase, tileBytes, B0, xblk, FAST_D, FAST_FS_HZ, c,
                            guard, acc);
        for (int i = 0; i < kNSum; ++i) q8[i] += __float2int_rn(fast_acc_get(acc, i) * 256.f);
    }
    for (int i = 0; i < kNSum; ++i) sums18[i] = (double)q8[i] / 256.0;
    return nExact;
}
extern "C" int sizeof_tab() { return (int)sizeof(FastTab); }
extern "C" int offsetof_u0() { return (int)offsetof(FastTab, u0); }
"""


def _cut(text, pat):
    m = re.search(pat, text)
    assert m, pat
    blk = _block(text[m.start():], pat)
    return text[:m.start()] + text[m.start() + len(blk) - 1:]


def _build(tmp, flags):
    trk = open(os.path.join(CSRC, "bds_track.cuh")).read()
    fast = open(os.path.join(CSRC, "bds_track_fast.cuh")).read()
    inc = open(os.path.join(CSRC, "bds_track_fast_gen.inc")).read().replace("static __constant__", "static const")
    for pat in GPU_ONLY:
        fast = _cut(fast, pat)
    fast = fast.replace("#pragma once", "").replace('#include "bds_track.cuh"', "")
    fast = fast.replace('#include "bds_track_fast_gen.inc"', inc)
    fast = re.sub(r"typedef unsigned long long f2_t;", "", fast)
    fast = re.sub(r"namespace bds \{", "", fast, count=1)
    fast = fast[:fast.rindex("}  // namespace bds")]
    parts = [SHIM.replace("#define BDS_TRK_B2A 2", ""), "#include <cstddef>\n", F2_SHIM,
             "constexpr int kNSum = 18;\nconstexpr int kPackedWordsDev = 320;\n",
             _block(trk, r"__host__ __device__ constexpr int sum_idx"),
             "enum { EPL_E = 0, EPL_P = 1, EPL_L = 2 };\n",
             _block(trk, r"struct EpochParams \{"),
             fast, DRIVER]
    src = tmp / "b1c_host.cpp"
    src.write_text("\n".join(parts))
    so = tmp / "b1c_host.so"
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-shared", "-fPIC", *flags, "-o", str(so), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    lib = C.CDLL(str(so))
    lib.run_epoch.restype = C.c_int
    return lib


def build_tab_bytes(lib, rem, step, carrFreq, remCarr, binrec):
    """fast_build_tab_warp, as bytes laid out like FastTab"""
    S = 1.0 / (12.0 * step)
    r = carrFreq / FS
    r -= math.floor(r)
    dphi = int(round(r * TWO64)) % TWO64
    r0 = remCarr / 6.283185307179586476925286766559
    r0 -= math.floor(r0)
    phi0 = int(round(r0 * TWO64)) % TWO64
    nw = 26
    w = np.zeros(nw * 8, dtype=np.int16)
    for t in range(4 * nw):
        ph = (t * dphi) % TWO64
        hi = ph >> 32
        a = (hi - TWO32 if hi & 0x80000000 else hi) * 4.656612873077392578125e-10 * math.pi
        w[(t >> 2) * 8 + (t & 3)] = int(round(math.cos(a) * 32767.0))
        w[(t >> 2) * 8 + 4 + (t & 3)] = int(round(-math.sin(a) * 32767.0))
    thr = []
    for k in range(1, 37):
        th = BETA[k] * S - R[k]
        assert 1e-6 < th < 1 - 1e-6
        thr.append(int(min(th * 4294967296.0, 4294967295.0)))
    pos = [sum((thr[j] < v) or (thr[j] == v and j < t) for j in range(36)) for t, v in enumerate(thr)]
    srt = np.full(40, 0xFFFFFFFF, dtype=np.uint32)
    for t, v in enumerate(thr):
        srt[pos[t]] = v
    mask = np.zeros(40, dtype=np.uint64)
    for j in range(37):
        mask[j] = sum(1 << (k - 1) for k in range(1, 37) if pos[k - 1] >= j)
    bins = np.zeros(144, dtype=np.uint8)
    for t in range(129):
        bins[t] = sum((v >> 25) < t for v in thr)
    posbin = np.zeros(80, dtype=np.uint8)
    parts = [w.tobytes(), srt.tobytes(), mask.tobytes(), bins.tobytes(), posbin.tobytes()]
    if binrec:
        rec = np.full((64, 2), 0xFFFFFFFF, dtype=np.uint32)
        for b in range(64):
            s0, s1 = int(bins[2 * b]), int(bins[2 * b + 2])
            assert s1 - s0 <= 2
            for i in range(s1 - s0):
                rec[b, i] = srt[s0 + i]
        parts.append(rec.tobytes())
    head = b"".join(parts)
    assert len(head) == lib.offsetof_u0(), (len(head), lib.offsetof_u0())
    tail = np.array([12.0 * rem, 12.0 * step, S], dtype=np.float64).tobytes() + np.array([dphi, phi0], dtype=np.uint64).tobytes() + \
        np.array([1, 0, 0, 0], dtype=np.int32).tobytes()
    buf = head + tail
    return buf + b"\0" * (lib.sizeof_tab() - len(buf))


def make_epoch(rng, prn, rem, codeFreq, carrFreq, remCarr):
    """one block of the SURVEY 8(d) B1C signal model + noise, int8"""
    step = codeFreq / FS
    blk = int(math.ceil((L - rem) / step))
    s = O.initSettings_B1C(samplingFreq=FS, pilotTRKflag=2)
    codes = O.make_track_codes("WB", s, prn)
    k = np.arange(blk)
    t = rem + k * step
    i2 = np.ceil(t * 2).astype(np.int64)
    i12 = np.ceil(t * 12).astype(np.int64)
    th = carrFreq * 2.0 * np.pi * (k / FS) + remCarr
    sig = math.sqrt(11 / 44) * codes["data"][i2] * np.cos(th) - (math.sqrt(29 / 44) * codes["pilot"][i2] * np.sin(th)
                                                                    + math.sqrt(4 / 44) * codes["pilot61"][i12] * np.cos(th))
    x = np.clip(np.rint(24.0 * rng.standard_normal(blk) + 30.0 * sig), -127, 127).astype(np.int8)
    return s, codes, x, step


NAMES = [f"{fam}_{iq}_{epl}" for fam in ("d", "p", "p61") for epl in ("E", "P", "L") for iq in ("I", "Q")]
VARIANTS = {"default": [], "f32x2": ["-DBDS_FAST_F32X2=1"], "binrec": ["-DBDS_FAST_BINREC=1"],
            "f32x2+binrec": ["-DBDS_FAST_F32X2=1", "-DBDS_FAST_BINREC=1"]}


@pytest.mark.parametrize("variant", list(VARIANTS))
def test_fast_chip_source_on_host_equals_oracle(variant, tmp_path):
    lib = _build(tmp_path, VARIANTS[variant])
    rng = np.random.default_rng(33)
    prn = 19
    for rem, cf, fc_, rc, B0, guard in ((0.0, 1.023e6 - 2.7, 14.58e6 + 1830.0, 0.0, 7, 16),
                                        (0.0061, 1.023e6 + 1.9, 14.58e6 - 3920.0, 4.2, 993750 * 2 + 13, 16),
                                        (0.0033, 1.023e6 - 0.4, 14.58e6 + 55.0, 2.2, 48, 1 << 24)):
        s, codes, x, step = make_epoch(rng, prn, rem, cf, fc_, rc)
        tab = build_tab_bytes(lib, rem, step, fc_, rc, "BINREC" in " ".join(VARIANTS[variant]))
        p = EpochParams(pos=B0, blksize=x.size, pad=0, rem=rem, step=step, carrFreq=fc_, remCarr=rc)
        bits = np.concatenate([_pack_bits(O.b1c_data_primary(prn)), _pack_bits(O.b1c_pilot_primary(prn))])
        tileBase = B0 & ~15
        tile = np.zeros((B0 - tileBase) + x.size + 256, dtype=np.int8)
        tile[B0 - tileBase:B0 - tileBase + x.size] = x
        sums = np.zeros(18)
        n_exact = lib.run_epoch(tab, C.byref(p), bits.ctypes.data_as(C.c_void_p), tile.ctypes.data_as(C.c_void_p),
                                C.c_longlong(tileBase), C.c_int(tile.size), C.c_longlong(B0), x.ctypes.data_as(C.c_void_p),
                                C.c_uint(guard), sums.ctypes.data_as(C.c_void_p))
        if guard == 16:
            assert n_exact <= 4, n_exact
        else:
            assert 0.1 < n_exact / L < 0.6, n_exact
        ref, _, _ = O.correlate_epoch("WB", s, x.astype(np.float64), codes, rem, step, fc_, rc)
        got = dict(zip(NAMES, sums))
        for fam in ("d", "p", "p61"):
            scale = max(abs(ref[f"{fam}_I_P"]), abs(ref[f"{fam}_Q_P"]))
            for nm in "EPL":
                for iq in "IQ":
                    k = f"{fam}_{iq}_{nm}"
                    assert abs(got[k] - ref[k]) <= 1e-4 * scale, (variant, k, got[k], ref[k], scale)

"""csrc/bds_track_fast.cuh compiled for the host and run over whole B1C epochs against the float64 oracle.

The header is taken as it is (its CUDA-only helpers - inline PTX, TMA, the warp vote - sit behind __CUDACC__); the CUDA
intrinsics it calls get host definitions; `bds_track.cuh` is replaced by a stub with the three definitions it needs.
What runs is the device source itself: the per-epoch table builder (fast_build_tab_lane for 32 lanes), the generated
chip body with the per-rank prefix masks, the one-compare rank search, the chip-sign combination on packed pairs and
the exact per-sample path.  18 sums must equal WB_tracking.m:289-380 as restated by the oracle within 1e-4, so a change
of the kernel's arithmetic is known to be right before it is ever given GPU time."""
import ctypes as C
import math
import os
import re
import subprocess

import numpy as np
import pytest

import bds_oracle as O
from test_fast_b2a_hostcompile import SHIM, EpochParams, _block, _pack_bits

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bds-3-b1c-b2a-sdr-receiver_b200", "csrc")
L = 10230
# chip-body geometries of the library (bds_track.cu): namespace, sampling rate, generated file
GEOMS = {"g99": (99.375e6, "bds_track_fast_gen.inc"),      # BASELINE
         "g53": (53e6, "bds_track_fast_gen_53.inc"),        # the reference's shipped B1C setting (B1C/initSettings.m:57)
         "g99n": (99.375e6, "bds_track_fast_gen_nb.inc"),   # narrow-band bodies (NB_tracking.m): six segments per chip
         "g53n": (53e6, "bds_track_fast_gen_53_nb.inc")}
TWO32, TWO64 = 1 << 32, 1 << 64

F2_SHIM = r"""
typedef unsigned long long f2_t;
static inline f2_t f2_pk(float a, float b) { uint32_t x, y; std::memcpy(&x, &a, 4); std::memcpy(&y, &b, 4); return ((f2_t)y << 32) | x; }
static inline void f2_unpk(f2_t v, float& a, float& b) { uint32_t x = (uint32_t)v, y = (uint32_t)(v >> 32); std::memcpy(&a, &x, 4); std::memcpy(&b, &y, 4); }
static inline f2_t f2_fma(f2_t a, f2_t b, f2_t c) { float ax, ay, bx, by, cx, cy; f2_unpk(a, ax, ay); f2_unpk(b, bx, by); f2_unpk(c, cx, cy); return f2_pk(std::fmaf(ax, bx, cx), std::fmaf(ay, by, cy)); }
static inline f2_t f2_mul(f2_t a, f2_t b) { float ax, ay, bx, by; f2_unpk(a, ax, ay); f2_unpk(b, bx, by); return f2_pk(ax * bx, ay * by); }
static inline f2_t f2_add(f2_t a, f2_t b) { float ax, ay, bx, by; f2_unpk(a, ax, ay); f2_unpk(b, bx, by); return f2_pk(ax + bx, ay + by); }
static inline f2_t f2_sub(f2_t a, f2_t b) { float ax, ay, bx, by; f2_unpk(a, ax, ay); f2_unpk(b, bx, by); return f2_pk(ax - bx, ay - by); }
struct uint2 { unsigned x, y; };
static inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r; r.x = x; r.y = y; return r; }
template <int BIT> static inline unsigned fast_psel(unsigned m, unsigned lo, unsigned hi) { return (m >> BIT) & 1u ? hi : lo; }
#define BDS_TRK_B1C_WB 0
#define BDS_TRK_B1C_NB 1
#define FAST_CONST static const
"""

DRIVER = r"""
using namespace bds;
// the table exactly as the builder warp of trk_fw_kernel makes it: 32 lanes, then the vote
static void build_tab(FastTab* tab, const FastStatic& fsx, const EpochParams& p, double fs) {
    int ok = 1;
    for (int lane = 0; lane < 32; ++lane) ok &= fast_build_tab_lane(tab, fsx, p, fs, lane);
    tab->valid = ok;
}
extern "C" int run_epoch(const EpochParams* p, const uint32_t* bits, const unsigned char* tile,
                         long long tileBase, int tileBytes, long long B0, const int8_t* xblk, unsigned guard, double* sums18,
                         int* valid) {
    static FastStatic fsx;
    static FastTab tab;
    fast_load_static(&fsx, 0, 1);
    build_tab(&tab, fsx, *p, FAST_FS_HZ);
    *valid = tab.valid;
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
        nExact += fast_chip(tab, fsx, *p, bits, bits + kPackedWordsDev, tile, tileBase, tileBytes, B0, xblk, FAST_D, FAST_FS_HZ, c,
                            guard, acc);
        for (int i = 0; i < kNSum; ++i) q8[i] += __float2int_rn(fast_acc_get(acc, i) * 256.f);
    }
    for (int i = 0; i < kNSum; ++i) sums18[i] = (double)q8[i] / 256.0;
    return nExact;
}
// rank of a sub-sample phase exactly as fast_chip derives it, next to the definition (number of thresholds < Psi)
extern "C" int rank_check(const EpochParams* p, unsigned Psi, int* by_definition) {
    static FastStatic fsx;
    static FastTab tab;
    fast_load_static(&fsx, 0, 1);
    build_tab(&tab, fsx, *p, FAST_FS_HZ);
    const int jlo = fsx.rankLo[Psi >> (32 - FAST_RANK_BITS)];
    const int j = jlo + (tab.thr[jlo] < Psi);
    int n = 0;
    for (int k = 1; k <= FAST_NSEG; ++k) {
        double th = kFastBeta[k] * tab.S - (double)kFastR[k];
        n += (unsigned)std::fmin(th * 4294967296.0, 4294967295.0) < Psi;
    }
    *by_definition = n;
    return tab.valid ? j : -1;
}
"""


def _build(tmp, flags=(), geom="g99"):
    """bds_track_fast.cuh as it is (its CUDA-only helpers sit behind __CUDACC__), behind a stub of bds_track.cuh"""
    trk = open(os.path.join(CSRC, "bds_track.cuh")).read()
    stub = tmp / "stub"
    stub.mkdir(exist_ok=True)
    (stub / "bds_track.cuh").write_text("\n".join([
        "#pragma once", SHIM.replace("#define BDS_TRK_B2A 2", ""), "#include <cstddef>", F2_SHIM, "namespace bds {",
        "constexpr int kNSum = 18;\nconstexpr int kPackedWordsDev = 320;",
        _block(trk, r"__host__ __device__ constexpr int sum_idx"), "enum { EPL_E = 0, EPL_P = 1, EPL_L = 2 };",
        _block(trk, r"struct EpochParams \{"), "}"]))
    inc = GEOMS[geom][1]
    for f in ("bds_track_fast.cuh", inc):
        (stub / f).write_text(open(os.path.join(CSRC, f)).read())
    flags = tuple(flags) + ("-DFAST_GEOM_NS=" + geom, '-DFAST_GEN_INC="%s"' % inc)
    src = tmp / "b1c_host.cpp"
    src.write_text('#include "bds_track_fast.cuh"\n' + DRIVER)
    so = tmp / "b1c_host.so"
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-I", str(stub), *flags, "-o", str(so),
                        str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    lib = C.CDLL(str(so))
    lib.run_epoch.restype = C.c_int
    lib.rank_check.restype = C.c_int
    return lib


def make_epoch(rng, prn, rem, codeFreq, carrFreq, remCarr, FS=99.375e6, mode="WB"):
    """one block of the SURVEY 8(d) B1C signal model + noise, int8 (the signal always carries all three components;
    settings and replica codes are those of the tracking mode)"""
    step = codeFreq / FS
    blk = int(math.ceil((L - rem) / step))
    s = O.initSettings_B1C(samplingFreq=FS, pilotTRKflag=2)
    codes = O.make_track_codes("WB", s, prn)
    if mode == "NB":
        s_nb = O.initSettings_B1C(samplingFreq=FS, pilotTRKflag=1)
        codes_nb = O.make_track_codes("NB", s_nb, prn)
    k = np.arange(blk)
    t = rem + k * step
    i2 = np.ceil(t * 2).astype(np.int64)
    i12 = np.ceil(t * 12).astype(np.int64)
    th = carrFreq * 2.0 * np.pi * (k / FS) + remCarr
    sig = math.sqrt(11 / 44) * codes["data"][i2] * np.cos(th) - (math.sqrt(29 / 44) * codes["pilot"][i2] * np.sin(th)
                                                                    + math.sqrt(4 / 44) * codes["pilot61"][i12] * np.cos(th))
    x = np.clip(np.rint(24.0 * rng.standard_normal(blk) + 30.0 * sig), -127, 127).astype(np.int8)
    if mode == "NB":
        return s_nb, codes_nb, x, step
    return s, codes, x, step


NAMES = [f"{fam}_{iq}_{epl}" for fam in ("d", "p", "p61") for epl in ("E", "P", "L") for iq in ("I", "Q")]


@pytest.fixture(scope="module", params=sorted(GEOMS))
def hostlib(request, tmp_path_factory):
    lib = _build(tmp_path_factory.mktemp("b1c_host_" + request.param), geom=request.param)
    lib.FS = GEOMS[request.param][0]
    lib.MODE = "NB" if request.param.endswith("n") else "WB"
    text = open(os.path.join(CSRC, GEOMS[request.param][1])).read()
    lib.R = [int(v) for v in re.search(r"kFastR\[\d+\] = \{(.*?)\}", text).group(1).split(",")]
    lib.BETA = [float(v) for v in re.search(r"kFastBeta\[\d+\] = \{(.*?)\}", text).group(1).split(",")]
    return lib


def test_fast_chip_source_on_host_equals_oracle(hostlib):
    lib = hostlib
    FS = lib.FS
    rng = np.random.default_rng(33)
    prn = 19
    for rem, cf, fc_, rc, B0, guard in ((0.0, 1.023e6 - 2.7, 14.58e6 + 1830.0, 0.0, 7, 16),
                                        (0.0061, 1.023e6 + 1.9, 14.58e6 - 3920.0, 4.2, 993750 * 2 + 13, 16),
                                        (0.0033, 1.023e6 - 0.4, 14.58e6 + 55.0, 2.2, 48, 1 << 24)):
        s, codes, x, step = make_epoch(rng, prn, rem, cf, fc_, rc, FS, lib.MODE)
        p = EpochParams(pos=B0, blksize=x.size, pad=0, rem=rem, step=step, carrFreq=fc_, remCarr=rc)
        bits = np.concatenate([_pack_bits(O.b1c_data_primary(prn)), _pack_bits(O.b1c_pilot_primary(prn))])
        tileBase = B0 & ~15
        tile = np.zeros((B0 - tileBase) + x.size + 256, dtype=np.int8)
        tile[B0 - tileBase:B0 - tileBase + x.size] = x
        sums = np.zeros(18)
        valid = C.c_int(0)
        n_exact = lib.run_epoch(C.byref(p), bits.ctypes.data_as(C.c_void_p), tile.ctypes.data_as(C.c_void_p),
                                C.c_longlong(tileBase), C.c_int(tile.size), C.c_longlong(B0), x.ctypes.data_as(C.c_void_p),
                                C.c_uint(guard), sums.ctypes.data_as(C.c_void_p), C.byref(valid))
        assert valid.value == 1
        if guard == 16:
            assert n_exact <= 4, n_exact
        else:   # guard 2^24 / 2^32 of a sample on either side of every threshold and of the chip edges
            nthr = len(lib.R) - 1
            assert 0.5 * 2 * (nthr + 2) / 256 < n_exact / L < 2.5 * 2 * (nthr + 2) / 256, n_exact
        ref, _, _ = O.correlate_epoch(lib.MODE, s, x.astype(np.float64), codes, rem, step, fc_, rc)
        got = dict(zip(NAMES, sums))
        # (narrow band: the BOC(6,1) sums only receive what the exact per-sample path adds; the kernel's epilogue zeroes them)
        for fam in (("d", "p") if lib.MODE == "NB" else ("d", "p", "p61")):
            scale = max(abs(ref[f"{fam}_I_P"]), abs(ref[f"{fam}_Q_P"]))
            for nm in "EPL":
                for iq in "IQ":
                    k = f"{fam}_{iq}_{nm}"
                    assert abs(got[k] - ref[k]) <= 1e-4 * scale, (k, got[k], ref[k], scale)


def test_code_rate_far_from_nominal_invalidates_the_table_and_stays_exact(hostlib):
    """thresholds further than half a rank bin from nominal (here: a code rate 60 ppm off) clear tab.valid; every chip then
    goes through the exact per-sample path and the sums are still the oracle's"""
    lib = hostlib
    rng = np.random.default_rng(5)
    prn = 7
    rem, cf, fc_, rc, B0 = 0.0042, 1.023e6 * (1 + 60e-6), 14.58e6 + 300.0, 1.0, 32
    s, codes, x, step = make_epoch(rng, prn, rem, cf, fc_, rc, lib.FS, lib.MODE)
    p = EpochParams(pos=B0, blksize=x.size, pad=0, rem=rem, step=step, carrFreq=fc_, remCarr=rc)
    bits = np.concatenate([_pack_bits(O.b1c_data_primary(prn)), _pack_bits(O.b1c_pilot_primary(prn))])
    tile = np.zeros(B0 + x.size + 256, dtype=np.int8)
    tile[B0:B0 + x.size] = x
    sums = np.zeros(18)
    valid = C.c_int(1)
    n_exact = lib.run_epoch(C.byref(p), bits.ctypes.data_as(C.c_void_p), tile.ctypes.data_as(C.c_void_p), C.c_longlong(0),
                            C.c_int(tile.size), C.c_longlong(B0), x.ctypes.data_as(C.c_void_p), C.c_uint(16),
                            sums.ctypes.data_as(C.c_void_p), C.byref(valid))
    assert valid.value == 0 and n_exact == L
    ref, _, _ = O.correlate_epoch(lib.MODE, s, x.astype(np.float64), codes, rem, step, fc_, rc)
    for k, v in zip(NAMES, sums):
        fam = k.split("_")[0]
        if k not in ref:
            continue   # narrow band: no BOC(6,1) family
        scale = max(abs(ref[f"{fam}_I_P"]), abs(ref[f"{fam}_Q_P"]))
        assert abs(v - ref[k]) <= 1e-4 * scale, (k, v, ref[k])


def test_one_compare_rank_search_equals_the_definition(hostlib):
    """rank = number of thresholds below the sub-sample phase: table lookup + one compare (fast_chip) against counting
    all 36, for random phases, phases hugging every threshold from both sides, and code rates across +-12 kHz of Doppler"""
    lib = hostlib
    FS, BETA, R = lib.FS, lib.BETA, lib.R
    rng = np.random.default_rng(12)
    for dopp in (-12000.0, -4500.0, 0.0, 37.0, 4500.0, 12000.0):
        cf = 1.023e6 * (1 - dopp / 1575.42e6)
        step = cf / FS
        p = EpochParams(pos=0, blksize=993750, pad=0, rem=float(rng.uniform(0, step)), step=step, carrFreq=14.58e6 + dopp, remCarr=0.3)
        S = 1.0 / (12.0 * step)
        thr = [int(min((BETA[k] * S - R[k]) * 4294967296.0, 4294967295.0)) for k in range(1, len(R))]
        probes = list(rng.integers(0, 1 << 32, size=3000)) + [t + d for t in thr for d in (-2, -1, 0, 1, 2)] + [0, 1, (1 << 32) - 1]
        for Psi in probes:
            Psi = int(Psi) & 0xFFFFFFFF
            want = C.c_int(0)
            got = lib.rank_check(C.byref(p), C.c_uint(Psi), C.byref(want))
            assert got == want.value, (dopp, Psi, got, want.value)

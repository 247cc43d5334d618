"""The device code of the chip-synchronous B2a correlator, compiled for the host and run against the oracle.

tests/test_fast_b2a_model.py checks a Python restatement of the algorithm; this test runs the C++ text itself.  The
single-thread device functions of csrc/bds_track_b2a.cuh (fastb_unit with the generated body and combination,
fastb_exact_range, make_exact_ctx_b2a, fastb_code12) and the few definitions they use from bds_track.cuh /
bds_track_fast.cuh are cut out of the sources, given host definitions of the CUDA intrinsics they call (__dp2a_lo/hi,
__funnelshift_r, sel_bit_u, sincospif, ...), compiled with g++ and driven unit by unit over whole epochs.  The per-epoch
table comes from the Python model (the warp-cooperative builder cannot run on the host); the rotated code-bit array is
built as b2a_load_bits does.  The twelve sums must match the float64 oracle within the parity tolerance.  What this cannot
cover is what only exists on the GPU: the barriers, the TMA staging and the block reduction of the kernel around it."""
import ctypes as C
import math
import os
import re
import subprocess

import numpy as np
import pytest

import bds_oracle as O
from test_fast_b2a_model import Settings, build_tab, make_epoch, UNITS

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bds-3-b1c-b2a-sdr-receiver_b200", "csrc")

SHIM = r"""
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
using std::min; using std::max;
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __constant__
#define __align__(n) alignas(n)
struct int4 { int x, y, z, w; };
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline unsigned long long __double2ull_rn(double x) { return (unsigned long long)std::nearbyint(x); }
static inline int __float2int_rn(float x) { return (int)std::lrintf(x); }
static inline void sincospif(float x, float* s, float* c) { *s = (float)std::sin(M_PI * (double)x); *c = (float)std::cos(M_PI * (double)x); }
#define __sinf(x) std::sin((float)(x))   /* glibc declares functions of these names */
#define __cosf(x) std::cos((float)(x))
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
    return (unsigned)((((unsigned long long)hi << 32) | lo) >> (sh & 31));
}
static inline int __dp2a_lo(int a, int b, int c) {
    return c + (int)(int16_t)(a & 0xffff) * (int)(int8_t)(b & 0xff) + (int)(int16_t)((unsigned)a >> 16) * (int)(int8_t)((b >> 8) & 0xff);
}
static inline int __dp2a_hi(int a, int b, int c) {
    return c + (int)(int16_t)(a & 0xffff) * (int)(int8_t)((b >> 16) & 0xff) + (int)(int16_t)((unsigned)a >> 16) * (int)(int8_t)((b >> 24) & 0xff);
}
template <int BIT> static inline unsigned sel_bit_u(unsigned v, unsigned m) { return (m >> BIT) & 1u ? v : 0u; }
#define BDS_TRK_B2A 2
"""

DRIVER = r"""
extern "C" int run_epoch(const FastbTab* tab, const EpochParams* p, const uint32_t* bits, const uint32_t* ext,
                         const unsigned char* tile, long long tileBase, int tileBytes, long long B0, const int8_t* xblk,
                         unsigned guard, double* sums18) {
    long long q8[kNSum];
    for (int i = 0; i < kNSum; ++i) q8[i] = 0;
    int nExact = 0;
    if (p->rem == 0.0) {   // b2a_correlate: the t = 0 sample
        float a0[kNSum];
        for (int i = 0; i < kNSum; ++i) a0[i] = 0.f;
        ExactCtx ex;
        make_exact_ctx_b2a(*p, 0.5, FASTB_FS_HZ, ex);
        fastb_exact_range(ex, xblk, bits, bits + kPackedWordsDev, 0, 0, -100, 0, a0);
        for (int i = 0; i < kNSum; ++i) q8[i] += __float2int_rn(a0[i] * 256.f);
    }
    for (int u = 0; u < kB2aUnits; ++u) {
        float acc[kNSum];
        for (int i = 0; i < kNSum; ++i) acc[i] = 0.f;
        nExact += fastb_unit(*tab, *p, bits, bits + kPackedWordsDev, ext, ext + kPackedWordsDev, tile, tileBase, tileBytes, B0,
                             xblk, 0.5, FASTB_FS_HZ, u, guard, acc);
        for (int i = 0; i < kNSum; ++i) q8[i] += __float2int_rn(acc[i] * 256.f);
    }
    for (int i = 0; i < kNSum; ++i) sums18[i] = (double)q8[i] / 256.0;
    return nExact;
}
extern "C" int sizeof_tab() { return (int)sizeof(FastbTab); }
"""


def _block(text, start_pat):
    """source text from the match of start_pat up to the brace that closes the first '{' after it (+ a trailing ';')"""
    m = re.search(start_pat, text)
    assert m, start_pat
    i = text.index("{", m.end() - 1 if text[m.end() - 1] == "{" else m.end())
    depth, j = 0, i
    while True:
        depth += text[j] == "{"
        depth -= text[j] == "}"
        j += 1
        if depth == 0:
            break
    if text[j:j + 1] == ";":
        j += 1
    return text[m.start():j] + "\n"


@pytest.fixture(scope="module")
def hostlib(tmp_path_factory):
    trk = open(os.path.join(CSRC, "bds_track.cuh")).read()
    fast = open(os.path.join(CSRC, "bds_track_fast.cuh")).read()
    b2a = open(os.path.join(CSRC, "bds_track_b2a.cuh")).read()
    inc = open(os.path.join(CSRC, "bds_track_fast_b2a_gen.inc")).read().replace("static __constant__", "static const")
    parts = [SHIM,
             "constexpr int kNSum = 18;\nconstexpr int kPackedWordsDev = 320;\nconstexpr int kFastBins = 128;\n",
             _block(trk, r"__host__ __device__ constexpr int sum_idx"),
             "enum { EPL_E = 0, EPL_P = 1, EPL_L = 2 };\n",
             _block(trk, r"struct EpochParams \{"),
             _block(fast, r"struct ExactCtx \{"),
             _block(fast, r"__device__ __forceinline__ double colon_elem_f"),
             _block(fast, r"__device__ __forceinline__ int bit_of"),
             inc,
             "constexpr int kB2aUnits = 10230 / FASTB_CHIPS;\n",
             _block(b2a, r"struct __align__\(16\) FastbTab \{"),
             _block(b2a, r"__device__ inline void make_exact_ctx_b2a"),
             _block(b2a, r"__device__ __noinline__ void fastb_exact_range"),
             _block(b2a, r"__device__ __forceinline__ unsigned fastb_code12"),
             _block(b2a, r"__device__ __forceinline__ bool fastb_unit"),
             DRIVER]
    d = tmp_path_factory.mktemp("b2a_host")
    src = d / "b2a_host.cpp"
    src.write_text("\n".join(parts))
    so = d / "b2a_host.so"
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    lib = C.CDLL(str(so))
    lib.run_epoch.restype = C.c_int
    return lib


class FastbTab(C.Structure):
    _fields_ = [("w", C.c_int32 * (26 * 4)), ("thr", C.c_uint32 * 24), ("mask", C.c_uint32 * 24),
                ("binStart", C.c_uint8 * 144), ("u0", C.c_double), ("sigma", C.c_double), ("S", C.c_double),
                ("dphi", C.c_uint64), ("phi0", C.c_uint64), ("valid", C.c_int32), ("pad", C.c_int32 * 3),
                ("tail", C.c_uint8 * 8)]


class EpochParams(C.Structure):
    _fields_ = [("pos", C.c_longlong), ("blksize", C.c_int), ("pad", C.c_int), ("rem", C.c_double), ("step", C.c_double),
                ("carrFreq", C.c_double), ("remCarr", C.c_double)]


def _pack_bits(code_pm):
    """bit k of word w = chip 32 w + k, set <=> chip is -1 (bds_codes.h); padded to 320 words"""
    neg = np.zeros(320 * 32, dtype=np.uint8)
    neg[:code_pm.size] = code_pm < 0
    return np.packbits(neg.reshape(320, 32), axis=1, bitorder="little").view("<u4").reshape(320).copy()


def _rotate(words):
    """b2a_load_bits"""
    ext = np.zeros(320, dtype=np.uint32)
    for k in range(320):
        carry = (int(words[k - 1]) >> 31) if k else (int(words[10229 >> 5]) >> (10229 & 31)) & 1
        e = ((int(words[k]) << 1) | carry) & 0xFFFFFFFF
        if k == (10231 >> 5):
            e |= (int(words[0]) & 1) << (10231 & 31)
        ext[k] = e
    return ext


def _run(lib, x, B0, code_d, code_p, rem, step, carrFreq, remCarr, guard=16):
    assert lib.sizeof_tab() == C.sizeof(FastbTab), (lib.sizeof_tab(), C.sizeof(FastbTab))
    t = build_tab(rem, step, carrFreq, remCarr)
    tab = FastbTab()
    for i in range(26):
        for h in range(2):                                   # {wr01, wr23, wi01, wi23}
            lo, hi = t["w"][4 * i + 2 * h], t["w"][4 * i + 2 * h + 1]
            tab.w[4 * i + h] = C.c_int32(((hi[0] & 0xFFFF) << 16) | (lo[0] & 0xFFFF)).value
            tab.w[4 * i + 2 + h] = C.c_int32(((hi[1] & 0xFFFF) << 16) | (lo[1] & 0xFFFF)).value
    for i in range(24):
        tab.thr[i] = t["thr"][i]
    for i in range(21):
        tab.mask[i] = t["mask"][i]
    for i in range(129):
        tab.binStart[i] = t["bins"][i]
    tab.u0, tab.sigma, tab.S = t["u0"], 2.0 * step, t["S"]
    tab.dphi, tab.phi0, tab.valid = t["dphi"], t["phi0"], int(t["valid"])
    p = EpochParams(pos=B0, blksize=x.size, pad=0, rem=rem, step=step, carrFreq=carrFreq, remCarr=remCarr)
    bits = np.concatenate([_pack_bits(code_d), _pack_bits(code_p)])
    ext = np.concatenate([_rotate(bits[:320]), _rotate(bits[320:])])
    tileBase = B0 & ~15
    tile = np.zeros(99584, dtype=np.int8)
    n_in = min(tile.size - (B0 - tileBase), x.size)
    tile[B0 - tileBase:B0 - tileBase + n_in] = x[:n_in]
    xx = np.ascontiguousarray(x)
    sums = np.zeros(18)
    n_exact = lib.run_epoch(C.byref(tab), C.byref(p), bits.ctypes.data_as(C.c_void_p), ext.ctypes.data_as(C.c_void_p),
                            tile.ctypes.data_as(C.c_void_p), C.c_longlong(tileBase), C.c_int(tile.size), C.c_longlong(B0),
                            xx.ctypes.data_as(C.c_void_p), C.c_uint(guard), sums.ctypes.data_as(C.c_void_p))
    return sums, n_exact


NAMES = [f"{fam}_{iq}_{epl}" for fam in ("d", "p", "p61") for epl in ("E", "P", "L") for iq in ("I", "Q")]


@pytest.mark.parametrize("case", [
    dict(rem=0.0, codeFreq=10.23e6, carrFreq=13.55e6 + 1234.5, remCarr=0.0, B0=7, guard=16),
    dict(rem=0.0731, codeFreq=10.23e6 - 17.3, carrFreq=13.55e6 - 4321.0, remCarr=2.5, B0=123456, guard=16),
    dict(rem=0.0049, codeFreq=10.23e6 + 31.9, carrFreq=13.55e6 + 77.7, remCarr=5.9, B0=99375 * 3 + 11, guard=16),
    dict(rem=0.0387, codeFreq=10.23e6 + 5.0, carrFreq=13.55e6 + 2500.0, remCarr=1.0, B0=40, guard=1 << 24),
])
def test_device_source_on_host_equals_oracle(hostlib, case):
    rng = np.random.default_rng(21)
    x, step, code_d, code_p = make_epoch(rng, case["rem"], case["codeFreq"], case["carrFreq"], case["remCarr"])
    sums, n_exact = _run(hostlib, x, case["B0"], code_d, code_p, case["rem"], step, case["carrFreq"], case["remCarr"],
                         case["guard"])
    if case["guard"] == 16:
        assert n_exact <= (25 if case["rem"] == 0.0 else 3), n_exact
    else:
        assert 0.05 < n_exact / UNITS < 0.6
    codes = {"data": np.concatenate([code_d[-1:], code_d, code_d[:1]]), "pilot": np.concatenate([code_p[-1:], code_p, code_p[:1]])}
    ref, _, _ = O.correlate_epoch("B2a", Settings, x.astype(np.float64), codes, case["rem"], step, case["carrFreq"], case["remCarr"])
    got = dict(zip(NAMES, sums))
    for fam in "dp":
        scale = max(abs(ref[f"{fam}_I_P"]), abs(ref[f"{fam}_Q_P"]))
        for nm in "EPL":
            for iq in "IQ":
                k = f"{fam}_{iq}_{nm}"
                assert abs(got[k] - ref[k]) <= 1e-4 * scale, (k, got[k], ref[k], scale)
    assert all(got[k] == 0.0 for k in NAMES[12:])

"""The building blocks of the compile-time specialised inverse FFT passes (csrc/bds_acq.cu: brevq, inv_tw_offset,
ifft_regs_const, ifft_step_ct, the host's inv_step_twiddles) compiled for the host and run against numpy.

The device source is cut out of bds_acq.cu as it is; threadIdx is a plain global the driver sets before each call - a fused
step touches only the elements of its own thread between two barriers, so the threads of a CTA can be run one after the other,
one step at a time.  Checked for every transform length and both shared-memory layouts the kernels are instantiated with
(rows: element (i, b) at buf[b * seqStride + pidx(i)]; columns: buf[pidx(i) * 8 + b]): the result of all steps equals
the unnormalised inverse DFT of the bit-reversed input, i.e. index arithmetic of the padded layout, step twiddle tables
([m'][j] layout, offsets) and butterflies are right before the kernels get GPU time."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from test_fast_b2a_hostcompile import _block

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bds-3-b1c-b2a-sdr-receiver_b200", "csrc")

SHIM = r"""
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
struct Tid { unsigned x; };
static Tid threadIdx;
namespace bds {
constexpr int kColTile = 8;
constexpr int kLgColTile = 3;
"""

DRIVER = r"""
// all steps of a 2^LG-point inverse over a CTA of NT threads, one step at a time (the __syncthreads of ifft_ct)
template <int LG, int LGB, bool kCols, int SEQ, int NT, int D = 0>
static void run_all(float2* buf, const float2* tws) {
    if constexpr (D < LG) {
        constexpr int Q = (LG - D) < 4 ? (LG - D) : 4;
        for (unsigned t = 0; t < (unsigned)NT; ++t) {
            threadIdx.x = t;
            ifft_step_ct<LG, LGB, kCols, SEQ, NT, D, Q>(buf, tws + inv_tw_offset(LG, D));
        }
        run_all<LG, LGB, kCols, SEQ, NT, D + Q>(buf, tws);
    }
}
}  // namespace bds
using namespace bds;
extern "C" int tw_size(int LG) { return inv_tw_size(LG); }
extern "C" int tw_offset(int LG, int D) { return inv_tw_offset(LG, D); }
extern "C" int tw_table(int LG, float* out, int cap) {
    std::vector<float2> t = inv_step_twiddles(LG);
    if ((int)t.size() > cap) return -1;
    std::memcpy(out, t.data(), t.size() * 8);
    return (int)t.size();
}
extern "C" int pidx_host(int i) { return pidx(i); }
// shape: 0 = rows 2^12 x 2 (B1C), 1 = rows 2^12 x 1, 2 = rows 2^10 x 4 (B2a), 3 = rows 2^10 x 2,
//        4 = columns 2^10 x 8 (B1C), 5 = columns 2^9 x 8 (B2a, B1C at 53 MHz)
extern "C" int run_shape(int shape, float* buf, const float* tw) {
    float2* b = reinterpret_cast<float2*>(buf);
    const float2* t = reinterpret_cast<const float2*>(tw);
    switch (shape) {
        case 0: run_all<12, 1, false, 4096 + 256, 512>(b, t); return 0;
        case 1: run_all<12, 0, false, 4096 + 256, 256>(b, t); return 0;
        case 2: run_all<10, 2, false, 1024 + 64, 256>(b, t); return 0;
        case 3: run_all<10, 1, false, 1024 + 64, 128>(b, t); return 0;
        case 4: run_all<10, 3, true, 0, 512>(b, t); return 0;
        case 5: run_all<9, 3, true, 0, 256>(b, t); return 0;
    }
    return -1;
}
"""

SHAPES = {0: (12, 1, False), 1: (12, 0, False), 2: (10, 2, False), 3: (10, 1, False), 4: (10, 3, True), 5: (9, 3, True)}


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    src = open(os.path.join(CSRC, "bds_acq.cu")).read()
    parts = [SHIM]
    for pat in (r"__device__ __forceinline__ float2 cmul\(", r"__device__ __forceinline__ float2 cmulc\(",
                r"__device__ __forceinline__ int pidx\(", r"__device__ __forceinline__ float2 mul_w16\(",
                r"__device__ __forceinline__ float2 mul_w16c\(", r"template <int Q>\n__host__ __device__ constexpr int brevq",
                r"__host__ __device__ constexpr int inv_tw_offset", r"__host__ __device__ constexpr int inv_tw_size",
                r"template <int Q>\n__device__ __forceinline__ void ifft_regs_const",
                r"template <int LG, int LGB, bool kCols, int SEQ, int NT, int D, int Q>\n__device__ __forceinline__ void ifft_step_ct",
                r"std::vector<float2> inv_step_twiddles"):
        parts.append(_block(src, pat))
    tmp = tmp_path_factory.mktemp("acq_ct")
    cpp = tmp / "acq_ct.cpp"
    cpp.write_text("\n".join(parts) + DRIVER)
    so = tmp / "acq_ct.so"
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so), str(cpp)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return C.CDLL(str(so))


def _brev(x, bits):
    r = 0
    for k in range(bits):
        r |= ((x >> k) & 1) << (bits - 1 - k)
    return r


def _table(lib, LG):
    n = lib.tw_size(LG)
    buf = np.zeros(2 * max(n, 1), dtype=np.float32)
    got = lib.tw_table(LG, buf.ctypes.data_as(C.c_void_p), max(n, 1))
    assert got == max(n, 1)
    return buf


@pytest.mark.parametrize("LG", [8, 9, 10, 12])
def test_step_twiddle_table_layout(lib, LG):
    """T[(m' - 1) 2^D + j] = exp(+2 pi i j m' / 2^(D+Q)) at offset inv_tw_offset(LG, D), steps D = 4, 8 with Q = min(4, LG - D)"""
    t = _table(lib, LG).view(np.complex64)
    off = 0
    for D in range(4, LG, 4):
        Q = min(4, LG - D)
        assert lib.tw_offset(LG, D) == off
        h = 1 << D
        for mp in (1, (1 << Q) - 1):
            j = np.arange(h)
            want = np.exp(2j * np.pi * j * mp / (1 << (D + Q)))
            np.testing.assert_allclose(t[off + (mp - 1) * h: off + mp * h], want, atol=1e-7)
        off += ((1 << Q) - 1) * h
    assert lib.tw_size(LG) == off


@pytest.mark.parametrize("shape", sorted(SHAPES))
def test_all_steps_equal_the_inverse_dft_of_the_bit_reversed_input(lib, shape):
    LG, LGB, cols = SHAPES[shape]
    n, nb = 1 << LG, 1 << LGB
    rng = np.random.default_rng(100 + shape)
    X = (rng.standard_normal((nb, n)) + 1j * rng.standard_normal((nb, n))).astype(np.complex64)
    rev = np.array([_brev(k, LG) for k in range(n)])
    pid = np.arange(n) + (np.arange(n) >> 4)
    assert all(lib.pidx_host(int(i)) == int(pid[i]) for i in (0, 15, 16, 17, n - 1))
    stride = n + (n >> 4)
    buf = np.zeros(stride * nb, dtype=np.complex64)
    for b in range(nb):
        if cols:
            buf[pid * nb + b] = X[b, rev]
        else:
            buf[b * stride + pid] = X[b, rev]
    tw = _table(lib, LG)
    raw = buf.view(np.float32)
    assert lib.run_shape(shape, raw.ctypes.data_as(C.c_void_p), tw.ctypes.data_as(C.c_void_p)) == 0
    for b in range(nb):
        got = buf[pid * nb + b] if cols else buf[b * stride + pid]
        want = np.fft.ifft(X[b].astype(np.complex128)) * n
        assert np.max(np.abs(got - want)) <= 2e-6 * np.max(np.abs(want)) * LG, (shape, b)

"""The warp-cooperative per-epoch table builders (fast_build_tab_warp, fastb_build_tab_warp) run on the host.

A warp is emulated by 32 OS threads: threadIdx.x is thread-local, __syncwarp() is a 32-party barrier and __all_sync() a
barrier-protected vote, so the code keeps exactly the synchronisation it has on the GPU (a missing __syncwarp shows up
as a data race here too).  The tables built by the device source are compared with the Python restatement the other
host-compiled tests use, and a whole epoch is correlated through a table built this way:
  * B2a: thresholds, rank masks, bins, scalars; epoch sums against the oracle;
  * B1C: default and -DBDS_FAST_BINREC=1 (per-bin records), fresh build and the reuse path (same sort order as the
    previous epoch's table: only thresholds, records, rotation table and scalars are rewritten)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import bds_oracle as O
import test_fast_b1c_hostcompile as B1
import test_fast_b2a_hostcompile as B2
from test_fast_b2a_model import Settings, build_tab, make_epoch

CSRC = B2.CSRC

WARP_SHIM = r"""
#include <thread>
#include <barrier>
#include <vector>
struct TidX { unsigned x; };
static thread_local TidX threadIdx;
static std::barrier<> g_warp_bar(32);
static int g_warp_pred[32];
static inline void __syncwarp() { g_warp_bar.arrive_and_wait(); }
static inline int __all_sync(unsigned, int p) {
    g_warp_pred[threadIdx.x & 31] = p;
    g_warp_bar.arrive_and_wait();
    int r = 1;
    for (int i = 0; i < 32; ++i) r &= g_warp_pred[i] != 0;
    g_warp_bar.arrive_and_wait();
    return r;
}
template <typename F> static void run_warp(F f) {
    std::vector<std::thread> th;
    for (unsigned l = 0; l < 32; ++l) th.emplace_back([=] { threadIdx.x = l; f(); });
    for (auto& t : th) t.join();
}
"""

B2A_DRIVER = r"""
extern "C" void build_tab_host(FastbTab* tab, const EpochParams* p, double fs) {
    static unsigned scratch[64];
    run_warp([=] { fastb_build_tab_warp(tab, *p, fs, scratch); });
}
"""

B1C_DRIVER = r"""
extern "C" void build_tab_host(FastTab* tab, const EpochParams* p, double fs, const unsigned char* prev) {
    static unsigned scratch[128];
    run_warp([=] { fast_build_tab_warp(tab, *p, fs, scratch, prev); });
}
extern "C" int offsetof_posbin() { return (int)offsetof(FastTab, posbin); }
"""


def _compile(tmp, name, text, flags=()):
    src = tmp / (name + ".cpp")
    src.write_text(text)
    so = tmp / (name + ".so")
    r = subprocess.run(["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-pthread", "-shared", "-fPIC", *flags, "-o", str(so), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return C.CDLL(str(so))


@pytest.fixture(scope="module")
def b2a_lib(tmp_path_factory):
    trk = open(os.path.join(CSRC, "bds_track.cuh")).read()
    fast = open(os.path.join(CSRC, "bds_track_fast.cuh")).read()
    b2a = open(os.path.join(CSRC, "bds_track_b2a.cuh")).read()
    inc = open(os.path.join(CSRC, "bds_track_fast_b2a_gen.inc")).read().replace("static __constant__", "static const")
    blk = B2._block
    parts = [B2.SHIM, WARP_SHIM,
             "constexpr int kNSum = 18;\nconstexpr int kPackedWordsDev = 320;\nconstexpr int kFastBins = 128;\n",
             blk(trk, r"__host__ __device__ constexpr int sum_idx"), "enum { EPL_E = 0, EPL_P = 1, EPL_L = 2 };\n",
             blk(trk, r"struct EpochParams \{"), blk(fast, r"struct ExactCtx \{"),
             blk(fast, r"__device__ __forceinline__ double colon_elem_f"), blk(fast, r"__device__ __forceinline__ int bit_of"),
             inc, "constexpr int kB2aUnits = 10230 / FASTB_CHIPS;\n",
             blk(b2a, r"struct __align__\(16\) FastbTab \{"), blk(b2a, r"__device__ void fastb_build_tab_warp"),
             blk(b2a, r"__device__ inline void make_exact_ctx_b2a"), blk(b2a, r"__device__ __noinline__ void fastb_exact_range"),
             blk(b2a, r"__device__ __forceinline__ unsigned fastb_code12"), blk(b2a, r"__device__ __forceinline__ bool fastb_unit"),
             B2.DRIVER, B2A_DRIVER]
    lib = _compile(tmp_path_factory.mktemp("b2a_tab"), "b2a_tab", "\n".join(parts))
    lib.run_epoch.restype = C.c_int
    return lib


def test_b2a_table_builder_on_host(b2a_lib):
    rng = np.random.default_rng(3)
    for rem, cf, fc_, rc in ((0.0, 10.23e6, 13.55e6 + 1234.5, 0.0), (0.0731, 10.23e6 - 17.3, 13.55e6 - 4321.0, 2.5),
                             (0.0049, 10.23e6 + 31.9, 13.55e6 + 77.7, 5.9)):
        x, step, code_d, code_p = make_epoch(rng, rem, cf, fc_, rc)
        tab = B2.FastbTab()
        p = B2.EpochParams(pos=7, blksize=x.size, pad=0, rem=rem, step=step, carrFreq=fc_, remCarr=rc)
        b2a_lib.build_tab_host(C.byref(tab), C.byref(p), C.c_double(99.375e6))
        t = build_tab(rem, step, fc_, rc)
        assert list(tab.thr)[:24] == t["thr"][:24]
        assert list(tab.mask)[:21] == t["mask"]
        assert list(tab.binStart)[:129] == t["bins"]
        assert (tab.u0, tab.S, tab.dphi, tab.phi0, tab.valid) == (t["u0"], t["S"], t["dphi"], t["phi0"], int(t["valid"]))
        w = np.frombuffer(bytes(tab.w), dtype=np.int16).reshape(26, 8)
        for i in range(26 * 4):
            wr, wi = int(w[i >> 2, i & 3]), int(w[i >> 2, 4 + (i & 3)])
            assert abs(wr - t["w"][i][0]) <= 1 and abs(wi - t["w"][i][1]) <= 1
        # a whole epoch through the table the device source built
        bits = np.concatenate([B2._pack_bits(code_d), B2._pack_bits(code_p)])
        ext = np.concatenate([B2._rotate(bits[:320]), B2._rotate(bits[320:])])
        B0 = 7
        tile = np.zeros(99584, dtype=np.int8)
        n_in = min(tile.size - B0, x.size)
        tile[B0:B0 + n_in] = x[:n_in]
        sums = np.zeros(18)
        b2a_lib.run_epoch(C.byref(tab), C.byref(p), bits.ctypes.data_as(C.c_void_p), ext.ctypes.data_as(C.c_void_p),
                          tile.ctypes.data_as(C.c_void_p), C.c_longlong(0), C.c_int(tile.size), C.c_longlong(B0),
                          x.ctypes.data_as(C.c_void_p), C.c_uint(16), sums.ctypes.data_as(C.c_void_p))
        codes = {"data": np.concatenate([code_d[-1:], code_d, code_d[:1]]), "pilot": np.concatenate([code_p[-1:], code_p, code_p[:1]])}
        ref, _, _ = O.correlate_epoch("B2a", Settings, x.astype(np.float64), codes, rem, step, fc_, rc)
        got = dict(zip(B2.NAMES, sums))
        for k, v in ref.items():
            scale = max(abs(ref[f"{k[0]}_I_P"]), abs(ref[f"{k[0]}_Q_P"]))
            assert abs(got[k] - v) <= 1e-4 * scale, (k, got[k], v)


@pytest.mark.parametrize("binrec", [0, 1])
def test_b1c_table_builder_on_host(binrec, tmp_path):
    trk = open(os.path.join(CSRC, "bds_track.cuh")).read()
    fast = open(os.path.join(CSRC, "bds_track_fast.cuh")).read()
    inc = open(os.path.join(CSRC, "bds_track_fast_gen.inc")).read().replace("static __constant__", "static const")
    for pat in B1.GPU_ONLY[1:]:                               # keep fast_build_tab_warp
        fast = B1._cut(fast, pat)
    fast = fast.replace("#pragma once", "").replace('#include "bds_track.cuh"', "").replace('#include "bds_track_fast_gen.inc"', inc)
    fast = fast.replace("typedef unsigned long long f2_t;", "").replace("namespace bds {", "", 1)
    fast = fast[:fast.rindex("}  // namespace bds")]
    parts = [B2.SHIM.replace("#define BDS_TRK_B2A 2", ""), "#include <cstddef>\n", B1.F2_SHIM, WARP_SHIM,
             "constexpr int kNSum = 18;\nconstexpr int kPackedWordsDev = 320;\n",
             B2._block(trk, r"__host__ __device__ constexpr int sum_idx"), "enum { EPL_E = 0, EPL_P = 1, EPL_L = 2 };\n",
             B2._block(trk, r"struct EpochParams \{"), fast, B1.DRIVER, B1C_DRIVER]
    lib = _compile(tmp_path, "b1c_tab", "\n".join(parts), ["-DBDS_FAST_BINREC=%d" % binrec])
    size, off_u0, off_pb = lib.sizeof_tab(), lib.offsetof_u0(), lib.offsetof_posbin()

    def build(rem, step, fc_, rc, prev=None):
        buf = (C.c_uint8 * size)()
        if prev is not None:                                   # the table is rewritten in place on the GPU
            C.memmove(buf, prev[0], size)
        p = B2.EpochParams(pos=0, blksize=993750, pad=0, rem=rem, step=step, carrFreq=fc_, remCarr=rc)
        pb = (C.c_uint8 * 80).from_buffer_copy(bytes(prev[0])[off_pb:off_pb + 80]) if prev is not None else None
        lib.build_tab_host(buf, C.byref(p), C.c_double(99.375e6), pb)
        return buf, bytes(buf)

    def check(raw, rem, step, fc_, rc):
        want = B1.build_tab_bytes(lib, rem, step, fc_, rc, bool(binrec))
        w_end = 26 * 16
        got_w, want_w = np.frombuffer(raw[:w_end], dtype=np.int16), np.frombuffer(want[:w_end], dtype=np.int16)
        assert np.max(np.abs(got_w.astype(int) - want_w.astype(int))) <= 1
        thr_end = w_end + 160
        assert raw[w_end:thr_end] == want[w_end:thr_end]                       # sorted thresholds
        got_m = np.frombuffer(raw[thr_end:thr_end + 320], dtype=np.uint64)[:37]
        want_m = np.frombuffer(want[thr_end:thr_end + 320], dtype=np.uint64)[:37]
        assert np.array_equal(got_m, want_m)                                   # rank masks
        assert raw[thr_end + 320:thr_end + 320 + 129] == want[thr_end + 320:thr_end + 320 + 129]   # bins
        if binrec:
            assert raw[off_pb + 80:off_u0] == want[off_pb + 80:off_u0]         # per-bin records
        assert raw[off_u0:off_u0 + 44] == want[off_u0:off_u0 + 44]             # u0, sigma, S, dphi, phi0, valid

    fs = 99.375e6
    a = ((0.0061, (1.023e6 + 1.9) / fs, 14.58e6 - 3920.0, 4.2), (0.0033, (1.023e6 + 1.9003) / fs, 14.58e6 - 3919.2, 1.7))
    first = build(*a[0])
    check(first[1], *a[0])
    again = build(*a[1], prev=first)                           # same order, same bins: the reuse path
    check(again[1], *a[1])
    other = build(0.004, (1.023e6 - 3.0) / fs, 14.58e6 + 10.0, 0.3, prev=again)   # thresholds move across bins: full rebuild
    check(other[1], 0.004, (1.023e6 - 3.0) / fs, 14.58e6 + 10.0, 0.3)

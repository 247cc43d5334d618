"""The warp-cooperative per-epoch table builder of the B2a kernel (fastb_build_tab_warp) run on the host.

A warp is emulated by 32 OS threads: threadIdx.x is thread-local, __syncwarp() is a 32-party barrier and __all_sync() a
barrier-protected vote, so the code keeps exactly the synchronisation it has on the GPU (a missing __syncwarp shows up
as a data race here too).  The tables built by the device source are compared with the Python restatement the other
host-compiled tests use, and a whole epoch is correlated through a table built this way:
thresholds, rank masks, bins, scalars; epoch sums against the oracle.  (The B1C builder has no cross-lane dependency
any more - tests/test_fast_b1c_hostcompile.py runs its 32 lanes in turn.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import bds_oracle as O
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
             blk(b2a, r"struct __align__\(16\) FastbTab \{"), blk(b2a, r"__device__ inline void fastb_build_rot"),
             blk(b2a, r"__device__ void fastb_build_tab_warp"),
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

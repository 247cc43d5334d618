// Frame-synchronisation correlation on the device: the first consumer of the tracking output.
//
//   B1C  BDS-3_B1C/include/BCNAV1decoding.m:66-91  bits = sign(Pilot_I_P | Pilot_Q_P); XcorrResult = xcorr(bits,
//        Secondary) kept for lags >= 0; index = find(abs(XcorrResult) >= 1799.5); Secondary = the PRN's 1800-chip pilot
//        secondary Weil code, generate2ndCode.m:44-84 (Legendre sequence of length 3607, per-PRN (w, p) table :44-56)
//   B2a  BDS-3_B2a/include/BCNAV2decoding.m:69-97  bits = sign(I_P); pattern = kron(preamble_bits, secondCode)
//        (24 x 5 = 120 taps); index = find(abs(xcorr) > 115)
// MATLAB's xcorr(x, y) pads y with zeros: XcorrResult(lag + 1) = sum_n x(n + lag) y(n) over the n that exist.
//
// Integer work, bit-exact: the signs are packed 32 per word, a lag is 2 * popcount(agreements) - (taps inside the
// record); one thread per lag.
#include <algorithm>
#include <cstring>
#include <vector>

#include "bds_common.cuh"

namespace bds {
namespace {

// generate2ndCode.m:44-56 (ICD constants): pilot secondary code (w, p) per PRN
const uint16_t k2ndW[63] = {269,  1448, 1028, 1324, 822,  5,    155,  458,  310,  959,  1238, 1180, 1288, 334,  885,  1362,
                            181,  1648, 838,  313,  750,  225,  1477, 309,  108,  1457, 149,  322,  271,  576,  1103, 450,
                            399,  241,  1045, 164,  513,  687,  422,  303,  324,  495,  725,  780,  367,  882,  631,  37,
                            647,  1043, 24,   120,  134,  136,  158,  214,  335,  340,  661,  889,  929,  1002, 1149};
const uint16_t k2ndP[63] = {1889, 1268, 1593, 1186, 1239, 1930, 176,  1696, 26,   1344, 1271, 1182, 1381, 1604, 1333, 1185,
                            31,   704,  1190, 1646, 1385, 113,  860,  1656, 1921, 1173, 1928, 57,   150,  1214, 1148, 1458,
                            1519, 1635, 1257, 1687, 1382, 1514, 1,    1583, 1806, 1664, 1338, 1111, 1706, 1543, 1813, 228,
                            2871, 2884, 1823, 75,   11,   63,   1937, 22,   1768, 1526, 1402, 1445, 1680, 1290, 1245};
constexpr int k2ndN = 3607, k2ndLen = 1800;

// +1 -> bit 1, -1 -> bit 0
void secondary_bits(int prn, std::vector<uint8_t>& bits) {
    std::vector<uint8_t> leg(k2ndN, 0);
    for (long long k = 1; k < k2ndN; ++k) leg[(k * k) % k2ndN] = 1;   // Legendre symbol +1 (JacobiSymbol.m for prime N)
    bits.resize(k2ndLen);
    const int w = k2ndW[prn - 1], p = k2ndP[prn - 1];
    for (int ind = 0; ind < k2ndLen; ++ind) {
        const int k = (ind + p - 1) % k2ndN;
        const int chip = leg[k] ^ leg[(k + w) % k2ndN];   // Secondary = 1 - 2*chip: chip 0 -> +1
        bits[ind] = (uint8_t)(chip == 0);
    }
}
void b2a_preamble_bits(std::vector<uint8_t>& bits) {
    static const int pre[24] = {-1, -1, -1, 1, 1, 1, -1, 1, 1, -1, 1, 1, -1, -1, 1, -1, -1, -1, -1, 1, -1, 1, 1, 1};
    static const int sec[5] = {1, 1, 1, -1, 1};
    bits.clear();
    for (int i = 0; i < 24; ++i)
        for (int j = 0; j < 5; ++j) bits.push_back((uint8_t)(pre[i] * sec[j] > 0));   // kron(preamble_bits, secondCode)
}

__global__ void nav_pack_sign_kernel(const double* v, int n, uint32_t* words, int nWords) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nWords) return;
    uint32_t b = 0;
    for (int k = 0; k < 32; ++k) {
        const int i = w * 32 + k;
        if (i < n && v[i] > 0.0) b |= 1u << k;   // bits(bits > 0) = 1; bits(bits <= 0) = -1
    }
    words[w] = b;
}

// x: n sign bits (+ two zero words of padding), y: K pattern bits; out[lag] = sum over n < min(K, N - lag) of x[n+lag]*y[n]
__global__ void nav_xcorr_kernel(const uint32_t* x, int n, const uint32_t* y, int K, int* out, double* outD, int thr2 /*2*threshold*/,
                                 int strict, int* hits, int hitCap, int* nHits) {
    const int lag = blockIdx.x * blockDim.x + threadIdx.x;
    if (lag >= n) return;
    const int valid = min(K, n - lag);
    const int sh = lag & 31, w0 = lag >> 5;
    int agree = 0;
    for (int i = 0; i * 32 < valid; ++i) {
        const uint32_t xs = __funnelshift_r(x[w0 + i], x[w0 + i + 1], sh);
        const int rem = valid - i * 32;
        const uint32_t m = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
        agree += __popc(~(xs ^ y[i]) & m);
    }
    const int c = 2 * agree - valid;
    out[lag] = c;
    if (outD) outD[lag] = (double)c;
    const int a2 = 2 * (c < 0 ? -c : c);
    if (strict ? a2 > thr2 : a2 >= thr2) {
        const int pos = atomicAdd(nHits, 1);
        if (pos < hitCap) hits[pos] = lag + 1;   // MATLAB find: 1-based
    }
}

}  // namespace
}  // namespace bds

using namespace bds;

extern "C" int bds_secondary_code(int prn, int8_t* out1800) {
    if (prn < 1 || prn > 63 || !out1800) return set_error(BDS_ERR_ARG, "bds_secondary_code: bad arguments");
    std::vector<uint8_t> b;
    secondary_bits(prn, b);
    for (int i = 0; i < k2ndLen; ++i) out1800[i] = b[i] ? 1 : -1;
    return BDS_OK;
}

extern "C" int bds_frame_sync(int signal, int prn, const double* prompt, int n, int loc, double* xcorr, int32_t* index,
                              int index_cap, int32_t* n_index) {
    if (!prompt || n <= 0 || !n_index || (index_cap > 0 && !index)) return set_error(BDS_ERR_ARG, "bds_frame_sync: bad arguments");
    if (signal != BDS_SIG_B1C && signal != BDS_SIG_B2A) return set_error(BDS_ERR_ARG, "bds_frame_sync: unknown signal %d", signal);
    if (signal == BDS_SIG_B1C && (prn < 1 || prn > 63)) return set_error(BDS_ERR_ARG, "bds_frame_sync: PRN %d out of range", prn);
    int rc = require_device();
    if (rc) return rc;
    std::vector<uint8_t> pat;
    int thr2, strict;
    if (signal == BDS_SIG_B1C) {
        secondary_bits(prn, pat);
        thr2 = 3599;   // abs(X) >= 1799.5
        strict = 0;
    } else {
        b2a_preamble_bits(pat);
        thr2 = 230;    // abs(X) > 115
        strict = 1;
    }
    const int K = (int)pat.size(), kw = (K + 31) / 32, nw = (n + 31) / 32;
    std::vector<uint32_t> yw(kw, 0);
    for (int i = 0; i < K; ++i)
        if (pat[i]) yw[i >> 5] |= 1u << (i & 31);
    double* dV = nullptr;
    uint32_t *dX = nullptr, *dY = nullptr;
    int *dOut = nullptr, *dHits = nullptr, *dN = nullptr;
    double* dOutD = nullptr;
    const int hitCap = std::max(index_cap, 1);
    auto cleanup = [&]() {
        if (loc == BDS_LOC_HOST) cudaFree(dV);
        cudaFree(dX);
        cudaFree(dY);
        cudaFree(dOut);
        cudaFree(dHits);
        cudaFree(dN);
        cudaFree(dOutD);
    };
#define TRYN(x_)                                                      \
    if ((x_) != cudaSuccess) {                                        \
        cleanup();                                                    \
        return set_error(BDS_ERR_CUDA, "bds_frame_sync: %s failed", #x_); \
    }
    if (loc == BDS_LOC_HOST) {
        TRYN(cudaMalloc(&dV, sizeof(double) * (size_t)n));
        TRYN(cudaMemcpy(dV, prompt, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
    } else {
        dV = const_cast<double*>(prompt);
    }
    TRYN(cudaMalloc(&dX, 4 * (size_t)(nw + kw + 2)));
    TRYN(cudaMemset(dX, 0, 4 * (size_t)(nw + kw + 2)));
    TRYN(cudaMalloc(&dY, 4 * (size_t)kw));
    TRYN(cudaMemcpy(dY, yw.data(), 4 * (size_t)kw, cudaMemcpyHostToDevice));
    TRYN(cudaMalloc(&dOut, 4 * (size_t)n));
    TRYN(cudaMalloc(&dHits, 4 * (size_t)hitCap));
    TRYN(cudaMalloc(&dN, 4));
    TRYN(cudaMemset(dN, 0, 4));
    if (xcorr) TRYN(cudaMalloc(&dOutD, sizeof(double) * (size_t)n));
    nav_pack_sign_kernel<<<(nw + 127) / 128, 128>>>(dV, n, dX, nw);
    nav_xcorr_kernel<<<(n + 127) / 128, 128>>>(dX, n, dY, K, dOut, dOutD, thr2, strict, dHits, hitCap, dN);
    count_launch(2);
    TRYN(cudaGetLastError());
    int nh = 0;
    TRYN(cudaMemcpy(&nh, dN, 4, cudaMemcpyDeviceToHost));
    *n_index = nh;
    const int take = std::min(nh, index_cap);
    if (take > 0) {
        TRYN(cudaMemcpy(index, dHits, 4 * (size_t)take, cudaMemcpyDeviceToHost));
        std::sort(index, index + take);   // ascending, like find()
    }
    if (xcorr) TRYN(cudaMemcpy(xcorr, dOutD, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
#undef TRYN
    cleanup();
    return BDS_OK;
}

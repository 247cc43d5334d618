// Synthetic int8 IF generator.  The reference has no signal generator;
// this one produces the signal model that the reference trackers' discriminators lock
// onto with negative feedback:
//   B1C (WB_tracking.m:341-346,375-380; NB_tracking.m:357):
//     x = A[ sqrt(11/44) D c_d11 cos(th) - S ( sqrt(29/44) c_p11 sin(th) + sqrt(4/44) c_p61 cos(th) ) ]
//     code rate = codeFreqBasis (1 - doppler/carrFreqBasis)        (include/preRun.m:71-73)
//   B2a (tracking.m:309-314,345-348):
//     x = A[ D c_d sin(th) + S c_p cos(th) ],  code rate = codeFreqBasis (preRun.m:70)
// plus AWGN, rounded to nearest and clipped to +-127.  D / S are +-1, constant over one
// primary-code period, drawn from a counter-based hash so any window of the record can be
// generated independently.
#include <cmath>
#include <vector>

#include "bds_codes.h"
#include "bds_common.cuh"

namespace bds {

struct SynthSat {
    unsigned long long dphi;   // carrier turns/sample * 2^64
    unsigned long long phi0;   // carrier phase at sample 0 * 2^64
    double codeRatePerSample;  // chips per sample
    double codeDelay;          // samples
    float amp;
    int idx;
};

__host__ __device__ inline unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// +-1 symbol of satellite s, code period p, stream k (0 = data D, 1 = pilot secondary S)
__host__ __device__ inline float sym(unsigned long long seed, int s, long long period, int k) {
    unsigned long long h = splitmix64(seed ^ (0xD1B54A32D192ED03ull * (unsigned long long)(s * 2 + k + 1)) ^
                                      (unsigned long long)period * 0x2545F4914F6CDD1Dull);
    return (h >> 40) & 1 ? -1.f : 1.f;
}

__global__ void synth_kernel(int signal, const SynthSat* sats, int nSats, const uint32_t* bits, float sigma,
                             unsigned long long seed, long long first, size_t n, int8_t* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long smp = first + (long long)i;
    float acc = 0.f;
    for (int s = 0; s < nSats; ++s) {
        const SynthSat sat = sats[s];
        unsigned long long ph = sat.phi0 + (unsigned long long)smp * sat.dphi;
        float sn, cs;
        sincospif((float)(int)(ph >> 32) * 4.656612873077392578125e-10f, &sn, &cs);
        double tc = ((double)smp - sat.codeDelay) * sat.codeRatePerSample;
        double per = floor(tc / 10230.0);
        double tin = tc - per * 10230.0;
        int chip = (int)floor(tin);
        if (chip >= 10230) chip = 10229;
        if (chip < 0) chip = 0;
        float frac = (float)(tin - (double)chip);
        const uint32_t* bd = bits + (size_t)(s * 2) * kPackedWords;
        const uint32_t* bp = bd + kPackedWords;
        float cd = (bd[chip >> 5] >> (chip & 31)) & 1 ? -1.f : 1.f;
        float cp = (bp[chip >> 5] >> (chip & 31)) & 1 ? -1.f : 1.f;
        float D = sym(seed, sat.idx, (long long)per, 0), S = sym(seed, sat.idx, (long long)per, 1);
        if (signal == BDS_SIG_B1C) {
            float sc1 = frac < 0.5f ? -1.f : 1.f;                  // BOC(1,1): chip -> [-c, +c]
            int i6 = (int)(frac * 12.f);
            if (i6 > 11) i6 = 11;
            float sc6 = (i6 & 1) ? 1.f : -1.f;                     // BOC(6,1): (-1)^ii, ii = i6+1
            acc += sat.amp * (0.5f * D * cd * sc1 * cs -
                              S * (0.81184414f * cp * sc1 * sn + 0.30151134f * cp * sc6 * cs));
        } else {
            acc += sat.amp * (D * cd * sn + S * cp * cs);
        }
    }
    // AWGN: Box-Muller on a per-sample counter hash
    unsigned long long h = splitmix64(seed * 0x9E3779B97F4A7C15ull + (unsigned long long)smp);
    float u1 = ((float)(unsigned)(h >> 40) + 0.5f) * (1.0f / 16777216.0f);
    float u2 = (float)(unsigned)((h >> 8) & 0xFFFFFF) * (1.0f / 16777216.0f);
    float g = sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
    float v = rintf(acc + sigma * g);
    v = fminf(127.f, fmaxf(-127.f, v));
    out[i] = (int8_t)v;
}

}  // namespace bds

using namespace bds;

extern "C" int bds_synth_if(int signal, double fs, double IF, double carrFreqBasis, double codeFreqBasis,
                            const bds_sat* sats, int n_sats, double noise_sigma, uint64_t seed,
                            long long first_sample, size_t n, int8_t* out, int out_loc) {
    if (!sats || n_sats <= 0 || !out) return set_error(BDS_ERR_ARG, "bds_synth_if: bad arguments");
    if (signal != BDS_SIG_B1C && signal != BDS_SIG_B2A) return set_error(BDS_ERR_ARG, "unknown signal %d", signal);
    int rc = require_device();
    if (rc) return rc;
    std::vector<SynthSat> hs(n_sats);
    std::vector<uint32_t> bits((size_t)n_sats * 2 * kPackedWords);
    std::vector<uint8_t> chips;
    const double two64 = 18446744073709551616.0;
    for (int s = 0; s < n_sats; ++s) {
        const bds_sat& in = sats[s];
        if (in.PRN < 1 || in.PRN > 63) return set_error(BDS_ERR_ARG, "satellite %d: PRN %d out of range", s, in.PRN);
        double r = (IF + in.doppler) / fs;
        r -= std::floor(r);
        double r0 = in.carrPhase / 6.283185307179586476925286766559;
        r0 -= std::floor(r0);
        hs[s].dphi = (unsigned long long)(r * two64);
        hs[s].phi0 = (unsigned long long)(r0 * two64);
        double rate = signal == BDS_SIG_B1C ? codeFreqBasis * (1.0 - in.doppler / carrFreqBasis) : codeFreqBasis;
        hs[s].codeRatePerSample = rate / fs;
        hs[s].codeDelay = in.codeDelay;
        hs[s].amp = (float)in.amplitude;
        hs[s].idx = s;
        primary_bits(signal == BDS_SIG_B1C ? BDS_CODE_B1C_DATA_PRIMARY : BDS_CODE_B2A_DATA, in.PRN, chips);
        pack_bits(chips, &bits[(size_t)(s * 2) * kPackedWords]);
        primary_bits(signal == BDS_SIG_B1C ? BDS_CODE_B1C_PILOT_PRIMARY : BDS_CODE_B2A_PILOT, in.PRN, chips);
        pack_bits(chips, &bits[(size_t)(s * 2 + 1) * kPackedWords]);
    }
    SynthSat* dS = nullptr;
    uint32_t* dB = nullptr;
    int8_t* dOut = out;
    auto cleanup = [&]() {
        cudaFree(dS);
        cudaFree(dB);
        if (out_loc == BDS_LOC_HOST && dOut != out) cudaFree(dOut);
    };
#define TRYS(x_)                                                    \
    if ((x_) != cudaSuccess) {                                      \
        cleanup();                                                  \
        return set_error(BDS_ERR_CUDA, "synth: %s failed", #x_);    \
    }
    TRYS(cudaMalloc(&dS, sizeof(SynthSat) * n_sats));
    TRYS(cudaMalloc(&dB, bits.size() * 4));
    TRYS(cudaMemcpy(dS, hs.data(), sizeof(SynthSat) * n_sats, cudaMemcpyHostToDevice));
    TRYS(cudaMemcpy(dB, bits.data(), bits.size() * 4, cudaMemcpyHostToDevice));
    if (out_loc == BDS_LOC_HOST) {
        dOut = nullptr;
        TRYS(cudaMalloc(&dOut, n));
    }
    const size_t chunk = (size_t)1 << 28;
    for (size_t off = 0; off < n; off += chunk) {
        size_t m = std::min(chunk, n - off);
        synth_kernel<<<(unsigned)((m + 255) / 256), 256>>>(signal, dS, n_sats, dB, (float)noise_sigma, seed,
                                                           first_sample + (long long)off, m, dOut + off);
        count_launch();
    }
    TRYS(cudaGetLastError());
    TRYS(cudaDeviceSynchronize());
    if (out_loc == BDS_LOC_HOST) TRYS(cudaMemcpy(out, dOut, n, cudaMemcpyDeviceToHost));
#undef TRYS
    cleanup();
    return BDS_OK;
}

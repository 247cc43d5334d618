// Library lifetime, error reporting, code-generation entry points and raw device helpers.
#include <cmath>
#include <cstring>
#include <vector>

#include "bds_codes.h"
#include "bds_common.cuh"

namespace bds {

thread_local std::string g_last_error;
std::atomic<long long> g_launches{0};
int g_device = -1;
int g_num_sms = 0;

int set_error(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

int require_device() {
    if (g_device >= 0) {
        cudaError_t e = cudaSetDevice(g_device);
        if (e != cudaSuccess) return set_error(BDS_ERR_CUDA, "cudaSetDevice(%d): %s", g_device, cudaGetErrorString(e));
        return BDS_OK;
    }
    return bds_init(0);
}

}  // namespace bds

using namespace bds;

extern "C" {

int bds_abi_version(void) { return BDS_ABI_VERSION; }

int bds_init(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return set_error(BDS_ERR_NO_DEVICE, "no CUDA device available (%s); libbdsgpu has no CPU fallback",
                         e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return set_error(BDS_ERR_ARG, "device ordinal %d out of range [0,%d)", device, n);
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return set_error(BDS_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return set_error(BDS_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                         prop.major, prop.minor);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return set_error(BDS_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    g_device = device;
    g_num_sms = prop.multiProcessorCount;
    return BDS_OK;
}

void bds_shutdown(void) {
    if (g_device >= 0) {
        cudaDeviceSynchronize();
        acq_pool_release();
        acq_streams_release();
    }
    g_device = -1;
}

const char* bds_last_error(void) { return g_last_error.c_str(); }
long long bds_launch_count(void) { return g_launches.load(); }

int bds_device_ok(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return 0;
    cudaDeviceProp prop;
    int d = g_device >= 0 ? g_device : 0;
    if (cudaGetDeviceProperties(&prop, d) != cudaSuccess) return 0;
    return prop.major == 10;
}

int bds_gen_code(int component, int prn, int8_t* out, int n) {
    if (!out) return set_error(BDS_ERR_ARG, "null output");
    int len = component_length(component);
    if (len < 0) return set_error(BDS_ERR_ARG, "unknown code component %d", component);
    if (n != len) return set_error(BDS_ERR_ARG, "component %d has %d elements, caller passed %d", component, len, n);
    std::vector<int8_t> v;
    if (!gen_component(component, prn, v)) return set_error(BDS_ERR_ARG, "PRN %d out of range 1..63", prn);
    std::memcpy(out, v.data(), len);
    return BDS_OK;
}

int bds_make_code_table(int component, int prn, double fs, double codeFreqBasis, int codeLength, int8_t* out,
                        int n_samples) {
    if (!out) return set_error(BDS_ERR_ARG, "null output");
    bool b1c = component == BDS_CODE_B1C_DATA_BOC11 || component == BDS_CODE_B1C_PILOT_BOC11;
    bool b2a = component == BDS_CODE_B2A_DATA || component == BDS_CODE_B2A_PILOT;
    if (!b1c && !b2a) return set_error(BDS_ERR_ARG, "component %d has no sampled table in the reference", component);
    if (codeLength != kCodeLen) return set_error(BDS_ERR_UNSUPPORTED, "codeLength must be 10230");
    long spc = std::lround(fs / (codeFreqBasis / codeLength));
    if (n_samples != spc) return set_error(BDS_ERR_ARG, "n_samples %d != samplesPerCode %ld", n_samples, spc);
    std::vector<int8_t> code;
    if (!gen_component(component, prn, code)) return set_error(BDS_ERR_ARG, "PRN %d out of range 1..63", prn);
    // makeDataTable.m:49-63: ts = 1/fs; tc = 1/codeFreqBasis/2 (B1C half chip) or 1/codeFreqBasis (B2a);
    // idx = ceil((ts*k)/tc), k = 1..spc; idx(end) forced to the last element; B1C also forces idx(1) = 1.
    const double ts = 1.0 / fs;
    const double tc = b1c ? (1.0 / codeFreqBasis) / 2.0 : 1.0 / codeFreqBasis;
    const long last = (long)code.size();
    for (long k = 1; k <= spc; ++k) {
        long idx = (long)std::ceil((ts * (double)k) / tc);
        if (k == spc) idx = last;
        if (b1c && k == 1) idx = 1;
        if (idx < 1 || idx > last) return set_error(BDS_ERR_ARG, "code index out of range (fs/codeFreqBasis mismatch)");
        out[k - 1] = code[idx - 1];
    }
    return BDS_OK;
}

int bds_dev_alloc(void** p, size_t bytes) {
    int rc = require_device();
    if (rc) return rc;
    BDS_CUDA(cudaMalloc(p, bytes));
    return BDS_OK;
}
int bds_dev_free(void* p) {
    BDS_CUDA(cudaFree(p));
    return BDS_OK;
}
int bds_host_alloc_pinned(void** p, size_t bytes) {
    int rc = require_device();
    if (rc) return rc;
    BDS_CUDA(cudaMallocHost(p, bytes));
    return BDS_OK;
}
int bds_host_free_pinned(void* p) {
    BDS_CUDA(cudaFreeHost(p));
    return BDS_OK;
}
int bds_memcpy_h2d(void* dst, const void* src, size_t bytes) {
    BDS_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    return BDS_OK;
}
int bds_memcpy_d2h(void* dst, const void* src, size_t bytes) {
    BDS_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return BDS_OK;
}
int bds_dev_sync(void) {
    BDS_CUDA(cudaDeviceSynchronize());
    return BDS_OK;
}

}  // extern "C"

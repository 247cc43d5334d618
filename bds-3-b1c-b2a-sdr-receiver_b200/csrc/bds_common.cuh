// Shared host/device helpers for libbdsgpu (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/bdsgpu.h"

namespace bds {

extern thread_local std::string g_last_error;
extern std::atomic<long long> g_launches;
extern int g_device;          // -1 until bds_init succeeds
extern int g_num_sms;

int set_error(int code, const char* fmt, ...);
// Fails loudly (BDS_ERR_NO_DEVICE) when no usable sm_100 device: no CPU fallback.
int require_device();

#define BDS_CUDA(expr)                                                                        \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return ::bds::set_error(BDS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,             \
                                    cudaGetErrorString(_e), __FILE__, __LINE__);              \
    } while (0)

// frees the device work buffers pooled by bds_acquire (bds_acq.cu)
void acq_pool_release();
void acq_streams_release();

inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- device-side memory-model helpers --------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif

}  // namespace bds

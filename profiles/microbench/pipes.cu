// Instruction-throughput microbenchmark for the correlator inner loop (B200, sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
#define UNROLL 8
template <int OP>
__global__ void __launch_bounds__(512, 1) bench(int* out, int seed) {
    int a[UNROLL], b = seed + threadIdx.x, c = seed * 3 + 1;
    float fa[UNROLL], fb = (float)b * 1e-3f, fc = 1.0001f;
    unsigned long long pa[UNROLL];
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) { a[i] = i + b; fa[i] = (float)(i + b); pa[i] = ((unsigned long long)(i + b) << 32) | (unsigned)(b * 7 + i); }
    unsigned long long pb = ((unsigned long long)__float_as_uint(1.0001f) << 32) | __float_as_uint(0.9999f);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < UNROLL; ++i) {
            if (OP == 0) asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (OP == 1) asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (OP == 2) asm volatile("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (OP == 3) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(fa[i]) : "f"(fb), "f"(fc));
            if (OP == 4) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(pa[i]) : "l"(pb), "l"(pb));
            if (OP == 5) asm volatile("prmt.b32 %0, %0, %1, 0x9991;" : "+r"(a[i]) : "r"(b));
            if (OP == 6) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (OP == 7) asm volatile("add.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            if (OP == 8) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(fa[i]) : "f"(fb));
            if (OP == 9) { asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c)); asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(fa[i]) : "f"(fb), "f"(fc)); }
            if (OP == 10) { asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[(i + 4) % UNROLL]) : "r"(b), "r"(c)); }
            if (OP == 11) { asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[(i + 4) % UNROLL]) : "r"(b), "r"(c)); }
            if (OP == 12) asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(fa[i]) : "r"(a[i]));
        }
    }
    int s = 0;
    float fs = 0;
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) { s += a[i] + (int)pa[i]; fs += fa[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (int)fs;
}
template <int OP>
void run(const char* name, int perIter, int* d) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench<OP><<<148, 512>>>(d, 3);
    cudaEventRecord(e0);
    bench<OP><<<148, 512>>>(d, 3);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double winst = 148.0 * 16 * ITERS * UNROLL * perIter;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-28s %7.3f ms  %6.2f warp-inst/clk/SM (at %d MHz nominal)\n", name, ms, winst / (ms * 1e-3) / 148.0 / (clk * 1e3), clk / 1000);
}
int main() {
    int* d; cudaMalloc(&d, 148 * 512 * 4);
    run<0>("IMAD", 1, d); run<1>("IDP.2A", 1, d); run<2>("IDP.4A", 1, d); run<3>("FFMA", 1, d); run<4>("FFMA2", 1, d);
    run<5>("PRMT", 1, d); run<6>("LOP3", 1, d); run<7>("IADD", 1, d); run<8>("FADD", 1, d);
    run<9>("IMAD+FFMA", 2, d); run<10>("IMAD+LOP3", 2, d); run<11>("IDP.2A+LOP3", 2, d); run<12>("I2F", 1, d);
    return 0;
}

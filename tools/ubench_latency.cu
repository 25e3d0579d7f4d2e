// tools/ubench_latency.cu -- development aid: dependent-issue latency (cycles per instruction of a single dependent chain,
// one warp) of FFMA, FFMA2 (fma.rn.f32x2), FHFMA (fma.rn.f32.f16) and of the dequant+accumulate step on sm_100a.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define N 4096
__device__ __forceinline__ float fhfma(uint32_t a, uint32_t s, float c) {
    float d;
    asm volatile("{ .reg .b16 al, ah, sl, sh; mov.b32 {al, ah}, %1; mov.b32 {sl, sh}, %2; fma.rn.f32.f16 %0, al, sl, %3; }" : "=f"(d) : "r"(a), "r"(s), "f"(c));
    return d;
}
template <int MODE>
__global__ void k(float* out, long long* cyc, float x, uint32_t u) {
    float a = threadIdx.x * 1e-3f, b = a + 1.0f;
    unsigned long long acc;
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(acc) : "f"(a), "f"(b));
    unsigned long long xx;
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(xx) : "f"(x));
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) {
        if (MODE == 0) a = __fmaf_rn(a, x, b);
        if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(acc) : "l"(xx));
        if (MODE == 2) a = fhfma(u, u, a);
        if (MODE == 3) {   // chain step of the kernel: acc = fma2((w0,w1), xx, acc), w from FHFMA of independent inputs
            const float w0 = fhfma(u + i, u, x), w1 = fhfma(u + i, u, b);
            unsigned long long w;
            asm volatile("mov.b64 %0, {%1, %2};" : "=l"(w) : "f"(w0), "f"(w1));
            asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(w), "l"(xx));
        }
        if (MODE == 4) {   // same with two scalar FFMA chains
            const float w0 = fhfma(u + i, u, x), w1 = fhfma(u + i, u, x);
            a = __fmaf_rn(w0, x, a); b = __fmaf_rn(w1, x, b);
        }
    }
    long long t1 = clock64();
    float lo, hi;
    asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc));
    out[threadIdx.x] = a + b + lo + hi;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE> void run(const char* name) {
    float* out; long long* cyc;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
    k<MODE><<<1, 32>>>(out, cyc, 1.0001f, 0x3c003c01u);
    k<MODE><<<1, 32>>>(out, cyc, 1.0001f, 0x3c003c01u);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-44s %6.2f cycles per dependent step (err %d)\n", name, (double)c / N, (int)cudaGetLastError());
}
int main() {
    run<0>("FFMA chain");
    run<1>("FFMA2 chain");
    run<2>("FHFMA chain");
    run<3>("step: 2 FHFMA (independent) + FFMA2 chain");
    run<4>("step: 2 FHFMA (independent) + 2 FFMA chains");
    return 0;
}

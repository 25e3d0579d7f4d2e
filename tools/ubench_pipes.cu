// tools/ubench_pipes.cu -- development aid: issue/pipe throughput on sm_100a of the instructions the INT4
// GEMV inner loop is made of (FFMA, FFMA2, FHFMA = fma.rn.f32.f16, LOP3, HFMA2, cvt) and of the exact mix.
// Prints lane-ops per clock per SM.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench_pipes ubench_pipes.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ITERS 4096
#define CH 8

__device__ __forceinline__ float fhfma(uint32_t a, uint32_t s, float c) {
    float d;
    asm volatile("{ .reg .b16 al, ah, sl, sh;\n\t"
        "mov.b32 {al, ah}, %1;\n\t"
        "mov.b32 {sl, sh}, %2;\n\t"
        "fma.rn.f32.f16 %0, al, sl, %3; }"
        : "=f"(d) : "r"(a), "r"(s), "f"(c));
    return d;
}
__device__ __forceinline__ float fhfma_hi(uint32_t a, uint32_t s, float c) {
    float d;
    asm volatile("{ .reg .b16 al, ah, sl, sh;\n\t"
        "mov.b32 {al, ah}, %1;\n\t"
        "mov.b32 {sl, sh}, %2;\n\t"
        "fma.rn.f32.f16 %0, ah, sl, %3; }"
        : "=f"(d) : "r"(a), "r"(s), "f"(c));
    return d;
}
__device__ __forceinline__ void ffma2(float& a0, float& a1, float w0, float w1, float x) {
    asm volatile("{ .reg .b64 rw, rx, ra;\n\t"
        "mov.b64 rw, {%2, %3};\n\t"
        "mov.b64 rx, {%4, %4};\n\t"
        "mov.b64 ra, {%0, %1};\n\t"
        "fma.rn.f32x2 ra, rw, rx, ra;\n\t"
        "mov.b64 {%0, %1}, ra; }"
        : "+f"(a0), "+f"(a1) : "f"(w0), "f"(w1), "f"(x));
}
__device__ __forceinline__ uint32_t and_or(uint32_t a, uint32_t mask, uint32_t magic) {
    uint32_t d;
    asm volatile("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(mask), "r"(magic));
    return d;
}

// mode 0: FFMA   1: FFMA2   2: FHFMA   3: LOP3   4: HFMA2   5: mix (per 2 weights: 1 LOP3, 2 FHFMA, 1 FFMA2)
// 6: FMUL+FFMA (alt dequant)  7: cvt.f32.f16
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, uint32_t seed, long long* cyc) {
    float a[CH * 2];
    uint32_t u[CH];
    for (int i = 0; i < CH * 2; i++) a[i] = (float)(threadIdx.x + i) * 1e-3f;
    for (int i = 0; i < CH; i++) u[i] = seed * (threadIdx.x + i + 1);
    const float x = __int_as_float(0x3f800001u + (seed & 3)), y = 1e-9f;
    const uint32_t s = 0x2c002c00u + (seed & 7);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int c = 0; c < CH; c++) {
            if (MODE == 0) { a[c] = __fmaf_rn(a[c], x, y); a[c + CH] = __fmaf_rn(a[c + CH], x, y); }
            if (MODE == 1) { ffma2(a[c], a[c + CH], x, y, x); }
            if (MODE == 2) { a[c] = fhfma(u[c], s, a[c]); a[c + CH] = fhfma_hi(u[c], s, a[c + CH]); }
            if (MODE == 3) { u[c] = and_or(u[c], 0x000F000Fu + it, 0x64006400u); }
            if (MODE == 4) {
                __half2 h = *reinterpret_cast<__half2*>(&u[c]);
                h = __hfma2(h, *reinterpret_cast<const __half2*>(&s), h);
                u[c] = *reinterpret_cast<uint32_t*>(&h);
            }
            if (MODE == 5) {
                const uint32_t p = and_or(u[c] + it, 0x000F000Fu, 0x64006400u);
                const float w0 = fhfma(p, s, y), w1 = fhfma_hi(p, s, y);
                ffma2(a[c], a[c + CH], w0, w1, x);
            }
            if (MODE == 6) {
                const float w0 = __fmul_rn(a[c], x);
                a[c + CH] = __fmaf_rn(w0, y, a[c + CH]);
            }
            if (MODE == 7) {
                const __half2 h = *reinterpret_cast<__half2*>(&u[c]);
                a[c] += __low2float(h); a[c + CH] += __high2float(h);
            }
        }
    }
    long long t1 = clock64();
    float r = 0;
    for (int i = 0; i < CH * 2; i++) r += a[i];
    for (int i = 0; i < CH; i++) r += (float)u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, double lane_ops_per_iter_per_thread, int blocks_per_sm) {
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * sms * blocks_per_sm * 256);
    cudaMalloc(&cyc, 8);
    k<MODE><<<sms * blocks_per_sm, 256>>>(out, 3, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<sms * blocks_per_sm, 256>>>(out, 3, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double ops_per_sm = lane_ops_per_iter_per_thread * ITERS * 256.0 * blocks_per_sm;
    printf("%-28s blocks/SM %d: %8.1f lane-ops/clk/SM (block0 clocks %lld, %.3f ms, err %d)\n", name, blocks_per_sm, ops_per_sm / (double)c, c, ms,
           (int)cudaGetLastError());
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int b = 2; b <= 8; b *= 2) {
        run<0>("FFMA", CH * 2, b);
        run<1>("FFMA2 (2 fma each)", CH * 2, b);
        run<2>("FHFMA (fma.rn.f32.f16)", CH * 2, b);
        run<3>("LOP3", CH, b);
        run<4>("HFMA2 (2 fma each)", CH * 2, b);
        run<5>("mix LOP3+2FHFMA+FFMA2 /2w", CH * 2, b);   // counted in weights
        run<6>("FMUL+FFMA /w", CH, b);                      // counted in weights
        run<7>("cvt f16x2->2xf32 + 2 FADD", CH * 2, b);
    }
    return 0;
}

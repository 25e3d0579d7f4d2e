// tools/ubench_pipes2.cu -- development aid: issue throughput of the instructions of the INT4 trip body (FHFMA, FFMA2, FFMA,
// LOP3, SHF) alone and in the trip's mix, in warp-instructions per clock per SM sub-partition.  Eight independent chains per
// thread and 16 warps per SM, so dependent-issue latency is hidden and the figure is the pipe (or dispatch) limit.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/ubench_pipes2 tools/ubench_pipes2.cu
#include <stdio.h>
#include <stdint.h>
#include <cuda_runtime.h>

#define FH(d, a, s) asm volatile("{ .reg .b16 l, h, sl, sh; mov.b32 {l, h}, %1; mov.b32 {sl, sh}, %2; fma.rn.f32.f16 %0, l, sl, %0; }" : "+f"(d) : "r"(a), "r"(s))
#define FHH(d, a, s) asm volatile("{ .reg .b16 l, h, sl, sh; mov.b32 {l, h}, %1; mov.b32 {sl, sh}, %2; fma.rn.f32.f16 %0, h, sl, %0; }" : "+f"(d) : "r"(a), "r"(s))
#define FHC(d, a, s, c) asm volatile("{ .reg .b16 l, h, sl, sh; mov.b32 {l, h}, %1; mov.b32 {sl, sh}, %2; fma.rn.f32.f16 %0, l, sl, %3; }" : "=f"(d) : "r"(a), "r"(s), "f"(c))
#define FHHC(d, a, s, c) asm volatile("{ .reg .b16 l, h, sl, sh; mov.b32 {l, h}, %1; mov.b32 {sl, sh}, %2; fma.rn.f32.f16 %0, h, sl, %3; }" : "=f"(d) : "r"(a), "r"(s), "f"(c))
#define F2(d, w, x) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(w), "l"(x))
#define FF(d, w, x) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(d) : "f"(w), "f"(x))
#define LO(d, a) asm volatile("lop3.b32 %0, %0, %1, 0x64006400, 0xEA;" : "+r"(d) : "r"(a))
#define SH(d) asm volatile("shr.u32 %0, %0, 1;" : "+r"(d))

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* cyc, int iters, uint32_t seed) {
    float f[8];
    unsigned long long p[8];
    uint32_t u[8];
    for (int i = 0; i < 8; i++) { f[i] = (float)(threadIdx.x + i); p[i] = 0x3f8000003f800000ull + i + threadIdx.x; u[i] = seed * (i + 1) + threadIdx.x; }
    const uint32_t a = 0x3c003c00u ^ seed, s = 0x3c003c01u ^ seed;
    const unsigned long long w = 0x3f8000013f800001ull ^ seed, x = 0x3f7fffff3f7fffffull ^ seed;
    const float wf = __uint_as_float(0x3f800001u ^ seed), xf = __uint_as_float(0x3f7fffffu ^ seed);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 8
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) { FH(f[i], a, s); }
            if (MODE == 1) { F2(p[i], w, x); }
            if (MODE == 2) { FF(f[i], wf, xf); }
            if (MODE == 3) { LO(u[i], a); }
            if (MODE == 4) { SH(u[i]); }
            if (MODE == 5) { FH(f[i], a, s); FH(f[i], s, a); LO(u[i], a); }                                     // dequant: 2 FHFMA + 1 ALU
            if (MODE == 6) { FH(f[i], a, s); FH(f[i], s, a); FH(f[i], a, a); FH(f[i], s, s); F2(p[i], w, x); F2(p[i], x, w); }      // FMA pipe of the trip: 4 FHFMA + 2 FFMA2
            if (MODE == 7) { FH(f[i], a, s); FH(f[i], s, a); FH(f[i], a, a); FH(f[i], s, s); F2(p[i], w, x); F2(p[i], x, w); LO(u[i], a); LO(u[i], s); SH(u[i]); }   // the trip's mix 4:2:3
            if (MODE == 8) { FH(f[i], a, s); FF(f[i], wf, xf); }
            if (MODE == 9) { F2(p[i], w, x); LO(u[i], a); }
            if (MODE == 10) { FF(f[i], wf, xf); LO(u[i], a); }
            if (MODE == 12) { FHH(f[i], a, s); }
            if (MODE == 13) { FH(f[i], a, s); FHH(f[i], s, a); }
            if (MODE == 14) {      // the real dequant8 shape: 4 LOP3 + 1 SHF + 8 FHFMA (4 of them .H1) with fresh destinations, + 8 FADD to consume
                uint32_t p04, p15, p26, p37, w8;
                asm volatile("lop3.b32 %0, %1, 0x000F000F, 0x64006400, 0xEA;" : "=r"(p04) : "r"(u[i]));
                asm volatile("lop3.b32 %0, %1, 0x00F000F0, 0x54005400, 0xEA;" : "=r"(p15) : "r"(u[i]));
                asm volatile("shr.u32 %0, %1, 8;" : "=r"(w8) : "r"(u[i]));
                asm volatile("lop3.b32 %0, %1, 0x000F000F, 0x64006400, 0xEA;" : "=r"(p26) : "r"(w8));
                asm volatile("lop3.b32 %0, %1, 0x00F000F0, 0x54005400, 0xEA;" : "=r"(p37) : "r"(w8));
                float d0, d1, d2, d3, d4, d5, d6, d7;
                FHC(d0, p04, s, wf); FHC(d1, p15, s, xf); FHC(d2, p26, s, wf); FHC(d3, p37, s, xf);
                FHHC(d4, p04, s, wf); FHHC(d5, p15, s, xf); FHHC(d6, p26, s, wf); FHHC(d7, p37, s, xf);
                asm volatile("add.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(d0)); asm volatile("add.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(d1));
                asm volatile("add.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(d2)); asm volatile("add.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(d3));
                asm volatile("add.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(d4)); asm volatile("add.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(d5));
                asm volatile("add.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(d6)); asm volatile("add.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(d7));
                u[i] += 0x01234567u;
            }
            if (MODE == 15) {      // the same with the accumulate of the trip: 8 FFMA on one chain
                uint32_t p04, p15, p26, p37, w8;
                asm volatile("lop3.b32 %0, %1, 0x000F000F, 0x64006400, 0xEA;" : "=r"(p04) : "r"(u[i]));
                asm volatile("lop3.b32 %0, %1, 0x00F000F0, 0x54005400, 0xEA;" : "=r"(p15) : "r"(u[i]));
                asm volatile("shr.u32 %0, %1, 8;" : "=r"(w8) : "r"(u[i]));
                asm volatile("lop3.b32 %0, %1, 0x000F000F, 0x64006400, 0xEA;" : "=r"(p26) : "r"(w8));
                asm volatile("lop3.b32 %0, %1, 0x00F000F0, 0x54005400, 0xEA;" : "=r"(p37) : "r"(w8));
                float d0, d1, d2, d3, d4, d5, d6, d7;
                FHC(d0, p04, s, wf); FHC(d1, p15, s, xf); FHC(d2, p26, s, wf); FHC(d3, p37, s, xf);
                FHHC(d4, p04, s, wf); FHHC(d5, p15, s, xf); FHHC(d6, p26, s, wf); FHHC(d7, p37, s, xf);
                FF(f[i], d0, xf); FF(f[i], d1, wf); FF(f[i], d2, xf); FF(f[i], d3, wf); FF(f[i], d4, xf); FF(f[i], d5, wf); FF(f[i], d6, xf); FF(f[i], d7, wf);
                u[i] += 0x01234567u;
            }
            if (MODE == 11) { FH(f[i], a, s); FH(f[i], s, a); FH(f[i], a, a); FH(f[i], s, s); FF(f[i], wf, xf); FF(f[i], xf, wf); FF(f[i], wf, wf); FF(f[i], xf, xf); LO(u[i], a); LO(u[i], s); SH(u[i]); }   // FFMA instead of FFMA2
        }
    }
    const long long t1 = clock64();
    float acc = 0.f;
    for (int i = 0; i < 8; i++) acc += f[i] + (float)(p[i] >> 40) + (float)u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_iter, float* out, long long* cyc, int sms) {
    const int iters = 2048;
    k<MODE><<<sms, 512>>>(out, cyc, iters, 0);
    cudaDeviceSynchronize();
    k<MODE><<<sms, 512>>>(out, cyc, iters, 0);
    cudaDeviceSynchronize();
    long long h[256];
    cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double mx = 0;
    for (int b = 0; b < sms; b++) if ((double)h[b] > mx) mx = (double)h[b];
    const double instr = (double)iters * 8 * per_iter * 4;      // warp-instructions per sub-partition (16 warps / 4)
    printf("%-52s %6.3f warp-instr/clk/sub-partition  (%5.2f clk per group of %d)  err %d\n", name, instr / mx, mx / ((double)iters * 8 * 4), per_iter, (int)cudaGetLastError());
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * sms * 512); cudaMalloc(&cyc, sizeof(long long) * 256);
    run<0>("FHFMA", 1, out, cyc, sms);
    run<1>("FFMA2", 1, out, cyc, sms);
    run<2>("FFMA", 1, out, cyc, sms);
    run<3>("LOP3", 1, out, cyc, sms);
    run<4>("SHF", 1, out, cyc, sms);
    run<5>("2 FHFMA + 1 LOP3", 3, out, cyc, sms);
    run<6>("4 FHFMA + 2 FFMA2", 6, out, cyc, sms);
    run<7>("4 FHFMA + 2 FFMA2 + 2 LOP3 + 1 SHF (trip mix)", 9, out, cyc, sms);
    run<8>("FHFMA + FFMA", 2, out, cyc, sms);
    run<9>("FFMA2 + LOP3", 2, out, cyc, sms);
    run<10>("FFMA + LOP3", 2, out, cyc, sms);
    run<11>("4 FHFMA + 4 FFMA + 2 LOP3 + 1 SHF", 11, out, cyc, sms);
    run<12>("FHFMA .H1", 1, out, cyc, sms);
    run<13>("FHFMA + FHFMA .H1", 2, out, cyc, sms);
    run<14>("dequant8 (4 LOP3, SHF, 8 FHFMA) + 8 FADD + IADD", 22, out, cyc, sms);
    run<15>("dequant8 (4 LOP3, SHF, 8 FHFMA) + 8 FFMA + IADD", 22, out, cyc, sms);
    return 0;
}

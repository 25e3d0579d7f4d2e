// tools/ubench_trip.cu -- development aid: the INT4 trip loop of interp_sm100.cuh (q4_trip2 + col_meta) alone, fed from
// shared memory that is filled once, at the production launch geometry (one CTA of 12 warps per SM, 168 registers).  Prints
// cycles per warp-trip and weights per clock per SM for 1..11 active consumer warps and for alternative trip bodies.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o build/ubench_trip tools/ubench_trip.cu
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../llama_cu_awq_b200/csrc/interp_sm100.cuh"

using namespace lq4;

// ---- variant 1: the x pairs of a whole quarter-trip (qi) are loaded before its arithmetic starts ----
__device__ __forceinline__ void q4_trip2_v1(unsigned long long& acc0, unsigned long long& acc1, uint32_t xaddr, uint32_t w0, uint32_t w1,
                                            const ColMeta& m0, const ColMeta& m1) {
    const uint4 wa0 = lds_v4(w0), wb0 = lds_v4(w0 ^ 16), wa1 = lds_v4(w1), wb1 = lds_v4(w1 ^ 16);
    unsigned long long xp[2][8];
#pragma unroll
    for (int e2 = 0; e2 < 4; e2++) lds_v2_b64(xaddr + e2 * kRowBytes, xp[0][2 * e2], xp[0][2 * e2 + 1]);
#pragma unroll
    for (int qi = 0; qi < 4; qi++) {
        if (qi < 3) {
#pragma unroll
            for (int e2 = 0; e2 < 4; e2++) lds_v2_b64(xaddr + ((qi + 1) * 4 + e2) * kRowBytes, xp[(qi + 1) & 1][2 * e2], xp[(qi + 1) & 1][2 * e2 + 1]);
        }
        float da0[8], db0[8], da1[8], db1[8];
        dequant8(da0, word_of(wa0, qi), m0.s16, m0.nlo, m0.nhi);
        dequant8(db0, word_of(wb0, qi), m0.s16, m0.nlo, m0.nhi);
        dequant8(da1, word_of(wa1, qi), m1.s16, m1.nlo, m1.nhi);
        dequant8(db1, word_of(wb1, qi), m1.s16, m1.nlo, m1.nhi);
#pragma unroll
        for (int e = 0; e < 8; e++) {
            ffma2_pk(acc0, da0[e], db0[e], xp[qi & 1][e]);
            ffma2_pk(acc1, da1[e], db1[e], xp[qi & 1][e]);
        }
    }
}

// ---- variant 2: plain FFMA on four separate chains instead of FFMA2 ----
__device__ __forceinline__ void q4_trip2_v2(unsigned long long& acc0, unsigned long long& acc1, uint32_t xaddr, uint32_t w0, uint32_t w1,
                                            const ColMeta& m0, const ColMeta& m1) {
    const uint4 wa0 = lds_v4(w0), wb0 = lds_v4(w0 ^ 16), wa1 = lds_v4(w1), wb1 = lds_v4(w1 ^ 16);
    float a0, b0, a1, b1;
    unpack_f2(acc0, a0, b0); unpack_f2(acc1, a1, b1);
#pragma unroll
    for (int qi = 0; qi < 4; qi++) {
        float da0[8], db0[8], da1[8], db1[8];
        dequant8(da0, word_of(wa0, qi), m0.s16, m0.nlo, m0.nhi);
        dequant8(db0, word_of(wb0, qi), m0.s16, m0.nlo, m0.nhi);
        dequant8(da1, word_of(wa1, qi), m1.s16, m1.nlo, m1.nhi);
        dequant8(db1, word_of(wb1, qi), m1.s16, m1.nlo, m1.nhi);
#pragma unroll
        for (int e2 = 0; e2 < 4; e2++) {
            unsigned long long xp0, xp1;
            lds_v2_b64(xaddr + (qi * 4 + e2) * kRowBytes, xp0, xp1);
            float xa0, xb0, xa1, xb1;
            unpack_f2(xp0, xa0, xb0); unpack_f2(xp1, xa1, xb1);
            a0 = __fmaf_rn(da0[2 * e2], xa0, a0); b0 = __fmaf_rn(db0[2 * e2], xb0, b0);
            a1 = __fmaf_rn(da1[2 * e2], xa0, a1); b1 = __fmaf_rn(db1[2 * e2], xb0, b1);
            a0 = __fmaf_rn(da0[2 * e2 + 1], xa1, a0); b0 = __fmaf_rn(db0[2 * e2 + 1], xb1, b0);
            a1 = __fmaf_rn(da1[2 * e2 + 1], xa1, a1); b1 = __fmaf_rn(db1[2 * e2 + 1], xb1, b1);
        }
    }
    acc0 = pack_f2(a0, b0); acc1 = pack_f2(a1, b1);
}

// ---- variant 3: dequantisation only (no accumulate): what the FHFMA + LOP3 half costs ----
__device__ __forceinline__ void q4_trip2_v3(unsigned long long& acc0, unsigned long long& acc1, uint32_t xaddr, uint32_t w0, uint32_t w1,
                                            const ColMeta& m0, const ColMeta& m1) {
    const uint4 wa0 = lds_v4(w0), wb0 = lds_v4(w0 ^ 16), wa1 = lds_v4(w1), wb1 = lds_v4(w1 ^ 16);
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int qi = 0; qi < 4; qi++) {
        float da0[8], db0[8], da1[8], db1[8];
        dequant8(da0, word_of(wa0, qi), m0.s16, m0.nlo, m0.nhi);
        dequant8(db0, word_of(wb0, qi), m0.s16, m0.nlo, m0.nhi);
        dequant8(da1, word_of(wa1, qi), m1.s16, m1.nlo, m1.nhi);
        dequant8(db1, word_of(wb1, qi), m1.s16, m1.nlo, m1.nhi);
#pragma unroll
        for (int e = 0; e < 8; e++) { s0 += da0[e]; s1 += db0[e]; s0 += da1[e]; s1 += db1[e]; }   // 1 FADD per weight instead of 1/2 FFMA2
    }
    float a, b; unpack_f2(acc0, a, b); acc0 = pack_f2(a + s0, b + s1);
}

// ---- variant 4: accumulate only (weights taken as raw fp32 bit patterns: no dequantisation) ----
__device__ __forceinline__ void q4_trip2_v4(unsigned long long& acc0, unsigned long long& acc1, uint32_t xaddr, uint32_t w0, uint32_t w1,
                                            const ColMeta& m0, const ColMeta& m1) {
    const uint4 wa0 = lds_v4(w0), wb0 = lds_v4(w0 ^ 16), wa1 = lds_v4(w1), wb1 = lds_v4(w1 ^ 16);
#pragma unroll
    for (int qi = 0; qi < 4; qi++) {
        const float fa0 = __uint_as_float(word_of(wa0, qi)), fb0 = __uint_as_float(word_of(wb0, qi));
        const float fa1 = __uint_as_float(word_of(wa1, qi)), fb1 = __uint_as_float(word_of(wb1, qi));
#pragma unroll
        for (int e2 = 0; e2 < 4; e2++) {
            unsigned long long xp0, xp1;
            lds_v2_b64(xaddr + (qi * 4 + e2) * kRowBytes, xp0, xp1);
            ffma2_pk(acc0, fa0, fb0, xp0);
            ffma2_pk(acc1, fa1, fb1, xp0);
            ffma2_pk(acc0, fb0, fa0, xp1);
            ffma2_pk(acc1, fb1, fa1, xp1);
        }
    }
}

// ---- variant 5: another work split: a thread is ONE reference lane of FOUR columns (instead of two lanes of two columns), so a
// loaded x is used by four columns (8 instead of 16 x loads per trip) and the accumulate is the scalar FFMA.  x rows are plain
// fp32 in k order, one 128-byte row per lane padded to 144 bytes (conflict-free 16-byte loads). ----
constexpr int kRowBytes5 = 144, kTripBytes5 = 32 * kRowBytes5;
__device__ __forceinline__ void q4_trip4_v5(float (&acc)[4], uint32_t xaddr, const uint32_t (&w)[4], const ColMeta (&m)[4]) {
    uint4 wv[4];
#pragma unroll
    for (int c = 0; c < 4; c++) wv[c] = lds_v4(w[c]);
#pragma unroll
    for (int qi = 0; qi < 4; qi++) {
        const uint4 xa = lds_v4(xaddr + qi * 32), xb = lds_v4(xaddr + qi * 32 + 16);
        const float x[8] = {__uint_as_float(xa.x), __uint_as_float(xa.y), __uint_as_float(xa.z), __uint_as_float(xa.w),
                            __uint_as_float(xb.x), __uint_as_float(xb.y), __uint_as_float(xb.z), __uint_as_float(xb.w)};
#pragma unroll
        for (int c = 0; c < 4; c++) {
            float d[8];
            dequant8(d, word_of(wv[c], qi), m[c].s16, m[c].nlo, m[c].nhi);
#pragma unroll
            for (int e = 0; e < 8; e++) acc[c] = __fmaf_rn(d[e], x[e], acc[c]);
        }
    }
}

// MAXW: warps per CTA (consumers + the producer lane's warp).  12 is the production kernel (168 registers per thread); -DMAXW=24
// asks what a kernel that fitted 85 registers would gain from 23 consumer warps.
#ifndef MAXW
#define MAXW 12
#endif
static const int kActive[] = {1, 2, 4, 7, 8, 11, 12, 15, 16, 19, 20, 23};

template <int VAR>
__global__ void __launch_bounds__(32 * MAXW, 1) trip_kernel(const uint32_t* __restrict__ src, float* out, long long* cyc, int T, int ntasks, int active, int prod) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int colb = T * 512;
    const int xs_bytes = T * kTripBytes5, w_bytes = (MAXW - 1) * 4 * colb, meta_bytes = (MAXW - 1) * 4 * 64;
    uint32_t* s32 = reinterpret_cast<uint32_t*>(smem);
    for (int i = tid; i < (xs_bytes + w_bytes + meta_bytes) / 4; i += blockDim.x) {
        uint32_t v = src[(i * 7 + blockIdx.x) & 0xFFFFF];
        if (i < xs_bytes / 4) v = (v & 0x007FFFFFu) | 0x3C000000u;             // x: small positive floats
        else if (i >= (xs_bytes + w_bytes) / 4) v = (v & 0x03FF03FFu) | 0x20002000u;   // scales: small fp16, zeros: whatever
        s32[i] = v;
    }
    // the production kernel's twelfth warp: one lane keeping bulk copies in flight into shared memory (prod = 2: eight 8 KB copies
    // at a time, about the ring's fill rate; prod = 1: only spinning on an mbarrier that never completes)
    volatile int* stop = reinterpret_cast<volatile int*>(smem + xs_bytes + w_bytes + meta_bytes);
    const uint32_t pbar = smem_u32(smem) + xs_bytes + w_bytes + meta_bytes + 64, pring = pbar + 64 + 896;   // 1 KB after the meta area
    if (tid == 0) { stop[0] = 0; stop[1] = 0; for (int i = 0; i < 8; i++) mbar_init(pbar + i * 8, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    if (warp == MAXW - 1) {
        if (lane != 0 || prod == 0) return;
        uint64_t policy;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
        if (prod == 1) { while (!*stop) (void)mbar_try_wait(pbar, 0); return; }
        unsigned n = 0;
        while (!*stop) {
            const unsigned sl = n & 7, lap = n >> 3;
            if (lap > 0) while (!mbar_try_wait(pbar + sl * 8, (lap - 1) & 1)) {}
            mbar_arrive_expect_tx(pbar + sl * 8, 8192);
            bulk_g2s(pring + sl * 8192, reinterpret_cast<const uint8_t*>(src) + ((size_t)(n * 148 + blockIdx.x) * 8192 & 0x3FFFFF & ~8191u), 8192, pbar + sl * 8, policy);
            n++;
        }
        for (unsigned k = (n > 8 ? n - 8 : 0); k < n; k++) while (!mbar_try_wait(pbar + (k & 7) * 8, (k >> 3) & 1)) {}      // drain before exit
        cyc[blockIdx.x * MAXW + MAXW - 1] = (long long)n;
        return;
    }
    if (warp >= active) return;
    const uint32_t base = smem_u32(smem);
    const uint32_t xs = base, wb = base + xs_bytes + warp * 4 * colb, mb = base + xs_bytes + w_bytes + warp * 4 * 64;
    const int h = lane >> 4, j = lane & 15, sw = (j >> 2) & 1;
    const uint32_t w0 = wb + (2 * h) * colb + j * 32 + sw * 16, w1 = w0 + colb;
    const uint32_t scol0 = mb + (2 * h) * 64, scol1 = scol0 + 64, zcol0 = scol0 + 32, zcol1 = scol1 + 32;
    float sink = 0.f;
    const long long t0 = clock64();
    if (VAR == 5) {
        const uint32_t wc[4] = {wb + lane * 16, wb + colb + lane * 16, wb + 2 * colb + lane * 16, wb + 3 * colb + lane * 16};
#pragma unroll 1
        for (int task = 0; task < ntasks; task++) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
            for (int t = 0; t < T; t++) {
                ColMeta m[4];
#pragma unroll
                for (int c = 0; c < 4; c++) m[c] = col_meta(mb + c * 64, mb + c * 64 + 32, t & 1, lane >> 1);
                const uint32_t wt[4] = {wc[0] + t * 512, wc[1] + t * 512, wc[2] + t * 512, wc[3] + t * 512};
                q4_trip4_v5(acc, xs + t * kTripBytes5 + lane * kRowBytes5, wt, m);
            }
#pragma unroll
            for (int c = 0; c < 4; c++) {
                float v = acc[c];
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) v = v + __shfl_xor_sync(0xffffffffu, v, d);
                sink += v;
            }
        }
    } else
#pragma unroll 1
    for (int task = 0; task < ntasks; task++) {
        unsigned long long acc0 = 0ull, acc1 = 0ull;
#pragma unroll 1
        for (int t = 0; t < T; t++) {
            const ColMeta m0 = col_meta(scol0, zcol0, t & 1, j), m1 = col_meta(scol1, zcol1, t & 1, j);
            const uint32_t xa = xs + t * kTripBytes + j * 16;
            if (VAR == 0) q4_trip2(acc0, acc1, xa, w0 + t * 512, w1 + t * 512, m0, m1);
            if (VAR == 1) q4_trip2_v1(acc0, acc1, xa, w0 + t * 512, w1 + t * 512, m0, m1);
            if (VAR == 2) q4_trip2_v2(acc0, acc1, xa, w0 + t * 512, w1 + t * 512, m0, m1);
            if (VAR == 3) q4_trip2_v3(acc0, acc1, xa, w0 + t * 512, w1 + t * 512, m0, m1);
            if (VAR == 4) q4_trip2_v4(acc0, acc1, xa, w0 + t * 512, w1 + t * 512, m0, m1);
        }
        sink += halfwarp_total(acc0) + halfwarp_total(acc1);
    }
    const long long t1 = clock64();
    if (lane == 0) {
        cyc[blockIdx.x * MAXW + warp] = t1 - t0;
        int* done = const_cast<int*>(stop) + 1;
        if (atomicAdd(done, 1) == active - 1) { *stop = 1; *done = 0; }
    }
    out[blockIdx.x * blockDim.x + tid] = sink;
}

template <int VAR>
void run(const char* name, const uint32_t* src, float* out, long long* cyc, int sms) {
    const int T = 4, ntasks = 64;
    const size_t smem = (size_t)T * kTripBytes5 + (MAXW - 1) * 4 * T * 512 + (MAXW - 1) * 4 * 64 + 1024 + 8 * 8192;
    cudaFuncSetAttribute(trip_kernel<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, trip_kernel<VAR>);
    for (int prod : {0, 1, 2})
    for (int active : kActive) {
        if (active > MAXW - 1) continue;
        if (prod != 0 && active != 7 && active != 11) continue;
        cudaMemset(cyc, 0, sizeof(long long) * sms * MAXW);
        trip_kernel<VAR><<<sms, 32 * MAXW, smem>>>(src, out, cyc, T, ntasks, active, prod);
        cudaDeviceSynchronize();
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        trip_kernel<VAR><<<sms, 32 * MAXW, smem>>>(src, out, cyc, T, ntasks, active, prod);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        std::vector<long long> h(sms * MAXW);
        cudaMemcpy(h.data(), cyc, sizeof(long long) * sms * MAXW, cudaMemcpyDeviceToHost);
        double sum = 0, mx = 0, copies = 0; int n = 0;
        for (int b = 0; b < sms; b++) { copies += (double)h[b * MAXW + MAXW - 1]; for (int w = 0; w < active; w++) { double c = (double)h[b * MAXW + w]; sum += c; if (c > mx) mx = c; n++; } }
        const double per_trip = sum / n / (ntasks * T);
        printf("%-34s regs %3d  active warps %2d producer %d: %7.0f clk per warp-trip (slowest warp %7.0f), %5.1f weights/clk/SM  [%0.3f ms, err %d", name, fa.numRegs, active, prod, per_trip,
               mx / (ntasks * T), active * 4096.0 / (mx / (ntasks * T)), ms, (int)cudaGetLastError());
        if (prod == 2) printf(", %.1f B/clk/SM copied", copies / sms * 8192.0 / mx);
        printf("]\n");
    }
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t* src; float* out; long long* cyc;
    cudaMalloc(&src, 4 << 20); cudaMalloc(&out, sizeof(float) * sms * 32 * MAXW); cudaMalloc(&cyc, sizeof(long long) * sms * MAXW);
    std::vector<uint32_t> h(1 << 20);
    uint64_t s = 88172645463325252ull;
    for (auto& v : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; v = (uint32_t)s; }
    cudaMemcpy(src, h.data(), 4 << 20, cudaMemcpyHostToDevice);
    run<0>("v0 production q4_trip2", src, out, cyc, sms);
    run<1>("v1 x prefetched per quarter-trip", src, out, cyc, sms);
    run<2>("v2 FFMA instead of FFMA2", src, out, cyc, sms);
    run<3>("v3 dequant only (+FADD)", src, out, cyc, sms);
    run<4>("v4 FFMA2 accumulate only", src, out, cyc, sms);
    run<5>("v5 one lane x four columns, FFMA", src, out, cyc, sms);
    return 0;
}

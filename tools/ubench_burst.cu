// tools/ubench_burst.cu -- development aid: what a ring refill costs.  Every SM's producer lane issues NCH bulk copies of CH
// bytes back to back (no waits in between) from cold HBM, all SMs at the same moment -- the situation after a ring drain or
// after a round of warp-tasks released its slots together.  Prints when each copy was issued and when it had landed.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/ubench_burst tools/ubench_burst.cu
#include <stdio.h>
#include <stdint.h>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include "../llama_cu_awq_b200/csrc/interp_sm100.cuh"

using namespace lq4;

__global__ void __launch_bounds__(384, 1) burst(const uint8_t* __restrict__ src, size_t stride_cta, int ch, int nch, int split, unsigned long long* out, unsigned* sync) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t base = smem_u32(smem);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 32; i++) mbar_init(base + i * 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    // all SMs start together
    atomicAdd(sync, 1u);
    while (*reinterpret_cast<volatile unsigned*>(sync) < gridDim.x) {}
    const unsigned long long t0 = global_ns();
    const uint8_t* p = src + (size_t)blockIdx.x * stride_cta;
    for (int k = 0; k < nch; k++) {
        mbar_arrive_expect_tx(base + k * 8, (uint32_t)ch);
        for (int s = 0; s < split; s++)
            bulk_g2s(base + 1024 + k * ch + s * (ch / split), p + (size_t)k * ch + s * (ch / split), (uint32_t)(ch / split), base + k * 8, policy);
    }
    out[(size_t)blockIdx.x * 64] = global_ns() - t0;      // all issued (reading the global timer costs ~0.15 us: not once per copy)
    for (int k = 0; k < nch; k++) {
        while (!mbar_try_wait(base + k * 8, 0)) {}
        out[(size_t)blockIdx.x * 64 + 32 + k] = global_ns() - t0;
    }
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const size_t total = 2ull << 30;
    uint8_t* src; cudaMalloc(&src, total); cudaMemset(src, 1, total);
    unsigned long long* out; cudaMalloc(&out, sizeof(unsigned long long) * 64 * sms);
    unsigned* sync; cudaMalloc(&sync, 4);
    cudaFuncSetAttribute(burst, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct Case { int ch, nch, split; } cases[] = {{11008, 14, 1}, {8192, 19, 1}, {8192, 19, 2}, {8192, 11, 1}, {8192, 4, 1}, {16384, 9, 1}, {4096, 32, 1}};
    for (const Case& c : cases) {
        std::vector<unsigned long long> h(64 * sms);
        for (int rep = 0; rep < 3; rep++) {
            cudaMemset(sync, 0, 4);
            // a different, cold region every repetition: 2 GB / 3 per repetition, 4.5 MB apart per CTA
            burst<<<sms, 384, 1024 + c.ch * c.nch>>>(src + (size_t)rep * ((total / 3) & ~(size_t)4095), (size_t)4608 * 1024, c.ch, c.nch, c.split, out, sync);
            cudaDeviceSynchronize();
        }
        cudaMemcpy(h.data(), out, sizeof(unsigned long long) * 64 * sms, cudaMemcpyDeviceToHost);
        printf("%d copies of %d B (%d pieces each) per SM, %.1f MB in all, err %d\n  all issued after (us, median over SMs):", c.nch, c.ch, c.split, (double)c.ch * c.nch * sms / 1e6, (int)cudaGetLastError());
        for (int k = 0; k < 1; k++) { std::vector<double> v; for (int b = 0; b < sms; b++) v.push_back(h[b * 64 + k] / 1000.0); std::sort(v.begin(), v.end()); printf(" %.2f", v[sms / 2]); }
        printf("\n  landed (us, median over SMs): ");
        double last_max = 0;
        for (int k = 0; k < c.nch; k++) { std::vector<double> v; for (int b = 0; b < sms; b++) v.push_back(h[b * 64 + 32 + k] / 1000.0); std::sort(v.begin(), v.end()); printf(" %.2f", v[sms / 2]); if (k == c.nch - 1) last_max = v[sms - 1]; }
        printf("\n  last copy landed on the slowest SM at %.2f us: %.0f GB/s over the burst\n", last_max, (double)c.ch * c.nch * sms / last_max / 1e3);
    }
    return 0;
}

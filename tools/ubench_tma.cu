// tools/ubench_tma.cu -- development aid: how fast can one CTA per SM stream HBM into shared-memory rings with
// 1-D TMA bulk copies (cp.async.bulk + mbarrier), as a function of copy size, copies per slot, ring depth and the
// number of independent rings (producer lanes)?  Consumers only wait and release.  Prints GB/s.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/ubench_tma tools/ubench_tma.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol, int hint) {
    if (hint)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
    else
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

struct Params {
    const uint8_t* src; size_t bytes_per_cta; int copy_bytes, copies_per_slot, slots, rings, hint, stride_mode; unsigned* sink;
};
// ring r of a CTA streams its own contiguous share; slot = copies_per_slot copies of copy_bytes.
// stride_mode 1: the copies of a slot come from addresses `col_stride` apart (like column segments of a matrix)
__global__ void __launch_bounds__(512, 1) k(Params P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int R = P.rings, S = P.slots;
    const uint32_t bars = smem_u32(smem);            // full[R][S], empty[R][S]
    const uint32_t ring = bars + 4096;
    const int slot_bytes = P.copy_bytes * P.copies_per_slot;
    if (threadIdx.x == 0) {
        for (int i = 0; i < R * S; i++) { mbar_init(bars + i * 8, 1); mbar_init(bars + (R * S + i) * 8, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const size_t per_ring = (P.bytes_per_cta / R) & ~(size_t)65535;
    const int nchunks = (int)(((per_ring / P.copies_per_slot) & ~(size_t)4095) * P.copies_per_slot / slot_bytes);
    if (warp == R) {   // producer warp: lane r feeds ring r
        if (lane < R) {
            uint64_t pol;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
            const uint8_t* base = P.src + (size_t)blockIdx.x * P.bytes_per_cta + (size_t)lane * per_ring;
            const size_t col_stride = (per_ring / P.copies_per_slot) & ~(size_t)4095;
            int s = 0, ph = 0;
            for (int c = 0; c < nchunks;) {
                if (mbar_test(bars + (R * S + lane * S + s) * 8, ph ^ 1)) {
                    const uint32_t bar = bars + (lane * S + s) * 8, dst = ring + (lane * S + s) * slot_bytes;
                    mbar_expect(bar, slot_bytes);
                    for (int i = 0; i < P.copies_per_slot; i++) {
                        const uint8_t* src = P.stride_mode ? base + (size_t)i * col_stride + (size_t)c * P.copy_bytes
                                                           : base + (size_t)c * slot_bytes + (size_t)i * P.copy_bytes;
                        bulk(dst + i * P.copy_bytes, src, P.copy_bytes, bar, pol, P.hint);
                    }
                    c++;
                    if (++s == S) { s = 0; ph ^= 1; }
                }
            }
        }
        return;
    }
    if (warp < R) {    // consumer warp r
        int s = 0, ph = 0;
        unsigned acc = 0;
        for (int c = 0; c < nchunks; c++) {
            while (!mbar_try(bars + (warp * S + s) * 8, ph)) {}
            acc += *(volatile unsigned*)(smem + 4096 + (warp * S + s) * slot_bytes + lane * 4);
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + (R * S + warp * S + s) * 8);
            if (++s == S) { s = 0; ph ^= 1; }
        }
        if (acc == 0x12345678u) P.sink[0] = acc;
    }
}

int main(int argc, char** argv) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const size_t per_cta = 24u << 20;     // 24 MiB per CTA: 3.5 GB total, > L2
    uint8_t* src; unsigned* sink;
    cudaMalloc(&src, per_cta * sms); cudaMalloc(&sink, 4);
    cudaMemset(src, 1, per_cta * sms);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    struct Cfg { int copy, cps, slots, rings, hint, stride; };
    Cfg cfgs[] = {
        {512, 4, 5, 15, 1, 1}, {512, 4, 5, 15, 0, 1}, {512, 4, 5, 15, 1, 0}, {2048, 1, 5, 15, 1, 0}, {2048, 1, 5, 15, 0, 0},
        {2048, 4, 1, 15, 1, 0}, {4096, 1, 3, 15, 1, 0}, {8192, 1, 3, 8, 1, 0}, {16384, 1, 3, 4, 1, 0}, {16384, 1, 8, 1, 1, 0},
        {32768, 1, 6, 1, 1, 0}, {32768, 1, 6, 1, 0, 0}, {512, 4, 5, 8, 1, 1}, {512, 4, 5, 4, 1, 1}, {1024, 4, 3, 15, 1, 1}, {2048, 4, 2, 13, 1, 1},
    };
    for (Cfg c : cfgs) {
        Params P{src, per_cta, c.copy, c.cps, c.slots, c.rings, c.hint, c.stride, sink};
        const size_t smem = 4096 + (size_t)c.rings * c.slots * c.copy * c.cps;
        if (smem > 232448) { printf("skip (smem)\n"); continue; }
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<<<sms, 32 * (c.rings + 1), smem>>>(P);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        k<<<sms, 32 * (c.rings + 1), smem>>>(P);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const size_t per_ring = (per_cta / c.rings) & ~(size_t)65535;
        const double bytes = (double)((((per_ring / c.cps) & ~(size_t)4095) * c.cps) / (c.copy * c.cps)) * (c.copy * c.cps) * c.rings * sms;
        printf("copy %6d B x%d/slot, %d slots, %2d rings, hint %d, strided %d: %8.1f GB/s (%.3f ms, in flight/SM %d KB, err %d)\n", c.copy, c.cps,
               c.slots, c.rings, c.hint, c.stride, bytes / ms / 1e6, ms, (int)(c.rings * c.slots * c.copy * c.cps / 1024), (int)cudaGetLastError());
    }
    return 0;
}

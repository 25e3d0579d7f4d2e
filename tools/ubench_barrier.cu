// tools/ubench_barrier.cu -- development aid: cost of a grid-wide barrier among 148 persistent CTAs on B200,
// for several implementations.  Each barrier is preceded by a small global store per CTA and followed by a
// coherent load of another CTA's store (so that the memory ordering is actually exercised and checked).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/ubench_barrier tools/ubench_barrier.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <int MODE>
__global__ void __launch_bounds__(384, 1) k(unsigned* counter, unsigned* data, int iters, unsigned* err, long long* cycles) {
    const int nb = gridDim.x, b = blockIdx.x, tid = threadIdx.x;
    unsigned bad = 0;
    long long t0 = clock64();
    for (int it = 1; it <= iters; it++) {
        if (tid < 32) data[b * 32 + tid] = it;             // this CTA's contribution
        bar_sync(1, 384);
        if (tid == 0) {
            const unsigned target = (unsigned)it * nb;
            if (MODE == 0) {          // threadfence + atomicAdd + volatile poll + threadfence  (the textbook version)
                __threadfence();
                atomicAdd(counter, 1u);
                while (*(volatile unsigned*)counter < target) {}
                __threadfence();
            } else if (MODE == 1) {   // red.release + ld.acquire polls
                asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(counter), "r"(1u) : "memory");
                unsigned v;
                do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while (v < target);
            } else if (MODE == 2) {   // red.release + relaxed polls + acquire fence
                asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(counter), "r"(1u) : "memory");
                unsigned v;
                do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while (v < target);
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
            } else if (MODE == 3) {   // last arriver publishes a flag; everybody polls the flag (one writer)
                unsigned old;
                asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(counter), "r"(1u) : "memory");
                if (old + 1 == target) {
                    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(counter + 32), "r"((unsigned)it) : "memory");
                } else {
                    unsigned v;
                    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter + 32) : "memory"); } while (v < (unsigned)it);
                }
            }
        }
        bar_sync(1, 384);
        // read what another CTA wrote before the barrier
        if (tid < 32) {
            unsigned v;
            const unsigned* p = data + ((b + 37) % nb) * 32 + tid;
            asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
            if (v != (unsigned)it) bad++;
        }
    }
    long long t1 = clock64();
    if (bad) atomicAdd(err, bad);
    if (b == 0 && tid == 0) *cycles = t1 - t0;
}

template <int MODE>
void run(const char* name, int sms) {
    unsigned *counter, *data, *err; long long* cyc;
    cudaMalloc(&counter, 256); cudaMalloc(&data, sms * 32 * 4); cudaMalloc(&err, 4); cudaMalloc(&cyc, 8);
    cudaMemset(counter, 0, 256); cudaMemset(data, 0, sms * 32 * 4); cudaMemset(err, 0, 4);
    const int iters = 2000;
    void* args[] = {&counter, &data, (void*)&iters, &err, &cyc};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    cudaLaunchCooperativeKernel((void*)k<MODE>, dim3(sms), dim3(384), args, 0, 0);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned h_err; cudaMemcpy(&h_err, err, 4, cudaMemcpyDeviceToHost);
    printf("%-60s %7.3f us per barrier (incl. store + remote load), stale reads %u, err %d\n", name, ms * 1000.0f / iters, h_err, (int)cudaGetLastError());
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    run<0>("threadfence + atomicAdd + volatile poll + threadfence", sms);
    run<1>("red.release + ld.acquire polls", sms);
    run<2>("red.release + ld.relaxed polls + fence.acq_rel", sms);
    run<3>("atom.acq_rel; last arriver st.release flag; ld.acquire polls", sms);
    run<1>("red.release + ld.acquire polls (again)", sms);
    return 0;
}

// prefill_sm100.cuh -- batched prefill on the 5th-generation tensor cores (scope row f1, BASELINE.json configs[4]).
//
// The reference has no prefill: prompt tokens go one by one through the decode path (llama2_q4.cu:465-470).  Here a batch of
// M = batch x seq token rows goes through every projection as ONE dense GEMM  Y[M][N] = X[M][K] . dequant(W)[N][K]^T :
//
//   gemm_q4_tc_kernel -- persistent, warp-specialised, one CTA per SM, tile 256 (M) x 128 (N) x 64 (K):
//     warp 0      TMA producer: the X tile (256 rows x 64 k, fp16) by one cp.async.bulk.tensor per stage, 128-byte swizzle
//     warp 1      MMA issuer: one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (two M=128 x N=128 x K=16 MMAs per
//                 k-step, sharing the B tile), accumulators in tensor memory (2 tiles x 2 buffers x 128 columns = all 512)
//     warps 2-9   INT4 -> fp16 dequantisation: packed nibbles + group scale / zero from global memory, (q - z) exact in fp16
//                 (HSUB2 on the 1024+q / 64+q magic forms), x scale with one rounding (HMUL2), written straight into the
//                 K-major SWIZZLE_128B shared-memory layout the MMA descriptor names (no fp16 copy of W ever exists in HBM)
//     warps 10-13 epilogue: tcgen05.ld 32 columns at a time, optional residual add in fp32, fp16 stores
//   Four stages of {A 32 KB, B 16 KB}; mbarriers full_a (TMA tx), full_b (dequant warps), empty (tcgen05.commit),
//   tmem_full / tmem_empty (MMA <-> epilogue), so the epilogue of tile i overlaps the main loop of tile i + 1.
//
// Parity: NOT bit-exact by construction (tensor-core summation order, weights rounded to fp16 once); the bar is fp16
// tolerance against the sequential decode path on the same tokens (tests/test_gpu_prefill.py).
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "interp_sm100.cuh"

namespace lq4pf {

using lq4::smem_u32;
using lq4::mbar_init;
using lq4::mbar_arrive;
using lq4::mbar_arrive_expect_tx;
using lq4::mbar_wait;

constexpr int kBM = 256, kBK = 64;                 // CTA tile rows / k-block; the MMA atom is 128 x BN x 16, BN = 128 or 256
constexpr int kBNmin = 128;
constexpr int kABytes = kBM * kBK * 2;             // 32 KB
constexpr int kDequantWarps = 8, kEpiWarps = 4;
constexpr int kThreads = 32 * (2 + kDequantWarps + kEpiWarps);      // 448
constexpr int kFirstDequantWarp = 2, kFirstEpiWarp = 2 + kDequantWarps;
constexpr int kTmemCols = 512;
// BN = 256 halves the L2 -> SM traffic of the X tiles (measured: at BN = 128 the kernel moved 7.7 TB/s out of L2 and the tensor
// pipe was busy 52 % of the time); its two 128 x 256 accumulators fill the tensor memory, so the epilogue is not double-buffered
template <int BN> struct Geo {
    static constexpr int kBBytes = BN * kBK * 2;                     // 16 / 32 KB
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = (BN == 128) ? 4 : 3;              // 192 KB either way
    static constexpr int kAccBufs = (BN == 128) ? 2 : 1;
    static constexpr size_t kSmemBytes = 1024 + (size_t)kStages * kStageBytes + 1024;    // alignment slack + ring + barriers
    // instruction descriptor of tcgen05.mma.kind::f16 (cute/arch/mma_sm100_desc.hpp, InstrDescriptor): D = F32 (bits 4-5 = 1),
    // A = B = F16 (0), both K-major (0), N >> 3 at bit 17, M >> 4 at bit 24
    static constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
};

// shared-memory matrix descriptor, K-major, SWIZZLE_128B (SmemDescriptor): start address >> 4, leading byte offset 1 (ignored
// for swizzled K-major), stride byte offset = 8 rows x 128 B = 1024 >> 4, version 1 (Blackwell) at bit 46, layout type 2 at bit 61
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{ .reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {      // arrives on `bar` when every MMA issued so far has completed
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                   "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                   "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct GemmParams {
    const uint32_t* w;      // [N][K/8]
    const uint32_t* z;      // [N][zh]
    const uint16_t* s;      // [N][G]
    half* y;                // [M][ldy]
    const half* res;        // optional residual [M][ldy] added in fp32 before the fp16 store (may alias y), or nullptr
    int M, N, K, ldy;
};

// tile index -> (m block, n block): groups of 16 M-blocks sweep N together, so one wave of CTAs shares X rows and W rows in L2
__device__ __forceinline__ void tile_coord(int tile, int num_m, int num_n, int& m_blk, int& n_blk) {
    constexpr int GM = 16;
    const int per_group = GM * num_n;
    const int group = tile / per_group, r = tile - group * per_group;
    const int gm = (num_m - group * GM < GM) ? num_m - group * GM : GM;
    m_blk = group * GM + r % gm;
    n_blk = r / gm;
}

template <int BN>
__global__ void __launch_bounds__(kThreads, 1) gemm_q4_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const GemmParams p) {
    constexpr int kBN = BN, kStages = Geo<BN>::kStages, kStageBytes = Geo<BN>::kStageBytes, kAccBufs = Geo<BN>::kAccBufs;
    constexpr uint32_t kIdesc = Geo<BN>::kIdesc;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;       // SWIZZLE_128B tiles need 1024-byte alignment
    const uint32_t bars = base + kStages * kStageBytes;
    auto full_a = [&](int s) { return bars + s * 8; };
    auto full_b = [&](int s) { return bars + (kStages + s) * 8; };
    auto empty = [&](int s) { return bars + (2 * kStages + s) * 8; };
    auto tmem_full = [&](int b) { return bars + (3 * kStages + b) * 8; };
    auto tmem_empty = [&](int b) { return bars + (3 * kStages + 2 + b) * 8; };
    const uint32_t tmem_slot = bars + (3 * kStages + 4) * 8;
    auto stage_a = [&](int s) { return base + s * kStageBytes; };
    auto stage_b = [&](int s) { return base + s * kStageBytes + kABytes; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_m = (p.M + kBM - 1) / kBM, num_n = p.N / kBN, num_k = p.K / kBK;
    const int num_tiles = num_m * num_n;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; s++) { mbar_init(full_a(s), 1); mbar_init(full_b(s), kDequantWarps); mbar_init(empty(s), 1); }
        for (int b = 0; b < kAccBufs; b++) { mbar_init(tmem_full(b), 1); mbar_init(tmem_empty(b), kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {      // tensor memory: the whole 512 columns (one CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        // ===== TMA producer (X tiles) =====
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                int m_blk, n_blk;
                tile_coord(tile, num_m, num_n, m_blk, n_blk);
                for (int kb = 0; kb < num_k; kb++, it++) {
                    const int s = it % kStages;
                    mbar_wait(empty(s), ((it / kStages) & 1) ^ 1);
                    mbar_arrive_expect_tx(full_a(s), kABytes);
                    tma_load_2d(stage_a(s), &tmap_x, full_a(s), kb * kBK, m_blk * kBM);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        uint32_t it = 0, tcount = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, tcount++) {
            const int buf = tcount % kAccBufs;
            mbar_wait(tmem_empty(buf), ((tcount / kAccBufs) & 1) ^ 1);       // the epilogue has drained this accumulator pair
            tc_fence_after();
            const uint32_t d0 = tmem_base + buf * 256, d1 = d0 + kBN;
            for (int kb = 0; kb < num_k; kb++, it++) {
                const int s = it % kStages;
                const uint32_t ph = (it / kStages) & 1;
                mbar_wait(full_a(s), ph);
                mbar_wait(full_b(s), ph);
                tc_fence_after();
                if (lane == 0) {
                    const uint64_t da = umma_desc(stage_a(s)), db = umma_desc(stage_b(s));
#pragma unroll
                    for (int k = 0; k < kBK / 16; k++) {            // 32 bytes per k-step inside the 128-byte swizzle row: +2 in the address field
                        const uint32_t acc = (kb | k) ? 1u : 0u;
                        umma_f16(d0, da + 2 * k, db + 2 * k, kIdesc, acc);
                        umma_f16(d1, da + 2 * k + (128 * 128 >> 4), db + 2 * k, kIdesc, acc);
                    }
                    umma_commit(empty(s));                          // frees the stage once these MMAs have read it
                    if (kb == num_k - 1) umma_commit(tmem_full(buf));
                }
                __syncwarp();
            }
        }
    } else if (warp < kFirstEpiWarp) {
        // ===== INT4 -> fp16 dequantisation into the B stage =====
        // BN = 128: thread = (row, 16-byte half of the row's 32 packed bytes); BN = 256: thread = row, both halves
        constexpr int kHalves = BN / 128;                           // 16-byte pieces per thread and k-block
        const int dt = threadIdx.x - 32 * kFirstDequantWarp;       // 0..255
        const int row = (BN == 128) ? dt >> 1 : dt, hw0 = (BN == 128) ? (dt & 1) : 0;
        const int G = (p.K + 127) >> 7, zh = (G + 7) >> 3;
        const uint32_t dst_row = (uint32_t)(row >> 3) * 1024 + (uint32_t)(row & 7) * 128;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            int m_blk, n_blk;
            tile_coord(tile, num_m, num_n, m_blk, n_blk);
            const int n = n_blk * kBN + row;
            const uint32_t* wrow = p.w + (size_t)n * (p.K >> 3) + hw0 * 4;
            const uint16_t* srow = p.s + (size_t)n * G;
            const uint32_t* zrow = p.z + (size_t)n * zh;
            // Packed words, scale and zero word of k-blocks kb and kb + 1 sit in registers; those of kb + 2 are requested before kb is
            // converted (two stage-times of load latency hidden: with one, ncu showed the dequant warps waiting on these loads)
            uint4 wa[kHalves], wb[kHalves];
            uint32_t sa, sb, za, zb;
            auto load_kb = [&](int kb, uint4 (&wd)[kHalves], uint32_t& sd, uint32_t& zd) {
                if (kb < num_k) {
#pragma unroll
                    for (int hh = 0; hh < kHalves; hh++) wd[hh] = lq4::ldg_stream_v4(wrow + (size_t)kb * 8 + hh * 4);
                    const int gg = kb >> 1;
                    sd = lq4::ldg_stream_u16(srow + gg);
                    zd = lq4::ldg_stream_u32(zrow + (gg >> 3));
                }
            };
            load_kb(0, wa, sa, za);
            load_kb(1, wb, sb, zb);
            for (int kb = 0; kb < num_k; kb++, it++) {
                const int s = it % kStages;
                uint4 wcur[kHalves];
#pragma unroll
                for (int hh = 0; hh < kHalves; hh++) { wcur[hh] = wa[hh]; wa[hh] = wb[hh]; }
                const int g = kb >> 1;
                const uint32_t s16 = sa, zq = (za >> ((g & 7) * 4)) & 0xFu;
                sa = sb; za = zb;
                load_kb(kb + 2, wb, sb, zb);
                const uint32_t s2 = s16 | (s16 << 16);
                const uint32_t zlo = (0x6400u | zq) * 0x10001u;            // (1024 + z, 1024 + z)
                const uint32_t zhi = (0x5400u | (zq << 4)) * 0x10001u;     // (  64 + z,   64 + z)
                mbar_wait(empty(s), ((it / kStages) & 1) ^ 1);
                const uint32_t dst = stage_b(s) + dst_row;
#pragma unroll
                for (int hh = 0; hh < kHalves; hh++) {
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const uint32_t w = q == 0 ? wcur[hh].x : q == 1 ? wcur[hh].y : q == 2 ? wcur[hh].z : wcur[hh].w;
                        const uint32_t w8 = w >> 8;
                        uint32_t h[4];
                        h[0] = lq4::and_or(w, 0x000F000Fu, 0x64006400u);       // (1024 + q0, 1024 + q4)
                        h[1] = lq4::and_or(w, 0x00F000F0u, 0x54005400u);       // (  64 + q1,   64 + q5)
                        h[2] = lq4::and_or(w8, 0x000F000Fu, 0x64006400u);      // q2, q6
                        h[3] = lq4::and_or(w8, 0x00F000F0u, 0x54005400u);      // q3, q7
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            __half2 v = *reinterpret_cast<__half2*>(&h[i]);
                            const uint32_t zz = (i & 1) ? zhi : zlo;
                            v = __hsub2(v, *reinterpret_cast<const __half2*>(&zz));          // q - z, exact
                            v = __hmul2(v, *reinterpret_cast<const __half2*>(&s2));          // (q - z) * s, one rounding
                            h[i] = *reinterpret_cast<uint32_t*>(&v);
                        }
                        // k order in memory: (q0,q1) (q2,q3) (q4,q5) (q6,q7)
                        uint4 o;
                        o.x = __byte_perm(h[0], h[1], 0x5410); o.y = __byte_perm(h[2], h[3], 0x5410);
                        o.z = __byte_perm(h[0], h[1], 0x7632); o.w = __byte_perm(h[2], h[3], 0x7632);
                        const int chunk = (hw0 + hh) * 4 + q;                              // 16-byte chunk of the 128-byte row (8 k each)
                        lq4::sts_v4_u32(dst + (uint32_t)((chunk ^ (row & 7)) << 4), o);
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic-proxy writes -> visible to the MMA (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(full_b(s));
            }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> fp16 global =====
        const int q4 = warp & 3;                                   // the TMEM lane quarter this warp may read
        uint32_t tcount = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, tcount++) {
            int m_blk, n_blk;
            tile_coord(tile, num_m, num_n, m_blk, n_blk);
            const int buf = tcount % kAccBufs;
            mbar_wait(tmem_full(buf), (tcount / kAccBufs) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int half_m = 0; half_m < 2; half_m++) {
                const int m = m_blk * kBM + half_m * 128 + q4 * 32 + lane;
                const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + buf * 256 + half_m * kBN;
#pragma unroll 1
                for (int c0 = 0; c0 < kBN; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(taddr + c0, v);
                    tmem_ld_wait();
                    if (m < p.M) {
                        half* yrow = p.y + (size_t)m * p.ldy + n_blk * kBN + c0;
                        const half* rrow = p.res ? p.res + (size_t)m * p.ldy + n_blk * kBN + c0 : nullptr;
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            uint4 r = make_uint4(0, 0, 0, 0);
                            if (rrow) r = *reinterpret_cast<const uint4*>(rrow + j);
                            const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
                            uint32_t o[4];
#pragma unroll
                            for (int e = 0; e < 4; e++) {
                                float a = __uint_as_float(v[j + 2 * e]), b = __uint_as_float(v[j + 2 * e + 1]);
                                if (rrow) { a += lq4::h2f_bits(rr[e] & 0xFFFFu); b += lq4::h2f_bits(rr[e] >> 16); }
                                o[e] = lq4::f2h_bits(a) | (lq4::f2h_bits(b) << 16);
                            }
                            *reinterpret_cast<uint4*>(yrow + j) = make_uint4(o[0], o[1], o[2], o[3]);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty(buf));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
}

// ------------------------------------------------------------------------------------------------
// Row-wise kernels of the batched forward pass (fp32 arithmetic on fp16 storage, the decode path's formulas; the
// summation orders differ from the reference's, which is why prefill parity is a tolerance, not bit-exactness)
// ------------------------------------------------------------------------------------------------
__global__ void embed_rows_kernel(half* x, const half* __restrict__ table, const int* __restrict__ tokens, int dim) {
    const int row = blockIdx.x;
    const uint4* src = reinterpret_cast<const uint4*>(table + (size_t)tokens[row] * dim);
    uint4* dst = reinterpret_cast<uint4*>(x + (size_t)row * dim);
    for (int i = threadIdx.x; i < dim / 8; i += blockDim.x) dst[i] = src[i];
}

__global__ void __launch_bounds__(256) rmsnorm_rows_kernel(half* o, const half* __restrict__ x, const half* __restrict__ w, int dim) {
    __shared__ float red[8];
    const int row = blockIdx.x;
    const half* xr = x + (size_t)row * dim;
    float ss = 0.0f;
    for (int i = threadIdx.x; i < dim; i += 256) { const float v = __half2float(xr[i]); ss = fmaf(v, v, ss); }
    ss = lq4::warp_tree_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; i++) tot += red[i];
    const float scale = 1.0f / sqrtf(tot / (float)dim + 1e-5f);
    for (int i = threadIdx.x; i < dim; i += 256)
        o[(size_t)row * dim + i] = __float2half_rn(__half2float(xr[i]) * (scale * __half2float(w[i])));
}

// RoPE on q (n_heads) and k (n_kv_heads) rows: pairs (i, i + hs/2), position = row % seq (RoPERotation_kernel, gpu_kernels.h:332-355)
// 8 consecutive pair indices per thread: 16-byte loads and stores of both halves of the head (hs % 16 == 0)
__global__ void rope_rows_kernel(half* q, half* k, const float2* __restrict__ tab, int n_heads, int n_kv_heads, int hs, int seq) {
    const int row = blockIdx.x, pos = row % seq;
    const int half_hs = hs / 2, per = half_hs / 8;
    for (int idx = threadIdx.x; idx < (n_heads + n_kv_heads) * per; idx += blockDim.x) {
        const int h = idx / per, i = (idx - h * per) * 8;
        half* v = (h < n_heads) ? q + (size_t)row * n_heads * hs + h * hs : k + (size_t)row * n_kv_heads * hs + (h - n_heads) * hs;
        uint4 lo = *reinterpret_cast<const uint4*>(v + i), hi = *reinterpret_cast<const uint4*>(v + i + half_hs);
        half* a = reinterpret_cast<half*>(&lo);
        half* b = reinterpret_cast<half*>(&hi);
        const float4* cs4 = reinterpret_cast<const float4*>(tab + (size_t)pos * half_hs + i);
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
            const float4 cs = cs4[e >> 1];                 // (cos, sin) of pairs e and e + 1
            const float a0 = __half2float(a[e]), b0 = __half2float(b[e]), a1 = __half2float(a[e + 1]), b1 = __half2float(b[e + 1]);
            a[e] = __float2half_rn(a0 * cs.x - b0 * cs.y);         b[e] = __float2half_rn(a0 * cs.y + b0 * cs.x);
            a[e + 1] = __float2half_rn(a1 * cs.z - b1 * cs.w);     b[e + 1] = __float2half_rn(a1 * cs.w + b1 * cs.z);
        }
        *reinterpret_cast<uint4*>(v + i) = lo;
        *reinterpret_cast<uint4*>(v + i + half_hs) = hi;
    }
}

// 8 elements per thread (n % 8 == 0: the hidden size is a multiple of 8); the arithmetic of gpu_kernels.h:269-273
__global__ void silu_mul_kernel(half* out, const half* __restrict__ gate, const half* __restrict__ up, size_t n) {
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i < n) {
        uint4 gv = *reinterpret_cast<const uint4*>(gate + i);
        const uint4 uv = *reinterpret_cast<const uint4*>(up + i);
        half* gh = reinterpret_cast<half*>(&gv);
        const half* uh = reinterpret_cast<const half*>(&uv);
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const float g = __half2float(gh[e]);
            gh[e] = __float2half_rn(g * (1.0f / (1.0f + expf(-g))) * __half2float(uh[e]));
        }
        *reinterpret_cast<uint4*>(out + i) = gv;
    }
}

// Causal attention of one (sequence, head, block of 16 query rows) per CTA: scores, softmax and PV in fp32 with the scores and
// the probabilities rounded to fp16 where the decode path rounds them.  A CUDA-core kernel: the tensor-core part of this
// round is the projections (97 % of the prefill FLOPs at seq 2048); a tcgen05 attention is the next step.
constexpr int kPfQ = 16;
__global__ void __launch_bounds__(256) attn_prefill_kernel(half* out, const half* __restrict__ q, const half* __restrict__ k, const half* __restrict__ v,
                                                           int seq, int n_heads, int kv_mul, int hs, float alpha) {
    extern __shared__ float sm[];
    const int q0 = blockIdx.x * kPfQ, h = blockIdx.y, b = blockIdx.z;
    const int kvh = h / kv_mul, kv_dim = (n_heads / kv_mul) * hs, dim = n_heads * hs;
    const int nq = min(kPfQ, seq - q0), nk = q0 + nq;              // keys 0 .. q0+nq-1 matter to this block
    float* qs = sm;                                                // [kPfQ][hs]
    float* sc = qs + kPfQ * hs;                                    // [kPfQ][seq]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t row0 = (size_t)b * seq;
    for (int i = tid; i < nq * hs; i += 256) qs[i] = __half2float(q[(row0 + q0 + i / hs) * dim + h * hs + i % hs]);
    __syncthreads();
    // scores: one warp per key, lanes over the head dimension, all query rows of the block at once
    for (int t = warp; t < nk; t += 8) {
        const half* kr = k + (row0 + t) * kv_dim + kvh * hs;
        float kx[8];
        const int per = hs / 32;
        for (int e = 0; e < per; e++) kx[e] = __half2float(kr[e * 32 + lane]);
        for (int r = 0; r < nq; r++) {
            float s = 0.0f;
            for (int e = 0; e < per; e++) s = fmaf(kx[e], qs[r * hs + e * 32 + lane], s);
            s = lq4::warp_tree_sum(s);
            if (lane == 0) sc[r * seq + t] = (t <= q0 + r) ? __half2float(__float2half_rn(s * alpha)) : -INFINITY;
        }
    }
    __syncthreads();
    // softmax per query row (two rows per warp)
    for (int r = warp; r < nq; r += 8) {
        const int n = q0 + r + 1;
        float mx = (n < 1024) ? 0.0f : -INFINITY;                  // the decode path's quirk (gpu_kernels.h:374)
        for (int t = lane; t < n; t += 32) mx = fmaxf(mx, sc[r * seq + t]);
        mx = lq4::warp_max(mx);
        float sum = 0.0f;
        for (int t = lane; t < n; t += 32) { const float e = expf(sc[r * seq + t] - mx); sc[r * seq + t] = e; sum += e; }
        sum = lq4::warp_tree_sum(sum);
        for (int t = lane; t < n; t += 32) sc[r * seq + t] = __half2float(__float2half_rn(sc[r * seq + t] / sum));
    }
    __syncthreads();
    // PV: thread = (query row, group of output dims)
    const int groups = 256 / kPfQ;                                 // 16 threads per query row
    const int r = tid / groups, gidx = tid % groups, per = hs / groups;
    if (r < nq) {
        float acc[16];
        for (int e = 0; e < per; e++) acc[e] = 0.0f;
        const int n = q0 + r + 1;
        for (int t = 0; t < n; t++) {
            const float pt = sc[r * seq + t];
            const half* vr = v + (row0 + t) * kv_dim + kvh * hs + gidx * per;
            for (int e = 0; e < per; e++) acc[e] = fmaf(__half2float(vr[e]), pt, acc[e]);
        }
        half* o = out + (row0 + q0 + r) * dim + h * hs + gidx * per;
        for (int e = 0; e < per; e++) o[e] = __float2half_rn(acc[e]);
    }
}

// Causal attention on the legacy tensor-core path (mma.sync m16n8k16, fp16 in / fp32 accumulate): one CTA of 4 warps per
// (sequence, head, 64 query rows), 16 rows per warp; K and V tiles of 64 keys staged in shared memory by cp.async, online
// softmax in fp32 (scores rounded to fp16 where the decode path rounds them), P.V with V fragments by ldmatrix.trans.
// The projections are the tcgen05 path of this round; a tcgen05 attention (S and O in tensor memory) is the next step.
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}
constexpr int kFaQ = 64, kFaK = 64;
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
template <int HS>
constexpr int fa_smem_bytes() { return 4 * kFaK * (HS + 8) * 2; }      // K and V tiles, two buffers each
// Query tiles are taken heaviest first (the last rows of a sequence see the most keys); K/V tiles are double-buffered (the
// copies of tile i+1 travel while tile i is multiplied); K fragments come from ldmatrix.x4, V fragments from ldmatrix.x4.trans;
// the softmax runs in the base-2 domain (ex2.approx; scores are rounded to fp16 first, like the decode path); only the tile on
// the diagonal is masked.
template <int HS>
__global__ void __maxnreg__(168) attn_prefill_mma_kernel(half* out, const half* __restrict__ q, const half* __restrict__ k, const half* __restrict__ v,
                                                               int seq, int n_heads, int kv_mul, float alpha) {
    constexpr int LD = HS + 8;                                     // padded row (halfs): conflict-free fragment loads
    extern __shared__ __align__(16) uint8_t fa_smem[];
    half* Ks = reinterpret_cast<half*>(fa_smem);                   // [2][kFaK][LD]
    half* Vs = Ks + 2 * kFaK * LD;                                 // [2][kFaK][LD]
    const int q0 = ((int)gridDim.x - 1 - (int)blockIdx.x) * kFaQ, h = blockIdx.y, b = blockIdx.z;
    const int kvh = h / kv_mul, kv_dim = (n_heads / kv_mul) * HS, dim = n_heads * HS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const size_t row0 = (size_t)b * seq;
    const int r_lo = q0 + warp * 16 + g, r_hi = r_lo + 8;          // this thread's two query rows
    const half* kbase = k + row0 * kv_dim + kvh * HS;
    const half* vbase = v + row0 * kv_dim + kvh * HS;
    auto load_tile = [&](int buf, int k0) {
        half* kd = Ks + buf * kFaK * LD;
        half* vd = Vs + buf * kFaK * LD;
        for (int idx = tid; idx < kFaK * (HS / 8); idx += 128) {
            const int r = idx / (HS / 8), c = idx % (HS / 8);
            const size_t key = (size_t)min(k0 + r, seq - 1);
            lq4::cp_async16(smem_u32(kd + r * LD + c * 8), kbase + key * kv_dim + c * 8);
            lq4::cp_async16(smem_u32(vd + r * LD + c * 8), vbase + key * kv_dim + c * 8);
        }
        lq4::cp_async_commit();
    };
    load_tile(0, 0);
    // Q fragments (A operand), straight from global memory
    uint32_t qa[HS / 16][4];
    {
        const half* qlo = q + (row0 + min(r_lo, seq - 1)) * dim + h * HS;
        const half* qhi = q + (row0 + min(r_hi, seq - 1)) * dim + h * HS;
#pragma unroll
        for (int ks = 0; ks < HS / 16; ks++) {
            qa[ks][0] = *reinterpret_cast<const uint32_t*>(qlo + ks * 16 + 2 * t);
            qa[ks][1] = *reinterpret_cast<const uint32_t*>(qhi + ks * 16 + 2 * t);
            qa[ks][2] = *reinterpret_cast<const uint32_t*>(qlo + ks * 16 + 8 + 2 * t);
            qa[ks][3] = *reinterpret_cast<const uint32_t*>(qhi + ks * 16 + 8 + 2 * t);
        }
    }
    float o[HS / 8][4];
#pragma unroll
    for (int d = 0; d < HS / 8; d++) o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.0f;
    float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.0f, l_hi = 0.0f;      // running maxima in the base-2 domain
    const int last_key = min(q0 + kFaQ, seq) - 1;
    constexpr float kLog2e = 1.4426950408889634f;
    // ldmatrix lane addressing: matrix mi = lane / 8, row lane % 8
    const int lm_row = (lane & 7) + ((lane >> 4) << 3), lm_col = ((lane >> 3) & 1) * 8;      // K: two key blocks x two k halves
    int it = 0;
    for (int k0 = 0; k0 <= last_key; k0 += kFaK, it++) {
        const int buf = it & 1;
        const bool more = k0 + kFaK <= last_key;
        if (more) load_tile(buf ^ 1, k0 + kFaK);
        if (more) lq4::cp_async_wait<1>(); else lq4::cp_async_wait<0>();
        __syncthreads();
        const half* kt = Ks + buf * kFaK * LD;
        const half* vt = Vs + buf * kFaK * LD;
        // ---- S = Q K^T (16 rows x 64 keys per warp) ----
        float sacc[kFaK / 8][4];
#pragma unroll
        for (int n = 0; n < kFaK / 8; n++) sacc[n][0] = sacc[n][1] = sacc[n][2] = sacc[n][3] = 0.0f;
#pragma unroll
        for (int ks = 0; ks < HS / 16; ks++) {
#pragma unroll
            for (int n = 0; n < kFaK / 8; n += 2) {
                uint32_t b0, b1, b2, b3;
                const uint32_t addr = smem_u32(kt + (n * 8 + lm_row) * LD + ks * 16 + lm_col);
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3) : "r"(addr));
                mma16816(sacc[n], qa[ks], b0, b1);
                mma16816(sacc[n + 1], qa[ks], b2, b3);
            }
        }
        // ---- scale, round like the decode path, causal mask (diagonal tile only), online softmax in base 2 ----
        const bool diag = k0 + kFaK > q0;
        float mx_lo = m_lo, mx_hi = m_hi;
#pragma unroll
        for (int n = 0; n < kFaK / 8; n++) {
#pragma unroll
            for (int e = 0; e < 4; e++) {
                float sv = __half2float(__float2half_rn(sacc[n][e] * alpha)) * kLog2e;
                if (diag) {
                    const int key = k0 + n * 8 + 2 * t + (e & 1), row = (e < 2) ? r_lo : r_hi;
                    if (key > row || key >= seq) sv = -INFINITY;
                }
                sacc[n][e] = sv;
                if (e < 2) mx_lo = fmaxf(mx_lo, sv); else mx_hi = fmaxf(mx_hi, sv);
            }
        }
        mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1)); mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
        mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1)); mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
        // every row of the block sees key 0 in the first tile, so the running maxima are finite from the first tile on
        const float c_lo = ex2_approx(m_lo - mx_lo), c_hi = ex2_approx(m_hi - mx_hi);
        m_lo = mx_lo; m_hi = mx_hi;
        float s_lo = 0.0f, s_hi = 0.0f;
        uint32_t pa[kFaK / 16][4];
#pragma unroll
        for (int n = 0; n < kFaK / 8; n++) {
            const float p0 = ex2_approx(sacc[n][0] - m_lo), p1 = ex2_approx(sacc[n][1] - m_lo), p2 = ex2_approx(sacc[n][2] - m_hi), p3 = ex2_approx(sacc[n][3] - m_hi);
            s_lo += p0 + p1; s_hi += p2 + p3;
            pa[n >> 1][(n & 1) * 2] = pack_h2(p0, p1);
            pa[n >> 1][(n & 1) * 2 + 1] = pack_h2(p2, p3);
        }
        s_lo += __shfl_xor_sync(0xffffffffu, s_lo, 1); s_lo += __shfl_xor_sync(0xffffffffu, s_lo, 2);
        s_hi += __shfl_xor_sync(0xffffffffu, s_hi, 1); s_hi += __shfl_xor_sync(0xffffffffu, s_hi, 2);
        l_lo = l_lo * c_lo + s_lo; l_hi = l_hi * c_hi + s_hi;
#pragma unroll
        for (int d = 0; d < HS / 8; d++) { o[d][0] *= c_lo; o[d][1] *= c_lo; o[d][2] *= c_hi; o[d][3] *= c_hi; }
        // ---- O += P V: V fragments for two 8-column blocks per ldmatrix.x4.trans ----
#pragma unroll
        for (int kk = 0; kk < kFaK / 16; kk++) {
#pragma unroll
            for (int d = 0; d < HS / 8; d += 2) {
                uint32_t b0, b1, b2, b3;
                const uint32_t addr = smem_u32(vt + (kk * 16 + (lane & 15)) * LD + (d + (lane >> 4)) * 8);
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3) : "r"(addr));
                mma16816(o[d], pa[kk], b0, b1);
                mma16816(o[d + 1], pa[kk], b2, b3);
            }
        }
        __syncthreads();                                           // this buffer is refilled by the next iteration's prefetch
    }
    const float i_lo = 1.0f / l_lo, i_hi = 1.0f / l_hi;
#pragma unroll
    for (int d = 0; d < HS / 8; d++) {
        if (r_lo < seq) *reinterpret_cast<uint32_t*>(out + (row0 + r_lo) * dim + h * HS + d * 8 + 2 * t) = pack_h2(o[d][0] * i_lo, o[d][1] * i_lo);
        if (r_hi < seq) *reinterpret_cast<uint32_t*>(out + (row0 + r_hi) * dim + h * HS + d * 8 + 2 * t) = pack_h2(o[d][2] * i_hi, o[d][3] * i_hi);
    }
}

// rows of one sequence -> its KV-cache layer (so that decode can continue after the prefill)
__global__ void kv_store_kernel(half* kc, half* vc, const half* __restrict__ k, const half* __restrict__ v, int kv_dim) {
    const int pos = blockIdx.x;
    for (int i = threadIdx.x; i < kv_dim / 8; i += blockDim.x) {
        reinterpret_cast<uint4*>(kc + (size_t)pos * kv_dim)[i] = reinterpret_cast<const uint4*>(k + (size_t)pos * kv_dim)[i];
        reinterpret_cast<uint4*>(vc + (size_t)pos * kv_dim)[i] = reinterpret_cast<const uint4*>(v + (size_t)pos * kv_dim)[i];
    }
}

}  // namespace lq4pf

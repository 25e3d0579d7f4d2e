// kernels_sm100.cuh -- hand-written sm_100a device code for the batch-1 decode hot path.
//
// Every kernel reproduces the fp32 summation DAG of the reference kernel it replaces
// (ankan-ban/llama_cu_awq gpu_kernels.h, cited per kernel) so that fp16 results are bit-identical,
// while changing everything that is not arithmetic: which thread owns which partial sum, how many
// loads are in flight, how the INT4 nibbles are turned into fp32, and how many launches a layer takes.
//
// INT4 dequantisation (the issue-rate bound of this path on B200):
//   reference per weight:  shift/and -> I2F -> FSUB(z) -> FMUL(scale) -> FFMA        (5-6 issue slots)
//   here per weight:       1/2 LOP3  -> FHFMA (fma.rn.f32.f16)     -> 1/2 FFMA2      (2.1 issue slots)
//   * one LOP3 `(w & 0x000F000F) | 0x64006400` turns two nibbles into the fp16 pair (1024+q, 1024+q')
//     (mask 0x00F000F0 | 0x54005400 gives (64+q, 64+q') for the odd nibbles: no shift needed);
//   * sm_100's mixed-precision FMA computes (1024+q)*s - (1024+z)*s in ONE rounding. Both products are
//     exact (11-bit x 11-bit significands), and the result (q-z)*s has <= 15 significant bits, so it is
//     exactly the fp32 value the reference forms with `(float(q) - float(z)) * scale`;
//   * two output columns that share x[k] are accumulated by one packed `fma.rn.f32x2` (FFMA2), each half
//     being the same single-rounding fmaf the reference issues.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace lq4 {

// ------------------------------------------------------------------------------------------------
// PTX helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg_stream_v4(const void* p) {
    uint4 r;
    asm("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
        : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ldg_stream_u32(const void* p) {
    uint32_t r;
    asm("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ldg_stream_u16(const void* p) {
    uint16_t r;
    asm("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(r) : "l"(p));
    return (uint32_t)r;
}
// (a & mask) | magic in one LOP3
__device__ __forceinline__ uint32_t and_or(uint32_t a, uint32_t mask, uint32_t magic) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(mask), "r"(magic));
    return d;
}
// d = a.lo16(f16) * s.lo16(f16) + c(f32), one rounding (SASS: FHFMA)
__device__ __forceinline__ float fhfma_lo(uint32_t a, uint32_t s, float c) {
    float d;
    asm("{ .reg .b16 al, ah, sl, sh;\n\t"
        "mov.b32 {al, ah}, %1;\n\t"
        "mov.b32 {sl, sh}, %2;\n\t"
        "fma.rn.f32.f16 %0, al, sl, %3; }"
        : "=f"(d) : "r"(a), "r"(s), "f"(c));
    return d;
}
__device__ __forceinline__ float fhfma_hi(uint32_t a, uint32_t s, float c) {
    float d;
    asm("{ .reg .b16 al, ah, sl, sh;\n\t"
        "mov.b32 {al, ah}, %1;\n\t"
        "mov.b32 {sl, sh}, %2;\n\t"
        "fma.rn.f32.f16 %0, ah, sl, %3; }"
        : "=f"(d) : "r"(a), "r"(s), "f"(c));
    return d;
}
// a.lo*b.lo + c and a.hi*b.hi + c (both halves fp16)
__device__ __forceinline__ float fhfma_ll(uint32_t a, uint32_t b, float c) { return fhfma_lo(a, b, c); }
__device__ __forceinline__ float fhfma_hh(uint32_t a, uint32_t b, float c) {
    float d;
    asm("{ .reg .b16 al, ah, bl, bh;\n\t"
        "mov.b32 {al, ah}, %1;\n\t"
        "mov.b32 {bl, bh}, %2;\n\t"
        "fma.rn.f32.f16 %0, ah, bh, %3; }"
        : "=f"(d) : "r"(a), "r"(b), "f"(c));
    return d;
}
// (a0, a1) = (w0*x + a0, w1*x + a1): two independent single-rounding fmaf in one FFMA2
__device__ __forceinline__ void ffma2(float& a0, float& a1, float w0, float w1, float x) {
    asm("{ .reg .b64 rw, rx, ra;\n\t"
        "mov.b64 rw, {%2, %3};\n\t"
        "mov.b64 rx, {%4, %4};\n\t"
        "mov.b64 ra, {%0, %1};\n\t"
        "fma.rn.f32x2 ra, rw, rx, ra;\n\t"
        "mov.b64 {%0, %1}, ra; }"
        : "+f"(a0), "+f"(a1) : "f"(w0), "f"(w1), "f"(x));
}
// programmatic dependent launch: no-ops when the launch carries no programmatic edge
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float h2f_bits(uint32_t h) { return __half2float(__ushort_as_half((unsigned short)h)); }
__device__ __forceinline__ uint32_t f2h_bits(float f) { return (uint32_t)__half_as_ushort(__float2half_rn(f)); }

// ------------------------------------------------------------------------------------------------
// Reductions with the reference's association
// ------------------------------------------------------------------------------------------------
// cub::WarpReduce<float>::Sum = shfl.down 1,2,4,8,16 (CUB 2.8.2 warp_reduce_shfl.cuh:225-243,550-555).
// Lane 0's value there is the pairwise tree ((P0+P1)+(P2+P3))+...; an xor butterfly builds the same
// tree in every lane (fp add is commutative), so all lanes end with lane-0-of-cub's bits.
__device__ __forceinline__ float warp_tree_sum(float v) {
    v = v + __shfl_xor_sync(0xffffffffu, v, 1);
    v = v + __shfl_xor_sync(0xffffffffu, v, 2);
    v = v + __shfl_xor_sync(0xffffffffu, v, 4);
    v = v + __shfl_xor_sync(0xffffffffu, v, 8);
    v = v + __shfl_xor_sync(0xffffffffu, v, 16);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 8));
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 16));
    return v;
}
// Four per-lane partial sums (one per output column) -> four warp totals with the cub tree, in 6
// shuffles instead of 20.  On return lane L holds the total of column  2*(L&1) + ((L>>1)&1):
// lanes 0,1,2,3 hold columns 0,2,1,3.
__device__ __forceinline__ float warp_tree_sum4(float a0, float a1, float a2, float a3, int lane) {
    const bool b0 = lane & 1, b1 = lane & 2;
    float keep0 = b0 ? a2 : a0, keep1 = b0 ? a3 : a1;
    float send0 = b0 ? a0 : a2, send1 = b0 ? a1 : a3;
    float r0 = keep0 + __shfl_xor_sync(0xffffffffu, send0, 1);
    float r1 = keep1 + __shfl_xor_sync(0xffffffffu, send1, 1);
    float keep = b1 ? r1 : r0, send = b1 ? r0 : r1;
    float r = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    r = r + __shfl_xor_sync(0xffffffffu, r, 4);
    r = r + __shfl_xor_sync(0xffffffffu, r, 8);
    r = r + __shfl_xor_sync(0xffffffffu, r, 16);
    return r;
}

// ------------------------------------------------------------------------------------------------
// Activation staging: x (fp16, global) -> fp32 in shared memory, optionally through RMSNorm
// ------------------------------------------------------------------------------------------------
// Layout: element k = t*1024 + L*32 + j*4 + e  (t trip, L lane, j 16-byte chunk, e element) is stored
// at t*1024 + j*128 + L*4 + e, so that lane L's float4 reads of chunk j are conflict-free.
__device__ __forceinline__ int xs_index(int k) {
    return (k & ~1023) | ((k & 28) << 5) | ((k >> 3) & 124) | (k & 3);
}

// RMSNorm restated from rmsnorm_kernel (gpu_kernels.h:72-105) for a block of NT threads standing in for
// the reference's 1024: virtual thread vt accumulates x[vt + 1024*i]^2 (i ascending, FFMA), virtual
// warps (32 consecutive vt) are tree-summed, and the 32 warp aggregates are added in order 0..31.
// Returns the scale  1/sqrt(mean+eps)  in every thread.  `red` is 32 floats of shared memory.
template <int NT>
__device__ __forceinline__ float block_rms_scale(const half* __restrict__ x, int size, float* red) {
    static_assert(1024 % NT == 0 && NT % 32 == 0, "NT must divide 1024");
    constexpr int VPT = 1024 / NT;  // virtual threads per thread
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ept = (size - 1) / 1024 + 1;
#pragma unroll
    for (int v = 0; v < VPT; v++) {
        const int vt = tid + v * NT;
        float ss = 0.0f;
        for (int i = 0; i < ept; i++) {
            const int idx = vt + i * 1024;
            if (idx < size) {
                const float val = __half2float(x[idx]);
                ss = __fmaf_rn(val, val, ss);
            }
        }
        ss = warp_tree_sum(ss);
        if (lane == 0) red[warp + v * (NT / 32)] = ss;  // virtual warp id = vt/32
    }
    __syncthreads();
    float tot = red[0];
#pragma unroll
    for (int w = 1; w < 32; w++) tot = tot + red[w];
    tot = __fdiv_rn(tot, (float)size);
    tot = tot + 1e-5f;
    tot = __fdiv_rn(1.0f, __fsqrt_rn(tot));
    return tot;
}

// Stage x into xs (fp32, permuted).  With norm_w: xs = float(half(x * (scale * w))), the fp16 rounding
// being the reference's store of xb (gpu_kernels.h:100-102).  Ends with __syncthreads().
template <int NT>
__device__ __forceinline__ void stage_x_f32(float* xs, float* red, const half* __restrict__ x,
                                            const half* __restrict__ norm_w, int K) {
    if (norm_w != nullptr) {
        const float scale = block_rms_scale<NT>(x, K, red);
        for (int k = threadIdx.x; k < K; k += NT) {
            float val = __half2float(x[k]);
            val = __fmul_rn(val, __fmul_rn(scale, __half2float(norm_w[k])));
            xs[xs_index(k)] = __half2float(__float2half_rn(val));
        }
    } else {
        for (int k = threadIdx.x; k < K; k += NT) xs[xs_index(k)] = __half2float(x[k]);
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// INT4 GEMV core
// ------------------------------------------------------------------------------------------------
struct QW {  // device view of a reference QWeight (common.h:20-24)
    const uint32_t* w;
    const uint32_t* z;
    const uint16_t* s;
};

// One warp-task = 4 output columns:  cols {c0, c0+1} of matrix A and {c0+d, c0+d+1} of matrix B.
//   plain GEMV : A == B, d = 2            (4 consecutive columns)
//   RoPE pair  : A == B, d = head_size/2  (the two rotation partners of columns c0, c0+1)
//   gate/up    : A = gate, B = up, d = 0
struct Task4 {
    QW A, B;
    int c0, d;
};

struct Trip4 {       // one trip (1024 k) of a warp-task held in registers: what the lane needs
    uint4 w[4];      // 32 nibbles per column: k = trip*1024 + lane*32 + [0,32)
    uint32_t z[4];   // packed zero points of groups trip*8 .. trip*8+7
    uint32_t s[4];   // fp16 scale of group trip*8 + lane/4
};

__device__ __forceinline__ void load_trip4(Trip4& r, const Task4& t, int trip, int lane, int pwh, int zh, int G) {
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const QW& m = (c < 2) ? t.A : t.B;
        const size_t col = (size_t)(t.c0 + (c & 1) + ((c < 2) ? 0 : t.d));
        r.w[c] = ldg_stream_v4(m.w + col * pwh + trip * 128 + lane * 4);
        r.z[c] = ldg_stream_u32(m.z + col * zh + trip);
        r.s[c] = ldg_stream_u16(m.s + col * G + trip * 8 + (lane >> 2));
    }
}

// 8 nibbles of `w` -> 8 exact fp32 weights (q - z) * s, in k order
__device__ __forceinline__ void dequant8(float* d, uint32_t w, uint32_t s, float nlo, float nhi) {
    const uint32_t p04 = and_or(w, 0x000F000Fu, 0x64006400u);   // (1024+q0, 1024+q4)
    const uint32_t p15 = and_or(w, 0x00F000F0u, 0x54005400u);   // (  64+q1,   64+q5)
    const uint32_t w8 = w >> 8;
    const uint32_t p26 = and_or(w8, 0x000F000Fu, 0x64006400u);
    const uint32_t p37 = and_or(w8, 0x00F000F0u, 0x54005400u);
    d[0] = fhfma_lo(p04, s, nlo);
    d[1] = fhfma_lo(p15, s, nhi);
    d[2] = fhfma_lo(p26, s, nlo);
    d[3] = fhfma_lo(p37, s, nhi);
    d[4] = fhfma_hi(p04, s, nlo);
    d[5] = fhfma_hi(p15, s, nhi);
    d[6] = fhfma_hi(p26, s, nlo);
    d[7] = fhfma_hi(p37, s, nhi);
}

// acc[c] = fma(w, x[k], acc[c]) over the trip's 32 k in ascending order (gpu_kernels.h:188-201)
__device__ __forceinline__ void compute_trip4(float* acc, const Trip4& r, const float* xs_trip, int lane) {
    uint32_t s[4];
    float nlo[4], nhi[4];
    const int zshift = (lane >> 2) * 4;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const float zf = (float)((r.z[c] >> zshift) & 0xFu);
        const float sf = h2f_bits(r.s[c]);
        s[c] = r.s[c];
        nlo[c] = __fmul_rn(__fadd_rn(zf, 1024.0f), -sf);   // -(1024+z)*s, exact
        nhi[c] = __fmaf_rn(960.0f, sf, nlo[c]);            // -(64+z)*s, exact
    }
    const float4* xs4 = reinterpret_cast<const float4*>(xs_trip);
#pragma unroll
    for (int qi = 0; qi < 4; qi++) {
        const float4 xa = xs4[(2 * qi) * 32 + lane];
        const float4 xb = xs4[(2 * qi + 1) * 32 + lane];
        const float x[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
        for (int p = 0; p < 2; p++) {
            const int a = 2 * p, b = 2 * p + 1;
            const uint32_t wa = (qi == 0) ? r.w[a].x : (qi == 1) ? r.w[a].y : (qi == 2) ? r.w[a].z : r.w[a].w;
            const uint32_t wb = (qi == 0) ? r.w[b].x : (qi == 1) ? r.w[b].y : (qi == 2) ? r.w[b].z : r.w[b].w;
            float da[8], db[8];
            dequant8(da, wa, s[a], nlo[a], nhi[a]);
            dequant8(db, wb, s[b], nlo[b], nhi[b]);
#pragma unroll
            for (int i = 0; i < 8; i++) ffma2(acc[a], acc[b], da[i], db[i], x[i]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Fused INT4 GEMV kernel
// ------------------------------------------------------------------------------------------------
enum GemvKind { GEMV_PLAIN = 0, GEMV_QKV = 1, GEMV_FFN = 2 };

struct GemvParams {
    // activation in
    const half* x;           // [K] fp16 (ignored when emb_table != nullptr)
    const half* norm_w;      // RMSNorm weight to fuse (gpu_kernels.h:72-105) or nullptr
    const half* emb_table;   // layer 0: x = emb_table[tokens[*pPos]] (copy_embedding_kernel, :61-69)
    const int* tokens;
    half* x_copy;            // where block 0 copies the gathered embedding row (the residual stream)
    int K;
    // matrices: PLAIN uses m[0]; QKV uses m[0..2] = q,k,v; FFN uses m[0]=gate, m[1]=up
    QW m[3];
    int n[3];                // output columns per matrix
    half* out[3];
    int accum;               // PLAIN: out += (residual, gpu_kernels.h:229-230)
    int loff;                // cache offset (elements) added to out[1], out[2] (QKV) or out[0] (PLAIN) with pPos
    const int* pPos;         // device position; rows land at loff + pos*n
    // fused RoPE (QKV only): rope_tab[pos*(head_size/2) + i] = (cos, sin); nullptr = no rotation
    const float2* rope_tab;
    int head_size;
};

constexpr int kGemvThreads = 256;
constexpr int kGemvWarps = kGemvThreads / 32;

template <int KIND>
__device__ __forceinline__ int gemv_num_tasks(const GemvParams& p) {
    if (KIND == GEMV_QKV) return (p.n[0] + p.n[1] + p.n[2]) / 4;
    return p.n[0] / 4;   // PLAIN: 4 columns; FFN: 2 gate + 2 up columns => n/2... handled below
}

template <int KIND>
__device__ __forceinline__ void gemv_make_task(const GemvParams& p, int task, Task4& t, int& mat) {
    if (KIND == GEMV_PLAIN) {
        mat = 0;
        t.A = p.m[0]; t.B = p.m[0]; t.c0 = task * 4; t.d = 2;
    } else if (KIND == GEMV_FFN) {
        mat = 0;
        t.A = p.m[0]; t.B = p.m[1]; t.c0 = task * 2; t.d = 0;
    } else {
        const int t0 = p.n[0] / 4, t1 = t0 + p.n[1] / 4;
        mat = (task < t0) ? 0 : (task < t1) ? 1 : 2;
        const int local = task - ((mat == 0) ? 0 : (mat == 1) ? t0 : t1);
        t.A = p.m[mat]; t.B = p.m[mat];
        if (p.rope_tab != nullptr && mat < 2) {
            // head h, pair index i (even): columns {b+i, b+i+1, b+i+hs/2, b+i+hs/2+1}
            const int per_head = p.head_size / 4;
            const int h = local / per_head, i = (local - h * per_head) * 2;
            t.c0 = h * p.head_size + i; t.d = p.head_size / 2;
        } else {
            t.c0 = local * 4; t.d = 2;
        }
    }
}

template <int KIND>
__global__ void __launch_bounds__(kGemvThreads, 2) gemv_q4_kernel(const GemvParams p) {
    extern __shared__ __align__(16) float smem_f[];
    float* xs = smem_f;                                   // ceil(K/1024)*1024 floats
    float* red = smem_f + ((p.K + 1023) & ~1023);         // 32 floats
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int K = p.K;
    const int pwh = K >> 3, G = (K + 127) >> 7, zh = (G + 7) >> 3;   // K % 32 == 0
    const int ntrips = (pwh + 127) >> 7;
    const int ntasks = (KIND == GEMV_FFN) ? p.n[0] / 2 : gemv_num_tasks<KIND>(p);
    const int gwarp = blockIdx.x * kGemvWarps + warp, gstride = gridDim.x * kGemvWarps;

    griddep_launch();   // let the next kernel in the stream start its own weight prefetch

    // ---- stream position = (task, trip); ring of three trips in registers, two in flight ----
    Trip4 r0, r1, r2;
    Task4 tk;             // task of the position being computed
    int mat = 0;
    int task = gwarp;
    // load cursor
    int ltask = gwarp, ltrip = 0;
    Task4 ltk; int lmat = 0;
    if (ltask < ntasks) gemv_make_task<KIND>(p, ltask, ltk, lmat);
    auto lane_active = [&](int trip) { return trip * 128 + lane * 4 < pwh; };
    auto issue = [&](Trip4& r) {
        if (ltask < ntasks) {
            if (lane_active(ltrip)) load_trip4(r, ltk, ltrip, lane, pwh, zh, G);
            if (++ltrip == ntrips) {
                ltrip = 0; ltask += gstride;
                if (ltask < ntasks) gemv_make_task<KIND>(p, ltask, ltk, lmat);
            }
        }
    };
    issue(r0);
    issue(r1);

    // ---- activations (depend on the previous kernel) ----
    griddep_wait();
    const half* xin = p.x;
    int pos = 0;
    if (p.pPos != nullptr) pos = *p.pPos;
    if (p.emb_table != nullptr) {
        const int token = p.tokens[pos];
        xin = p.emb_table + (size_t)token * K;
        if (blockIdx.x == 0)
            for (int k = threadIdx.x; k < K; k += kGemvThreads) p.x_copy[k] = xin[k];
    }
    stage_x_f32<kGemvThreads>(xs, red, xin, p.norm_w, K);

    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    int trip = 0;
    if (task < ntasks) gemv_make_task<KIND>(p, task, tk, mat);

    auto step = [&](Trip4& cur, Trip4& nxt2) {
        issue(nxt2);
        if (lane_active(trip)) compute_trip4(acc, cur, xs + trip * 1024, lane);
        if (++trip == ntrips) {
            // ---- epilogue: cub-order reduction, then the reference's store ----
            const float tot = warp_tree_sum4(acc[0], acc[1], acc[2], acc[3], lane);
            const int cidx = 2 * (lane & 1) + ((lane >> 1) & 1);      // column slot held by this lane
            const int col = tk.c0 + (cidx & 1) + ((cidx < 2) ? 0 : tk.d);
            if (KIND == GEMV_FFN) {
                // lanes 0,2 hold gate(c0), gate(c0+1); lanes 1,3 hold up(c0), up(c0+1)
                const float u = __shfl_xor_sync(0xffffffffu, tot, 1);
                if (lane == 0 || lane == 2) {
                    float val = tot;                                   // gpu_kernels.h:269-273
                    val = __fmul_rn(val, __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-val))));
                    val = __fmul_rn(val, u);
                    p.out[0][col] = __float2half_rn(val);
                }
            } else if (KIND == GEMV_QKV) {
                half* dst = p.out[mat];
                if (mat > 0) dst += p.loff + (size_t)pos * p.n[mat];
                if (p.rope_tab != nullptr && mat < 2) {
                    // RoPERotation_kernel (gpu_kernels.h:332-355) on the fp16-rounded GEMV result
                    const float own = __half2float(__float2half_rn(tot));
                    const float oth = __shfl_xor_sync(0xffffffffu, own, 1);
                    if (lane < 4) {
                        const int i = (tk.c0 % p.head_size) + (lane >> 1);
                        const float2 cs = p.rope_tab[(size_t)pos * (p.head_size / 2) + i];
                        // contraction shapes as in the reference build's SASS (see rope_kernel below)
                        float o;
                        if ((lane & 1) == 0) o = __fmaf_rn(own, cs.x, -__fmul_rn(oth, cs.y));   // v0*c - v1*s
                        else if (mat == 0)   o = __fmaf_rn(own, cs.x, __fmul_rn(oth, cs.y));    // q: fma(q1,c,q0*s)
                        else                 o = __fmaf_rn(oth, cs.y, __fmul_rn(own, cs.x));    // k: fma(k0,s,k1*c)
                        dst[col] = __float2half_rn(o);
                    }
                } else if (lane < 4) {
                    dst[col] = __float2half_rn(tot);
                }
            } else {
                if (lane < 4) {
                    half* dst = p.out[0];
                    if (p.loff != -1 && p.pPos != nullptr) dst += p.loff + (size_t)pos * p.n[0];
                    float sum = tot;
                    if (p.accum) sum = sum + __half2float(dst[col]);
                    dst[col] = __float2half_rn(sum);
                }
            }
            acc[0] = acc[1] = acc[2] = acc[3] = 0.0f;
            trip = 0;
            task += gstride;
            if (task < ntasks) gemv_make_task<KIND>(p, task, tk, mat);
        }
    };

    while (task < ntasks) {
        step(r0, r2);
        if (task >= ntasks) break;
        step(r1, r0);
        if (task >= ntasks) break;
        step(r2, r1);
    }
}

// ------------------------------------------------------------------------------------------------
// fp16 classifier GEMV (mat_vec_kernel, gpu_kernels.h:109-139) with the final RMSNorm fused
// ------------------------------------------------------------------------------------------------
struct ClsParams {
    const half* x;        // [n]
    const half* norm_w;   // final RMSNorm weight or nullptr
    const half* w;        // [d][w_row_stride]
    half* out;            // [d]
    half* x_norm_out;     // optional: block 0 writes the normalised x here (reference normalises in place)
    int n, d, w_row_stride;
    float alpha;
};

constexpr int kClsThreads = 256;
constexpr int kClsRows = 4;   // rows per warp

__global__ void __launch_bounds__(kClsThreads, 2) gemv_f16_kernel(const ClsParams p) {
    extern __shared__ __align__(16) float smem_f[];
    const int n = p.n;
    uint32_t* xh = reinterpret_cast<uint32_t*>(smem_f);        // n fp16 (as u32 pairs), natural order
    float* red = smem_f + ((n / 2 + 3) & ~3);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ntrips = ((n - 1) / 32 + 1 - 1) / 8 + 1;         // divUp(divUp(n,32),8), llama2_q4.cu:216-217
    const int ngroups = (p.d + kClsRows - 1) / kClsRows;
    const int gwarp = blockIdx.x * (kClsThreads / 32) + warp, gstride = gridDim.x * (kClsThreads / 32);

    griddep_launch();
    // prefetch the first trip of the first row group (weights do not depend on the previous kernel)
    uint4 wr[2][kClsRows];
    int g = gwarp;
    auto load = [&](uint4* dst, int grp, int trip) {
        const int j = (trip * 32 + lane) * 8;
#pragma unroll
        for (int r = 0; r < kClsRows; r++) {
            const int row = grp * kClsRows + r;
            if (row < p.d && j < n) dst[r] = ldg_stream_v4(p.w + (size_t)row * p.w_row_stride + j);
        }
    };
    if (g < ngroups) load(wr[0], g, 0);

    griddep_wait();
    __half* xh16 = reinterpret_cast<__half*>(xh);
    if (p.norm_w != nullptr) {
        const float scale = block_rms_scale<kClsThreads>(p.x, n, red);
        for (int k = threadIdx.x; k < n; k += kClsThreads) {
            float val = __half2float(p.x[k]);
            val = __fmul_rn(val, __fmul_rn(scale, __half2float(p.norm_w[k])));
            const __half hv = __float2half_rn(val);
            xh16[k] = hv;
            if (p.x_norm_out != nullptr && blockIdx.x == 0) p.x_norm_out[k] = hv;
        }
    } else {
        for (int k = threadIdx.x; k < n; k += kClsThreads) xh16[k] = p.x[k];
    }
    __syncthreads();

    const uint4* xs4 = reinterpret_cast<const uint4*>(xh);
    for (; g < ngroups; g += gstride) {
        float acc[kClsRows];
#pragma unroll
        for (int r = 0; r < kClsRows; r++) acc[r] = 0.0f;
        for (int trip = 0; trip < ntrips; trip++) {
            uint4* cur = wr[trip & 1];
            uint4* nxt = wr[(trip + 1) & 1];
            if (trip + 1 < ntrips) load(nxt, g, trip + 1);
            else if (g + gstride < ngroups) load(nxt, g + gstride, 0);
            const int j = (trip * 32 + lane) * 8;
            if (j < n) {
                const uint4 xv = xs4[trip * 32 + lane];
#pragma unroll
                for (int r = 0; r < kClsRows; r++) {
                    // sum = fma(float(w[j+el]), float(x[j+el]), sum), el ascending (gpu_kernels.h:127-128)
                    float a = acc[r];
                    a = fhfma_ll(cur[r].x, xv.x, a); a = fhfma_hh(cur[r].x, xv.x, a);
                    a = fhfma_ll(cur[r].y, xv.y, a); a = fhfma_hh(cur[r].y, xv.y, a);
                    a = fhfma_ll(cur[r].z, xv.z, a); a = fhfma_hh(cur[r].z, xv.z, a);
                    a = fhfma_ll(cur[r].w, xv.w, a); a = fhfma_hh(cur[r].w, xv.w, a);
                    acc[r] = a;
                }
            }
        }
        // ntrips may be odd/even: the prefetch for the next group landed in wr[ntrips & 1]; realign to wr[0]
        if (ntrips & 1) {
#pragma unroll
            for (int r = 0; r < kClsRows; r++) wr[0][r] = wr[1][r];
        }
        const float tot = warp_tree_sum4(acc[0], acc[1], acc[2], acc[3], lane);
        if (lane < 4) {
            const int cidx = 2 * (lane & 1) + ((lane >> 1) & 1);
            const int row = g * kClsRows + cidx;
            if (row < p.d) p.out[row] = __float2half_rn(__fmul_rn(tot, p.alpha));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// RoPE angles.  One device function is the single source of the (cos, sin) bits: the table builder
// and the stand-alone rotation kernel both call it, restating gpu_kernels.h:338-342.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 rope_cos_sin(int pos, int i, int head_size, float rope_theta) {
    const int head_dim = (i * 2) % head_size;
    const float freq = 1.0f / powf(rope_theta, head_dim / (float)head_size);
    const float val = pos * freq;
    const float fcr = cosf(val);
    const float fci = sinf(val);
    return make_float2(fcr, fci);
}

__global__ void rope_table_kernel(float2* tab, int seq_len, int head_size, float rope_theta) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int half_hs = head_size / 2;
    if (idx >= seq_len * half_hs) return;
    const int pos = idx / half_hs, i = idx - pos * half_hs;
    tab[idx] = rope_cos_sin(pos, i, head_size, rope_theta);
}

// Stand-alone RoPERotation (operator API; gpu_kernels.h:332-355).  The FMA contraction is pinned to what
// nvcc 12.9 -O3 emits for the reference on sm_100a (cuobjdump -sass of oracle/_ref/llama2_q4_ref):
//   q: out0 = fma(q0,c,-(q1*s))  out1 = fma(q1,c,q0*s)      k: out0 = fma(k0,c,-(k1*s))  out1 = fma(k0,s,k1*c)
__global__ void rope_kernel(half* sq, half* sk_base, int num_kv_heads, int head_size, const int* pPos, int loff,
                            float rope_theta) {
    griddep_launch();
    griddep_wait();
    const int pos = *pPos;
    const int h = blockIdx.x, i = threadIdx.x;
    const float2 cs = rope_cos_sin(pos, i, head_size, rope_theta);
    half* q = sq + h * head_size;
    const float q0 = __half2float(q[i]), q1 = __half2float(q[i + head_size / 2]);
    q[i] = __float2half_rn(__fmaf_rn(q0, cs.x, -__fmul_rn(q1, cs.y)));
    q[i + head_size / 2] = __float2half_rn(__fmaf_rn(q1, cs.x, __fmul_rn(q0, cs.y)));
    if (h < num_kv_heads) {
        half* k = sk_base + loff + (size_t)pos * num_kv_heads * head_size + h * head_size;
        const float k0 = __half2float(k[i]), k1 = __half2float(k[i + head_size / 2]);
        k[i] = __float2half_rn(__fmaf_rn(k0, cs.x, -__fmul_rn(k1, cs.y)));
        k[i + head_size / 2] = __float2half_rn(__fmaf_rn(k0, cs.y, __fmul_rn(k1, cs.x)));
    }
}

// ------------------------------------------------------------------------------------------------
// Stand-alone RMSNorm (operator API; gpu_kernels.h:72-105).  One block of 1024 threads.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) rmsnorm_kernel(half* o, const half* x, const half* weight, int size) {
    __shared__ float red[32];
    griddep_launch();
    griddep_wait();
    const float scale = block_rms_scale<1024>(x, size, red);
    for (int k = threadIdx.x; k < size; k += 1024) {
        float val = __half2float(x[k]);
        val = __fmul_rn(val, __fmul_rn(scale, __half2float(weight[k])));
        o[k] = __float2half_rn(val);
    }
}

__global__ void copy_embedding_kernel(half* x, const half* __restrict__ table, int size, const int* tokens,
                                      const int* pPos) {
    const int index = blockIdx.x * blockDim.x + threadIdx.x;
    griddep_launch();
    griddep_wait();
    if (index >= size) return;
    const int token = tokens[*pPos];
    x[index] = table[(size_t)token * size + index];
}

__global__ void convert_fp16_to_fp32_kernel(float* out, const half* in, int elements) {
    const int index = blockIdx.x * blockDim.x + threadIdx.x;
    if (index < elements) out[index] = __half2float(in[index]);
}

// ------------------------------------------------------------------------------------------------
// Fused decode attention: QK^T -> softmax -> PV for one head per block, 1024 threads
// (mat_vec_kernel_simple :142-168, softmax_kernel :357-401, vec_mat_kernel :279-329).
// ------------------------------------------------------------------------------------------------
struct AttnParams {
    half* out;               // [n_heads*head_size]
    const half* q;           // [n_heads*head_size] (already rotated)
    const half* kcache;      // layer base, rows of kv_stride halfs
    const half* vcache;
    half* att_out;           // optional: probabilities [n_heads][pos+1] like the reference's s->att, or nullptr
    int head_size, kv_mul, kv_stride;
    const int* pPos;
    float alpha;             // (float)(1.0 / sqrt((double)head_size)), llama2_q4.cu:273
    int max_seq;             // capacity of the score buffer in shared memory
    int exp16;               // 1: arithmetic of softmax_kernel_no_smem (gpu_kernels.h:403-446), the reference's choice when max_seq_len > 8192
};

constexpr int kAttnThreads = 1024;

__global__ void __launch_bounds__(kAttnThreads, 1) attention_kernel(const AttnParams p) {
    extern __shared__ __align__(16) float smem_f[];
    const int hs = p.head_size;                       // multiple of 32, <= 256
    float* qs = smem_f;                               // hs
    float* red = qs + hs;                             // 32
    float* bc = red + 32;                             // 4 (broadcast slots)
    float* att = bc + 4;                              // max_seq
    float* part = att + ((p.max_seq + 3) & ~3);       // 32 * hs (tx-class partial sums)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int h = blockIdx.x;

    griddep_launch();
    griddep_wait();
    const int pos = *p.pPos;
    const int size = pos + 1;
    const half* kbase = p.kcache + (size_t)(h / p.kv_mul) * hs;
    const half* vbase = p.vcache + (size_t)(h / p.kv_mul) * hs;
    for (int j = tid; j < hs; j += kAttnThreads) qs[j] = __half2float(p.q[h * hs + j]);
    __syncthreads();

    // ---- scores: one warp per t; lane chain over j = 32*i + lane (gpu_kernels.h:154-159) ----
    const int nser = hs / 32;
    for (int t = warp; t < size; t += 32) {
        const half* krow = kbase + (size_t)t * p.kv_stride;
        float sum = 0.0f;
        for (int i = 0; i < nser; i++) {
            const int j = i * 32 + lane;
            sum = __fmaf_rn(__half2float(krow[j]), qs[j], sum);
        }
        sum = warp_tree_sum(sum);
        sum = __fmul_rn(sum, p.alpha);
        if (lane == 0) att[t] = __half2float(__float2half_rn(sum));      // scores round-trip through fp16
    }
    __syncthreads();

    // ---- softmax (gpu_kernels.h:373-400): idle threads seed the max with 0 ----
    float mx = (tid < size) ? att[tid] : 0.0f;
    for (int i = tid + kAttnThreads; i < size; i += kAttnThreads) mx = fmaxf(mx, att[i]);
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int w = 1; w < 32; w++) mx = fmaxf(mx, red[w]);
    __syncthreads();                                   // red is reused below
    float ssum = 0.0f;
    for (int i = tid; i < size; i += kAttnThreads) {
        const float e = expf(__fsub_rn(att[i], mx));
        att[i] = p.exp16 ? __half2float(__float2half_rn(e)) : e;   // no_smem variant: exp() is parked in the fp16 score buffer
        ssum = __fadd_rn(ssum, e);                     // FMUL (expf tail) + FADD in the reference SASS: not fused
    }
    ssum = warp_tree_sum(ssum);
    if (lane == 0) red[warp] = ssum;
    __syncthreads();
    float tot = red[0];
#pragma unroll
    for (int w = 1; w < 32; w++) tot = tot + red[w];
    for (int i = tid; i < size; i += kAttnThreads) {
        const __half pr = __float2half_rn(__fdiv_rn(att[i], tot));
        att[i] = __half2float(pr);
        if (p.att_out != nullptr) p.att_out[(size_t)h * size + i] = pr;
    }
    __syncthreads();

    // ---- PV: warp w owns the reference's lane tx = w (rows t = w, w+32, ...), lanes own outputs ----
    // reference chain for output i, lane tx: sum = fma(V[32e+tx][i], p[32e+tx], sum), e ascending (:311)
    {
        float a[8];
#pragma unroll
        for (int c = 0; c < 8; c++) a[c] = 0.0f;
        const int per_lane = hs / 32;                 // outputs per lane: i = lane*per_lane + c
        for (int t = warp; t < size; t += 32) {
            const half* vrow = vbase + (size_t)t * p.kv_stride + lane * per_lane;
            const float pt = att[t];
#pragma unroll
            for (int c = 0; c < 8; c++)
                if (c < per_lane) a[c] = __fmaf_rn(__half2float(vrow[c]), pt, a[c]);
        }
#pragma unroll
        for (int c = 0; c < 8; c++)
            if (c < per_lane) part[warp * hs + lane * per_lane + c] = a[c];
    }
    __syncthreads();
    for (int i = tid; i < hs; i += kAttnThreads) {
        // cub tree over tx = 0..31 (vec_mat_kernel :323-325)
        float v[32];
#pragma unroll
        for (int w = 0; w < 32; w++) v[w] = part[w * hs + i];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
#pragma unroll
            for (int w = 0; w < 32; w += 2 * o) v[w] = v[w] + v[w + o];
        p.out[h * hs + i] = __float2half_rn(v[0]);
    }
}

// ------------------------------------------------------------------------------------------------
// Greedy sampler (argmax_kernel, gpu_kernels.h:448-493).  Equal maxima: lowest index (the reference
// leaves the winner to a write race, :474-479).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) argmax_kernel(const half* __restrict__ x, int size, int* result,
                                                     volatile int* pPos, int* pPosGpu, int* result_dev,
                                                     bool write_token) {
    __shared__ float smax[32];
    __shared__ int sidx[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    griddep_wait();
    float max_val = (tid < size) ? __half2float(x[tid]) : -INFINITY;
    int max_pos = (tid < size) ? tid : 0x7fffffff;
    for (int i = tid + 1024; i < size; i += 1024) {
        const float v = __half2float(x[i]);
        if (v > max_val) { max_val = v; max_pos = i; }
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, max_val, o);
        const int op = __shfl_xor_sync(0xffffffffu, max_pos, o);
        if (ov > max_val || (ov == max_val && op < max_pos)) { max_val = ov; max_pos = op; }
    }
    if (lane == 0) { smax[warp] = max_val; sidx[warp] = max_pos; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < 32; w++)
            if (smax[w] > max_val || (smax[w] == max_val && sidx[w] < max_pos)) { max_val = smax[w]; max_pos = sidx[w]; }
        int token_pos = *pPos;
        token_pos++;
        if (write_token) {
            result[token_pos] = max_pos;
            if (result_dev != nullptr) result_dev[token_pos] = max_pos;
        }
        *pPos = token_pos;       // unblocks the CPU (pinned host memory)
        *pPosGpu = token_pos;
    }
}

}  // namespace lq4

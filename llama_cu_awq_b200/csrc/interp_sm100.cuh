// interp_sm100.cuh -- the persistent decode kernel for sm_100a ("op interpreter").
//
// One launch walks a table of ops (the whole per-token forward pass of llama2_q4.cu:286-340 plus the
// greedy sampler, or a single op for the operator API) with ONE CTA PER SM:
//
//   * a PRODUCER warp streams every weight byte the CTA will need through a ring of shared-memory slots
//     with 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx).  Weights never depend on
//     activations, so the producer runs ahead across op boundaries and grid barriers: HBM keeps
//     streaming while the consumers wait on a dependency.
//   * CONSUMER warps are arranged as SYSTOLIC PIPELINES along K.  Stage i of a pipeline owns the
//     reference's "trip" i (k in [1024 i, 1024 i + 1024), gpu_kernels.h:176-201) and keeps its slice of
//     the activation vector in REGISTERS for the whole op, so the inner loop reads nothing but packed
//     weights from shared memory.  The fp32 accumulator of a column is handed from stage to stage
//     through shared memory, which keeps every per-lane FMA chain of the reference in its original
//     order: results are bit-identical.
//   * one thread owns TWO reference lanes (2j, 2j+1) of one column, so that one packed FFMA2
//     (fma.rn.f32x2) advances both chains with a natural (x[k], x[k+32]) register pair, and the
//     per-group scale / zero-point preparation is shared by 64 weights.  A half-warp is one column.
//   * ops are separated by a grid-wide barrier (release/acquire counter in global memory) only where
//     the dataflow needs one.
//
// INT4 dequantisation: see kernels_sm100.cuh (LOP3 -> FHFMA -> FFMA2, exact).
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "kernels_sm100.cuh"

namespace lq4 {

// ------------------------------------------------------------------------------------------------
// op table
// ------------------------------------------------------------------------------------------------
enum OpKind {
    OP_GEMV = 0,     // INT4 GEMV(s): 1..3 matrices over the same x (plain, or q|k|v), optional fused RMSNorm/embedding
    OP_FFN = 1,      // gate/up INT4 GEMVs + SiLU (ffn_matvec_silu_kernel, gpu_kernels.h:256-275)
    OP_CLS = 2,      // fp16 GEMV (mat_vec_kernel, gpu_kernels.h:109-139), optional fused RMSNorm
    OP_ATTN = 3,     // [RoPE +] QK^T + softmax + PV for one head per CTA
    OP_ARGMAX = 4,   // greedy sampler (argmax_kernel, gpu_kernels.h:448-493)
};

struct Seg {
    const uint32_t* w;     // [ncols][K/8] packed nibbles   (OP_CLS: fp16 weights [rows][row_stride])
    const uint32_t* z;     // [ncols][zh]
    const uint16_t* s;     // [ncols][G]
    half* out;             // output vector of this matrix
    int ncols;
    int loff;              // element offset added to out together with pos*pos_stride (KV-cache row), see pos_stride
    int pos_stride;        // 0: out is used as is; else out += loff + pos * pos_stride  (gpu_kernels.h:224-226)
    int pad_;
};

struct Op {
    int kind;
    int K;                 // input length (OP_CLS: n)
    int T;                 // pipeline stages = ceil(K / 1024)
    int unit;              // CTA column ranges start on multiples of `unit` (16-byte alignment of the meta copies)
    int jc;                // columns (OP_CLS: rows) per ring slot
    int nseg;
    int accum;             // OP_GEMV: out = half(sum + float(out))  (residual, gpu_kernels.h:229-230)
    int sync_before;       // grid barrier before the consumers read this op's inputs
    Seg seg[3];
    // activation input
    const half* x;         // [K]
    const half* norm_w;    // fused RMSNorm weight (gpu_kernels.h:72-105) or nullptr
    const half* emb;       // x = emb[tokens[pos]] (copy_embedding_kernel, :61-69) when non-null
    const int* tokens;
    half* x_copy;          // CTA 0 stores the gathered embedding row here (the residual stream)
    // OP_CLS
    int row_stride;        // elements between rows
    float alpha;
    // OP_ATTN
    half* q;               // [n_heads*head_size]; rotated in place when rope_tab != nullptr
    const half* kraw;      // un-rotated k row of this step (written by the preceding OP_GEMV) or nullptr
    half* kcache;          // layer base [seq][kv_stride]
    const half* vcache;
    half* att_out;         // optional probabilities [n_heads][pos+1]
    half* attn_out;        // [n_heads*head_size]
    const float2* rope_tab;
    int n_heads, head_size, kv_mul, kv_stride, max_seq;
    float att_alpha;
    // OP_ARGMAX
    const half* logits;
    int vocab;
    int* tokens_out;           // pinned host token ring (SharedData::tokens)
    volatile int* pos_host;    // SharedData::pos
    int* pos_dev;              // RunState::pos
    int write_token;
};

struct InterpParams {
    const Op* ops;         // device op table, or nullptr: use `one`
    int nops;
    int nwc;               // consumer warps per CTA (blockDim.x = 32 * (nwc + 1))
    int nslots;            // ring slots
    int slot_bytes;        // bytes per slot (multiple of 128)
    int scratch_bytes;     // attention scratch
    int write_token;       // overrides Op::write_token of OP_ARGMAX when >= 0
    unsigned* sync;        // [2] grid barrier counter, exit counter (zero between launches)
    const int* pPos;       // device position
    Op one;                // inline single op (operator API)
};

constexpr int kMaxConsumerWarps = 15;
constexpr int kBarAll = 13;        // named barrier: all consumer warps
constexpr int kCtrlBytes = 1024;   // mbarriers + reduction scratch
constexpr int kHandBytes = 512;    // per consumer warp: float2[2][32]
constexpr unsigned kSpinLimit = 1u << 27;

// ------------------------------------------------------------------------------------------------
// PTX: mbarrier, bulk copy, named barriers, coherent loads
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}
// Bounded spin: a protocol bug must abort the launch (trap), never hang the device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity))
        if (++spins > kSpinLimit) asm volatile("trap;");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ void named_bar(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint4 ld_cg_v4(const void* p) {   // L2-coherent (skips L1): data written by other CTAs
    uint4 r;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ld_cg_v2(const void* p) {
    uint2 r;
    asm volatile("ld.global.cg.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ld_cg_u16(const void* p) {
    uint16_t r;
    asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(r) : "l"(p));
    return (uint32_t)r;
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t r;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(addr));
    return r;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    uint16_t r;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(r) : "r"(addr));
    return (uint32_t)r;
}
// (acc.lo, acc.hi) = (w0*x.lo + acc.lo, w1*x.hi + acc.hi): two single-rounding fmaf in one FFMA2
__device__ __forceinline__ void ffma2_pk(unsigned long long& acc, float w0, float w1, unsigned long long x) {
    asm("{ .reg .b64 rw;\n\t"
        "mov.b64 rw, {%1, %2};\n\t"
        "fma.rn.f32x2 %0, rw, %3, %0; }"
        : "+l"(acc) : "f"(w0), "f"(w1), "l"(x));
}
__device__ __forceinline__ unsigned long long pack_f2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack_f2(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

// ------------------------------------------------------------------------------------------------
// Grid barrier between dependent ops.  `target` = arrivals expected so far (ordinal * gridDim.x).
// Called by all consumer threads of every CTA; the producer warp never takes part.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target, int nthreads, int ctid) {
    named_bar(kBarAll, nthreads);                 // this CTA's stores are issued
    if (ctid == 0) {
        __threadfence();
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(counter), "r"(1u) : "memory");
        unsigned v, spins = 0;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
            if (++spins > kSpinLimit) asm volatile("trap;");
        } while (v < target);
        __threadfence();
    }
    named_bar(kBarAll, nthreads);
}

// ------------------------------------------------------------------------------------------------
// Work split: CTA b of nb owns columns [c0, c1) of the op's concatenated column space.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int op_total_cols(const Op& op) {
    int n = 0;
    for (int s = 0; s < op.nseg; s++) n += op.seg[s].ncols;
    return (op.kind == OP_FFN) ? op.seg[0].ncols : n;
}
__device__ __forceinline__ void cta_range(const Op& op, int b, int nb, int& c0, int& c1) {
    const int units = op_total_cols(op) / op.unit;
    c0 = (int)(((long long)units * b) / nb) * op.unit;
    c1 = (int)(((long long)units * (b + 1)) / nb) * op.unit;
}
struct Chunk {
    int seg;       // matrix
    int col;       // first column inside the matrix
    int n;         // columns
};
// next chunk starting at column c of the concatenated space (never crosses a matrix boundary)
__device__ __forceinline__ Chunk next_chunk(const Op& op, int c, int c1) {
    Chunk ch;
    int base = 0, s = 0;
    if (op.kind != OP_FFN) {
        while (s + 1 < op.nseg && c >= base + op.seg[s].ncols) { base += op.seg[s].ncols; s++; }
    }
    const int seg_end = base + op.seg[s].ncols;
    int n = op.jc;
    if (n > c1 - c) n = c1 - c;
    if (n > seg_end - c) n = seg_end - c;
    ch.seg = s; ch.col = c - base; ch.n = n;
    return ch;
}
// slot layout of a q4 chunk of n columns: [weights (A)][weights (B, FFN only)][scales A][scales B][zeros A][zeros B]
__device__ __forceinline__ int q4_col_bytes(int K) { return K >> 1; }
__device__ __forceinline__ int q4_groups(int K) { return (K + 127) >> 7; }
__device__ __forceinline__ int q4_zh(int K) { return (q4_groups(K) + 7) >> 3; }

// ------------------------------------------------------------------------------------------------
// Producer: one lane issues, in op order, every bulk copy this CTA will consume.
// ------------------------------------------------------------------------------------------------
__device__ void producer_loop(const InterpParams& P, const Op* ops, uint8_t* ring, uint32_t full0, uint32_t empty0) {
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    unsigned seq = 0;
    for (int o = 0; o < P.nops; o++) {
        const Op& op = ops[o];
        if (op.kind > OP_CLS) continue;
        int c0, c1;
        cta_range(op, blockIdx.x, gridDim.x, c0, c1);
        for (int c = c0; c < c1;) {
            const Chunk ch = next_chunk(op, c, c1);
            const unsigned slot = seq % P.nslots, use = seq / P.nslots;
            mbar_wait(empty0 + 8 * slot, (use & 1) ^ 1);
            const uint32_t dst = smem_u32(ring + (size_t)slot * P.slot_bytes);
            const uint32_t bar = full0 + 8 * slot;
            if (op.kind == OP_CLS) {
                const uint32_t bytes = (uint32_t)ch.n * op.K * 2;
                mbar_arrive_expect_tx(bar, bytes);
                if (op.row_stride == op.K) {
                    bulk_g2s(dst, (const half*)op.seg[0].w + (size_t)ch.col * op.row_stride, bytes, bar, policy);
                } else {
                    for (int r = 0; r < ch.n; r++)
                        bulk_g2s(dst + r * op.K * 2, (const half*)op.seg[0].w + (size_t)(ch.col + r) * op.row_stride,
                                 op.K * 2, bar, policy);
                }
            } else {
                const int nm = (op.kind == OP_FFN) ? 2 : 1;
                const uint32_t wb = (uint32_t)ch.n * q4_col_bytes(op.K), sb = (uint32_t)ch.n * q4_groups(op.K) * 2,
                               zb = (uint32_t)ch.n * q4_zh(op.K) * 4;
                mbar_arrive_expect_tx(bar, nm * (wb + sb + zb));
                for (int m = 0; m < nm; m++) {
                    const Seg& sg = op.seg[(op.kind == OP_FFN) ? m : ch.seg];
                    bulk_g2s(dst + m * wb, (const uint8_t*)sg.w + (size_t)ch.col * q4_col_bytes(op.K), wb, bar, policy);
                    bulk_g2s(dst + nm * wb + m * sb, (const uint8_t*)sg.s + (size_t)ch.col * q4_groups(op.K) * 2, sb, bar, policy);
                    bulk_g2s(dst + nm * (wb + sb) + m * zb, (const uint8_t*)sg.z + (size_t)ch.col * q4_zh(op.K) * 4, zb, bar, policy);
                }
            }
            seq++;
            c += ch.n;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Consumer context
// ------------------------------------------------------------------------------------------------
struct Ctx {
    const InterpParams* P;
    uint8_t* ring;
    uint32_t full0, empty0;   // shared addresses of the mbarrier arrays
    float* red;               // 64 floats of reduction scratch
    float2* hand;             // [nwc][2][32]
    uint8_t* scratch;         // attention scratch
    int nwc, nthreads;        // consumer warps / threads
    int warp, lane, ctid;
    int pos;
    unsigned seq;             // ring sequence number of the next chunk (same in every warp)
    unsigned nsync;           // grid barriers taken so far
};

// RMSNorm scale 1/sqrt(mean(x^2)+eps) in the reference's association (gpu_kernels.h:73-96): virtual
// thread vt (of 1024) chains x[vt + 1024 i]^2, virtual warps are tree-summed, thread 0 adds the 32 warp
// aggregates in order.  All consumer threads call this; returns the scale in every thread.
__device__ float cta_rms_scale(Ctx& c, const half* x, int size) {
    const int ept = (size - 1) / 1024 + 1;
    for (int vw = c.warp; vw < 32; vw += c.nwc) {
        const int vt = vw * 32 + c.lane;
        float ss = 0.0f;
        for (int i = 0; i < ept; i++) {
            const int idx = vt + i * 1024;
            if (idx < size) {
                const float v = h2f_bits(ld_cg_u16(x + idx));
                ss = __fmaf_rn(v, v, ss);
            }
        }
        ss = warp_tree_sum(ss);
        if (c.lane == 0) c.red[vw] = ss;
    }
    named_bar(kBarAll, c.nthreads);
    float tot = c.red[0];
#pragma unroll
    for (int w = 1; w < 32; w++) tot = tot + c.red[w];
    tot = __fdiv_rn(tot, (float)size);
    tot = tot + 1e-5f;
    tot = __fdiv_rn(1.0f, __fsqrt_rn(tot));
    named_bar(kBarAll, c.nthreads);   // red may be reused
    return tot;
}

// fp16 bits of the (optionally normalised) activation element: half(x * (scale * w)), gpu_kernels.h:100-102
__device__ __forceinline__ uint32_t norm_h(uint32_t xh, uint32_t wh, float scale, bool norm) {
    if (!norm) return xh;
    const float v = __fmul_rn(h2f_bits(xh), __fmul_rn(scale, h2f_bits(wh)));
    return f2h_bits(v);
}

// ------------------------------------------------------------------------------------------------
// INT4 GEMV / FFN consumer
// ------------------------------------------------------------------------------------------------
// One step = one column per half-warp (64 weights per thread).  `wcol` = shared address of the column's
// packed weights, `scol` / `zcol` of its scales / zero words.  acc = (chain of reference lane 2j+swap,
// chain of reference lane 2j+1-swap).
__device__ __forceinline__ void q4_step(unsigned long long& acc, const unsigned long long (&xp)[32], uint32_t wcol,
                                        uint32_t scol, uint32_t zcol, int stage, int j, int swap) {
    const uint32_t wa_addr = wcol + stage * 512 + j * 32 + (swap ? 16 : 0);
    const uint4 wa = lds_v4(wa_addr);
    const uint4 wb = lds_v4(wa_addr ^ 16);
    const uint32_t s16 = lds_u16(scol + (stage * 8 + (j >> 1)) * 2);
    const uint32_t zw = lds_u32(zcol + stage * 4);
    const uint32_t zp = ((zw >> ((j >> 1) * 4)) & 0xFu) | 0x6400u;            // fp16 bits of 1024 + z
    const float nlo = fhfma_lo(zp, s16 ^ 0x8000u, 0.0f);                       // -(1024+z)*s, exact
    const float nhi = fhfma_lo(0x6380u, s16, nlo);                             // -(64+z)*s = 960*s + nlo, exact
#pragma unroll
    for (int qi = 0; qi < 4; qi++) {
        const uint32_t a = (qi == 0) ? wa.x : (qi == 1) ? wa.y : (qi == 2) ? wa.z : wa.w;
        const uint32_t b = (qi == 0) ? wb.x : (qi == 1) ? wb.y : (qi == 2) ? wb.z : wb.w;
        float da[8], db[8];
        dequant8(da, a, s16, nlo, nhi);
        dequant8(db, b, s16, nlo, nhi);
#pragma unroll
        for (int e = 0; e < 8; e++) ffma2_pk(acc, da[e], db[e], xp[qi * 8 + e]);
    }
}

__device__ void run_q4(Ctx& c, const Op& op) {
    const int T = op.T, K = op.K;
    const int npipes = c.nwc / T;
    const bool active = c.warp < npipes * T;
    const int pl = c.warp / T, stage = c.warp - pl * T;
    const int half_id = c.lane >> 4, j = c.lane & 15, swap = (j >> 2) & 1;
    const bool dual = (op.kind == OP_FFN);
    const int colb = q4_col_bytes(K), G = q4_groups(K), zh = q4_zh(K);
    // a thread whose two reference lanes lie beyond K (partial last trip) carries the accumulator through
    const bool lanes_live = active && (stage * 1024 + j * 64 < K);

    // ---- activation slice -> registers: xp[m] = (x[k0 + m], x[k1 + m]), k0/k1 = first k of the two owned lanes ----
    unsigned long long xp[32];
    {
        const half* xin = op.x;
        if (op.emb != nullptr) {
            const int token = op.tokens[c.pos];
            xin = op.emb + (size_t)token * K;
            if (blockIdx.x == 0 && op.x_copy != nullptr)
                for (int k = c.ctid; k < K; k += c.nthreads) op.x_copy[k] = xin[k];
        }
        const bool norm = (op.norm_w != nullptr);
        // slot 0 = reference lane 2j+swap, slot 1 = reference lane 2j+1-swap (see q4_step)
        const int k0 = stage * 1024 + j * 64 + (swap ? 32 : 0), k1 = stage * 1024 + j * 64 + (swap ? 0 : 32);
        uint4 r0[4], r1[4], n0[4], n1[4];
        if (lanes_live) {
#pragma unroll
            for (int v = 0; v < 4; v++) { r0[v] = ld_cg_v4(xin + k0 + v * 8); r1[v] = ld_cg_v4(xin + k1 + v * 8); }
            if (norm) {
#pragma unroll
                for (int v = 0; v < 4; v++) { n0[v] = ldg_stream_v4(op.norm_w + k0 + v * 8); n1[v] = ldg_stream_v4(op.norm_w + k1 + v * 8); }
            }
        }
        float scale = 1.0f;
        if (norm) scale = cta_rms_scale(c, xin, K);
        if (lanes_live) {
#pragma unroll
            for (int m = 0; m < 32; m++) {
                const int wi = m >> 1;   // 32-bit word holding element m
                const uint32_t a0 = (wi & 3) == 0 ? r0[wi >> 2].x : (wi & 3) == 1 ? r0[wi >> 2].y : (wi & 3) == 2 ? r0[wi >> 2].z : r0[wi >> 2].w;
                const uint32_t a1 = (wi & 3) == 0 ? r1[wi >> 2].x : (wi & 3) == 1 ? r1[wi >> 2].y : (wi & 3) == 2 ? r1[wi >> 2].z : r1[wi >> 2].w;
                const uint32_t h0 = (m & 1) ? (a0 >> 16) : (a0 & 0xFFFFu), h1 = (m & 1) ? (a1 >> 16) : (a1 & 0xFFFFu);
                uint32_t g0 = 0, g1 = 0;
                if (norm) {
                    const uint32_t b0 = (wi & 3) == 0 ? n0[wi >> 2].x : (wi & 3) == 1 ? n0[wi >> 2].y : (wi & 3) == 2 ? n0[wi >> 2].z : n0[wi >> 2].w;
                    const uint32_t b1 = (wi & 3) == 0 ? n1[wi >> 2].x : (wi & 3) == 1 ? n1[wi >> 2].y : (wi & 3) == 2 ? n1[wi >> 2].z : n1[wi >> 2].w;
                    g0 = (m & 1) ? (b0 >> 16) : (b0 & 0xFFFFu);
                    g1 = (m & 1) ? (b1 >> 16) : (b1 & 0xFFFFu);
                }
                xp[m] = pack_f2(h2f_bits(norm_h(h0, g0, scale, norm)), h2f_bits(norm_h(h1, g1, scale, norm)));
            }
        } else {
#pragma unroll
            for (int m = 0; m < 32; m++) xp[m] = 0ull;
        }
    }

    // ---- stream this CTA's chunks ----
    int c0, c1;
    cta_range(op, blockIdx.x, gridDim.x, c0, c1);
    float2* hand_in = c.hand + (size_t)(c.warp - 1) * 64;   // written by stage-1 (only read when stage > 0)
    float2* hand_out = c.hand + (size_t)c.warp * 64;
    const int pipe_threads = T * 32;
    const int bar_id = 1 + pl;
    int step = stage;                 // lock-step counter of this pipeline
    if (active && T > 1)
        for (int s = 0; s < stage; s++) named_bar(bar_id, pipe_threads);   // fill skew
    int pair_base = 0;                // pairs of this CTA's range before the current chunk
    for (int cc = c0; cc < c1;) {
        const Chunk ch = next_chunk(op, cc, c1);
        const unsigned slot = c.seq % c.P->nslots, use = c.seq / c.P->nslots;
        mbar_wait(c.full0 + 8 * slot, use & 1);
        const int npairs = dual ? ch.n : (ch.n >> 1);
        if (active) {
            const int nm = dual ? 2 : 1;
            const uint32_t base = smem_u32(c.ring + (size_t)slot * c.P->slot_bytes);
            const uint32_t wb = (uint32_t)ch.n * colb, sb = (uint32_t)ch.n * G * 2, zb = (uint32_t)ch.n * zh * 4;
            const Seg& sg = op.seg[dual ? 0 : ch.seg];
            // first pair of this chunk owned by my pipeline: global pair index q = pl (mod npipes)
            int p = (pl - pair_base % npipes + npipes) % npipes;
            for (; p < npairs; p += npipes) {
                const int lcol = dual ? p : 2 * p + half_id;          // column inside the chunk
                const int m = dual ? half_id : 0;                      // matrix (gate / up)
                const uint32_t wcol = base + m * wb + (uint32_t)lcol * colb;
                const uint32_t scol = base + nm * wb + m * sb + (uint32_t)lcol * G * 2;
                const uint32_t zcol = base + nm * (wb + sb) + m * zb + (uint32_t)lcol * zh * 4;
                unsigned long long acc = 0ull;
                if (stage > 0) {
                    const float2 in = hand_in[((step - 1) & 1) * 32 + c.lane];
                    acc = pack_f2(in.x, in.y);
                }
                if (lanes_live) q4_step(acc, xp, wcol, scol, zcol, stage, j, swap);
                if (stage == T - 1) {
                    // cub::WarpReduce order over the 32 reference lanes (shfl_down 1,2,4,8,16): the first level
                    // pairs lanes (2j, 2j+1) = this thread's two chains, the rest is a butterfly over 16 threads
                    float a0, a1;
                    unpack_f2(acc, a0, a1);
                    float v = a0 + a1;
                    v = v + __shfl_xor_sync(0xffffffffu, v, 1);
                    v = v + __shfl_xor_sync(0xffffffffu, v, 2);
                    v = v + __shfl_xor_sync(0xffffffffu, v, 4);
                    v = v + __shfl_xor_sync(0xffffffffu, v, 8);
                    if (dual) {
                        const float u = __shfl_sync(0xffffffffu, v, 16);
                        if (c.lane == 0) {   // gpu_kernels.h:269-273
                            float val = v;
                            val = __fmul_rn(val, __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-val))));
                            val = __fmul_rn(val, u);
                            sg.out[ch.col + p] = __float2half_rn(val);
                        }
                    } else if (j == 0) {
                        half* dst = sg.out;
                        if (sg.pos_stride != 0) dst += sg.loff + (size_t)c.pos * sg.pos_stride;
                        const int col = ch.col + lcol;
                        float sum = v;
                        if (op.accum) sum = sum + h2f_bits(ld_cg_u16(dst + col));
                        dst[col] = __float2half_rn(sum);
                    }
                } else {
                    float a0, a1;
                    unpack_f2(acc, a0, a1);
                    hand_out[(step & 1) * 32 + c.lane] = make_float2(a0, a1);
                }
                if (T > 1) named_bar(bar_id, pipe_threads);
                step++;
            }
        }
        __syncwarp();
        if (c.lane == 0) mbar_arrive(c.empty0 + 8 * slot);
        pair_base += npairs;
        c.seq++;
        cc += ch.n;
    }
    if (active && T > 1)
        for (int s = 0; s < T - 1 - stage; s++) named_bar(bar_id, pipe_threads);   // drain skew
}

// ------------------------------------------------------------------------------------------------
// fp16 classifier consumer (mat_vec_kernel, gpu_kernels.h:109-139).  Reference lane L chains
// k = (trip*32 + L)*8 + el over trips; stage i of a pipeline owns trips 4i..4i+3 (1024 k).  One thread is
// one reference lane, a warp step is two rows.
// ------------------------------------------------------------------------------------------------
__device__ void run_cls(Ctx& c, const Op& op) {
    const int T = op.T, n = op.K;
    const int npipes = c.nwc / T;
    const bool active = c.warp < npipes * T;
    const int pl = c.warp / T, stage = c.warp - pl * T;
    const int lane = c.lane;

    uint4 xh[4];   // 8 halves per trip
    {
        const bool norm = (op.norm_w != nullptr);
        uint4 wn[4];
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int jx = ((stage * 4 + t) * 32 + lane) * 8;
            xh[t] = make_uint4(0, 0, 0, 0);
            wn[t] = make_uint4(0, 0, 0, 0);
            if (active && jx < n) {
                xh[t] = ld_cg_v4(op.x + jx);
                if (norm) wn[t] = ldg_stream_v4(op.norm_w + jx);
            }
        }
        if (norm) {
            const float scale = cta_rms_scale(c, op.x, n);
#pragma unroll
            for (int t = 0; t < 4; t++) {
                uint32_t* xv = &xh[t].x;
                const uint32_t* wv = &wn[t].x;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint32_t lo = norm_h(xv[q] & 0xFFFFu, wv[q] & 0xFFFFu, scale, true);
                    const uint32_t hi = norm_h(xv[q] >> 16, wv[q] >> 16, scale, true);
                    xv[q] = lo | (hi << 16);
                }
            }
        }
    }

    int c0, c1;
    cta_range(op, blockIdx.x, gridDim.x, c0, c1);
    float2* hand_in = c.hand + (size_t)(c.warp - 1) * 64;
    float2* hand_out = c.hand + (size_t)c.warp * 64;
    const int pipe_threads = T * 32, bar_id = 1 + pl;
    int step = stage;
    if (active && T > 1)
        for (int s = 0; s < stage; s++) named_bar(bar_id, pipe_threads);
    int pair_base = 0;
    for (int cc = c0; cc < c1;) {
        const Chunk ch = next_chunk(op, cc, c1);
        const unsigned slot = c.seq % c.P->nslots, use = c.seq / c.P->nslots;
        mbar_wait(c.full0 + 8 * slot, use & 1);
        const int npairs = (ch.n + 1) >> 1;
        if (active) {
            const uint32_t base = smem_u32(c.ring + (size_t)slot * c.P->slot_bytes);
            int p = (pl - pair_base % npipes + npipes) % npipes;
            for (; p < npairs; p += npipes) {
                float acc[2] = {0.0f, 0.0f};
                if (stage > 0) {
                    const float2 in = hand_in[((step - 1) & 1) * 32 + lane];
                    acc[0] = in.x; acc[1] = in.y;
                }
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    const int lrow = 2 * p + r;
                    if (lrow < ch.n) {
                        const uint32_t row = base + (uint32_t)lrow * n * 2;
                        float a = acc[r];
#pragma unroll
                        for (int t = 0; t < 4; t++) {
                            const int jx = ((stage * 4 + t) * 32 + lane) * 8;
                            if (jx < n) {
                                const uint4 w = lds_v4(row + jx * 2);
                                const uint4 xv = xh[t];
                                a = fhfma_ll(w.x, xv.x, a); a = fhfma_hh(w.x, xv.x, a);
                                a = fhfma_ll(w.y, xv.y, a); a = fhfma_hh(w.y, xv.y, a);
                                a = fhfma_ll(w.z, xv.z, a); a = fhfma_hh(w.z, xv.z, a);
                                a = fhfma_ll(w.w, xv.w, a); a = fhfma_hh(w.w, xv.w, a);
                            }
                        }
                        acc[r] = a;
                    }
                }
                if (stage == T - 1) {
#pragma unroll
                    for (int r = 0; r < 2; r++) {
                        const float tot = warp_tree_sum(acc[r]);
                        const int lrow = 2 * p + r;
                        if (lane == 0 && lrow < ch.n)
                            op.seg[0].out[ch.col + lrow] = __float2half_rn(__fmul_rn(tot, op.alpha));
                    }
                } else {
                    hand_out[(step & 1) * 32 + lane] = make_float2(acc[0], acc[1]);
                }
                if (T > 1) named_bar(bar_id, pipe_threads);
                step++;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(c.empty0 + 8 * slot);
        pair_base += npairs;
        c.seq++;
        cc += ch.n;
    }
    if (active && T > 1)
        for (int s = 0; s < T - 1 - stage; s++) named_bar(bar_id, pipe_threads);
}

// ------------------------------------------------------------------------------------------------
// Attention for one head per CTA: [RoPE on q and on the new k row] -> QK^T -> softmax -> PV
// (RoPERotation_kernel :332-355, mat_vec_kernel_simple :142-168, softmax_kernel :357-401,
//  vec_mat_kernel :279-329).  The reference's 1024-thread reductions are replayed with virtual threads.
// ------------------------------------------------------------------------------------------------
__device__ void run_attn(Ctx& c, const Op& op) {
    const int hs = op.head_size, nt = c.nthreads, tid = c.ctid, lane = c.lane, warp = c.warp;
    float* qs = reinterpret_cast<float*>(c.scratch);      // hs
    float* krow = qs + hs;                                 // hs: rotated k row of this step
    float* att = krow + hs;                                // max_seq
    float* part = att + ((op.max_seq + 3) & ~3);           // 32 * hs
    float* red = c.red;
    const int pos = c.pos, size = pos + 1;
    for (int h = blockIdx.x; h < op.n_heads; h += gridDim.x) {
        const int kvh = h / op.kv_mul;
        half* kbase = op.kcache + (size_t)kvh * hs;
        const half* vbase = op.vcache + (size_t)kvh * hs;
        // ---- q (and the new k row): load, rotate, keep in shared memory ----
        if (op.rope_tab != nullptr) {
            for (int i = tid; i < hs / 2; i += nt) {
                const float2 cs = op.rope_tab[(size_t)pos * (hs / 2) + i];
                half* q = op.q + (size_t)h * hs;
                const float q0 = h2f_bits(ld_cg_u16(q + i)), q1 = h2f_bits(ld_cg_u16(q + i + hs / 2));
                const half o0 = __float2half_rn(__fmaf_rn(q0, cs.x, -__fmul_rn(q1, cs.y)));
                const half o1 = __float2half_rn(__fmaf_rn(q1, cs.x, __fmul_rn(q0, cs.y)));
                q[i] = o0; q[i + hs / 2] = o1;
                qs[i] = __half2float(o0); qs[i + hs / 2] = __half2float(o1);
                const half* kr = op.kraw + (size_t)kvh * hs;
                const float k0 = h2f_bits(ld_cg_u16(kr + i)), k1 = h2f_bits(ld_cg_u16(kr + i + hs / 2));
                const half r0 = __float2half_rn(__fmaf_rn(k0, cs.x, -__fmul_rn(k1, cs.y)));
                const half r1 = __float2half_rn(__fmaf_rn(k0, cs.y, __fmul_rn(k1, cs.x)));
                krow[i] = __half2float(r0); krow[i + hs / 2] = __half2float(r1);
                if (h == kvh * op.kv_mul) {   // one head per kv group stores the rotated row into the cache
                    half* kd = kbase + (size_t)pos * op.kv_stride;
                    kd[i] = r0; kd[i + hs / 2] = r1;
                }
            }
        } else {
            for (int i = tid; i < hs; i += nt) {
                qs[i] = h2f_bits(ld_cg_u16(op.q + (size_t)h * hs + i));
                krow[i] = h2f_bits(ld_cg_u16(kbase + (size_t)pos * op.kv_stride + i));
            }
        }
        named_bar(kBarAll, nt);
        // ---- scores: one warp per t ----
        const int nser = hs / 32;
        for (int t = warp; t < size; t += c.nwc) {
            float sum = 0.0f;
            if (t == pos) {
                for (int i = 0; i < nser; i++) sum = __fmaf_rn(krow[i * 32 + lane], qs[i * 32 + lane], sum);
            } else {
                const half* kr = kbase + (size_t)t * op.kv_stride;
                for (int i = 0; i < nser; i++)
                    sum = __fmaf_rn(h2f_bits(ld_cg_u16(kr + i * 32 + lane)), qs[i * 32 + lane], sum);
            }
            sum = warp_tree_sum(sum);
            sum = __fmul_rn(sum, op.att_alpha);
            if (lane == 0) att[t] = __half2float(__float2half_rn(sum));
        }
        named_bar(kBarAll, nt);
        // ---- softmax (idle reference threads seed the max with 0, gpu_kernels.h:374) ----
        float mx = (size < 1024) ? 0.0f : -INFINITY;
        for (int i = tid; i < size; i += nt) mx = fmaxf(mx, att[i]);
        mx = warp_max(mx);
        if (lane == 0) red[32 + warp] = mx;
        named_bar(kBarAll, nt);
        mx = red[32];
        for (int w = 1; w < c.nwc; w++) mx = fmaxf(mx, red[32 + w]);
        for (int vw = warp; vw < 32; vw += c.nwc) {
            const int vt = vw * 32 + lane;
            float ssum = 0.0f;
            for (int i = vt; i < size; i += 1024) {
                const float e = expf(__fsub_rn(att[i], mx));
                att[i] = e;
                ssum = __fadd_rn(ssum, e);
            }
            ssum = warp_tree_sum(ssum);
            if (lane == 0) red[vw] = ssum;
        }
        named_bar(kBarAll, nt);
        float tot = red[0];
#pragma unroll
        for (int w = 1; w < 32; w++) tot = tot + red[w];
        for (int i = tid; i < size; i += nt) {
            const __half pr = __float2half_rn(__fdiv_rn(att[i], tot));
            att[i] = __half2float(pr);
            if (op.att_out != nullptr) op.att_out[(size_t)h * size + i] = pr;
        }
        named_bar(kBarAll, nt);
        // ---- PV: reference lane tx chains t = 32 e + tx (e ascending); then the cub tree over tx ----
        {
            const int per_lane = hs / 32;
            for (int tx = warp; tx < 32; tx += c.nwc) {
                float a[8];
#pragma unroll
                for (int q = 0; q < 8; q++) a[q] = 0.0f;
                for (int t = tx; t < size; t += 32) {
                    const half* vr = vbase + (size_t)t * op.kv_stride + lane * per_lane;
                    const float pt = att[t];
#pragma unroll
                    for (int q = 0; q < 8; q++)
                        if (q < per_lane) a[q] = __fmaf_rn(h2f_bits(ld_cg_u16(vr + q)), pt, a[q]);
                }
#pragma unroll
                for (int q = 0; q < 8; q++)
                    if (q < per_lane) part[tx * hs + lane * per_lane + q] = a[q];
            }
        }
        named_bar(kBarAll, nt);
        for (int i = tid; i < hs; i += nt) {
            float v[32];
#pragma unroll
            for (int w = 0; w < 32; w++) v[w] = part[w * hs + i];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
#pragma unroll
                for (int w = 0; w < 32; w += 2 * o) v[w] = v[w] + v[w + o];
            op.attn_out[(size_t)h * hs + i] = __float2half_rn(v[0]);
        }
        named_bar(kBarAll, nt);
    }
}

// ------------------------------------------------------------------------------------------------
// Greedy sampler on CTA 0 (argmax_kernel, gpu_kernels.h:448-493).  Equal maxima: lowest index.
// ------------------------------------------------------------------------------------------------
__device__ void run_argmax(Ctx& c, const Op& op, int write_token) {
    if (blockIdx.x != 0) return;
    float* smax = c.red;
    int* sidx = reinterpret_cast<int*>(c.red + 32);
    float max_val = -INFINITY;
    int max_pos = 0x7fffffff;
    for (int i = c.ctid; i < op.vocab; i += c.nthreads) {
        const float v = h2f_bits(ld_cg_u16(op.logits + i));
        if (v > max_val) { max_val = v; max_pos = i; }
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, max_val, o);
        const int op_ = __shfl_xor_sync(0xffffffffu, max_pos, o);
        if (ov > max_val || (ov == max_val && op_ < max_pos)) { max_val = ov; max_pos = op_; }
    }
    if (c.lane == 0) { smax[c.warp] = max_val; sidx[c.warp] = max_pos; }
    named_bar(kBarAll, c.nthreads);
    if (c.ctid == 0) {
        for (int w = 1; w < c.nwc; w++)
            if (smax[w] > max_val || (smax[w] == max_val && sidx[w] < max_pos)) { max_val = smax[w]; max_pos = sidx[w]; }
        int token_pos = *op.pos_host;
        token_pos++;
        if (write_token) op.tokens_out[token_pos] = max_pos;
        __threadfence_system();
        *op.pos_host = token_pos;      // unblocks the CPU (pinned host memory)
        *op.pos_dev = token_pos;
    }
    named_bar(kBarAll, c.nthreads);
}

// ------------------------------------------------------------------------------------------------
// The kernel
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512, 1) interp_kernel(const __grid_constant__ InterpParams P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const Op* ops = (P.ops != nullptr) ? P.ops : &P.one;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);                 // full[nslots], empty[nslots]
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * P.nslots;
    float* red = reinterpret_cast<float*>(smem + 512);                   // 64 floats
    float2* hand = reinterpret_cast<float2*>(smem + kCtrlBytes);
    uint8_t* scratch = smem + kCtrlBytes + P.nwc * kHandBytes;
    uint8_t* ring = scratch + P.scratch_bytes;

    if (threadIdx.x == 0) {
        for (int s = 0; s < P.nslots; s++) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, P.nwc);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (warp == P.nwc) {
        if (lane == 0) producer_loop(P, ops, ring, full0, empty0);
        return;
    }

    Ctx c;
    c.P = &P; c.ring = ring; c.full0 = full0; c.empty0 = empty0; c.red = red; c.hand = hand; c.scratch = scratch;
    c.nwc = P.nwc; c.nthreads = P.nwc * 32; c.warp = warp; c.lane = lane; c.ctid = threadIdx.x;
    c.pos = (P.pPos != nullptr) ? *P.pPos : 0;
    c.seq = 0; c.nsync = 0;

    for (int o = 0; o < P.nops; o++) {
        const Op& op = ops[o];
        if (op.sync_before) {
            c.nsync++;
            grid_barrier(P.sync, c.nsync * gridDim.x, c.nthreads, c.ctid);
        }
        switch (op.kind) {
            case OP_GEMV:
            case OP_FFN: run_q4(c, op); break;
            case OP_CLS: run_cls(c, op); break;
            case OP_ATTN: run_attn(c, op); break;
            case OP_ARGMAX: run_argmax(c, op, (P.write_token >= 0) ? P.write_token : op.write_token); break;
            default: break;
        }
    }
    // leave the barrier counters at zero for the next launch: the last CTA out resets them
    if (c.nsync > 0) {
        named_bar(kBarAll, c.nthreads);
        if (c.ctid == 0) {
            const unsigned old = atomicAdd(P.sync + 1, 1u);
            if (old == gridDim.x - 1) {
                P.sync[0] = 0;
                P.sync[1] = 0;
                __threadfence();
            }
        }
    }
}

}  // namespace lq4

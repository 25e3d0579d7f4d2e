// interp_sm100.cuh -- the persistent decode kernel for sm_100a ("op interpreter").
//
// One launch walks a table of ops (the whole per-token forward pass of llama2_q4.cu:286-340 plus the
// greedy sampler, or a single op for the operator API) with ONE CTA PER SM, 11 consumer warps + 1 producer warp:
//
//   * the PRODUCER lane streams every weight byte the CTA will need with 1-D TMA bulk copies
//     (cp.async.bulk + mbarrier complete_tx) into ONE FIFO ring of shared-memory slots, a slot being a run of
//     whole columns (4-11 KB: the TMA unit sustains about one copy per 70 cycles per SM, so copies must be
//     large).  Weights never depend on activations, so the producer runs ahead across op boundaries: HBM keeps
//     streaming while the consumers wait on a dependency.  The slot size is per op ("ring epochs"); scales and
//     zero points of the CTA's column range travel the same way into a ping-pong pair of buffers.
//   * each CONSUMER warp owns whole output columns: a warp-task is 4 columns (two per half-warp) of an
//     INT4 matrix, or 2 gate/up column pairs, or 1-4 classifier rows, streamed trip by trip
//     (1024 k per trip, the reference's loop trip, gpu_kernels.h:176-201).  There is NO inter-warp
//     synchronisation inside an op: the accumulators of a column stay in one thread's registers from the
//     first trip to the last, which keeps every per-lane FMA chain of the reference in its original order
//     (results are bit-identical).
//   * one thread owns TWO reference lanes (2j, 2j+1) of two columns: one packed FFMA2 (fma.rn.f32x2)
//     advances both lane chains of a column with a natural (x[k], x[k+32]) register pair that is read
//     from shared memory once and used for both columns.  The activation vector is staged per op into
//     shared memory as fp32 pairs in exactly that order; a fused RMSNorm is computed on the way, in registers,
//     in the reference's summation order (stage_norm).
//   * ACTIVATIONS CARRY THEIR OWN SYNCHRONISATION ("flag-in-data"): every element travels as one 32-bit word
//     tag<<16 | fp16, the tag naming the op and launch that wrote it; a reader re-reads a word until the tag is
//     the expected one.  No grid barrier separates two ops, except in front of the argmax.
//   * attention runs one head per CTA; K and V rows stream through a four-deep ring of 32-row tiles filled by
//     cp.async (run_attn_t).
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
    half* out;             // output vector of this matrix (fp16), or nullptr when only the tagged copy is wanted
    uint32_t* out32;       // tagged output vector (see "flag-in-data" below) or nullptr
    int ncols;
    int loff;              // element offset added to out together with pos*pos_stride (KV-cache row), see pos_stride
    int pos_stride;        // 0: out is used as is; else out += loff + pos * pos_stride  (gpu_kernels.h:224-226)
    int bcast;             // tensor parallel: the tagged output goes to every rank's copy of the buffer (same offset), not only to ours
};

struct Op {
    int kind;
    int K;                 // input length (OP_CLS: n)
    int T;                 // trips: ceil(K / 1024)  (OP_CLS: ceil(n / 256))
    int cps;               // columns (gate/up pairs, rows) per ring slot: 4, 2 or 1
    int spt;               // ring slots per warp-task
    int slot_bytes;        // ring geometry of this op: bytes per slot (multiple of 128) ...
    int nslots;            // ... and number of slots; consecutive ops with the same slot size share an epoch of the ring
    int rpt;               // OP_CLS: rows per warp-task (4, or 1 when a row is long enough to keep a warp busy)
    int au;                // CTA task ranges start on multiples of `au` tasks (16-byte alignment of the scale/zero copies)
    int ntasks;            // warp-tasks of the whole op (4 columns | 2 gate/up pairs | 4 rows each)
    int nseg;
    int accum;             // OP_GEMV: out = half(sum + float(out))  (residual, gpu_kernels.h:229-230)
    int sync_before;       // grid barrier before the consumers read this op's inputs
    Seg seg[3];
    // activation input
    const uint32_t* xt;    // [K] tagged activations written earlier in this launch by other CTAs, or nullptr: use x
    const half* x;         // [K]
    const half* norm_w;    // fused RMSNorm weight (gpu_kernels.h:72-105) or nullptr
    const half* emb;       // x = emb[tokens[pos]] (copy_embedding_kernel, :61-69) when non-null
    const int* tokens;
    half* x_copy;          // CTA 0 stores the gathered embedding row here (the residual stream); op-by-op path only
    const half* res_emb;   // accumulating GEMV: the residual is this embedding table's row tokens[pos] instead of the old output
    int res_stride;        // elements per embedding row (res_emb is already offset to this rank's first column)
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
    const uint32_t* qt;    // tagged q | un-rotated k row | v row of this step (fused path) or nullptr: q / kraw / the cache row
    const uint32_t* krawt;
    const uint32_t* vrawt;
    uint32_t* attn_out32;  // tagged output or nullptr
    int attn_bcast;        // tensor parallel: broadcast attn_out32 to every rank
    const float2* rope_tab;
    int n_heads, head_size, kv_mul, kv_stride, max_seq;
    int exp16_from;        // softmax over more than this many positions uses the arithmetic of softmax_kernel_no_smem (gpu_kernels.h:403-446)
    int att_split;         // long contexts: CTAs per head (1, 2 or 4; needs att_sc / att_flags and n_heads * att_split CTAs), see run_attn_split
    int att_sc_stride;     // elements per head of att_sc
    uint16_t* att_sc;      // [n_heads][att_sc_stride] fp16 scores exchanged between the CTAs of a head
    unsigned* att_flags;   // [n_heads][4] "my scores are written" flags of the parts (value: Ctx::op_seq)
    float att_alpha;
    // OP_ARGMAX
    const half* logits;
    int vocab;
    int* tokens_out;           // pinned host token ring (SharedData::tokens)
    volatile int* pos_host;    // SharedData::pos
    int* pos_dev;              // RunState::pos
    int write_token;
    uint32_t* cand;        // tensor parallel: [world][3] tagged (value, index low, index high) candidates of the ranks' vocabulary slices, or nullptr
    int vocab0;            // first vocabulary row of this rank's slice
};

static_assert(sizeof(Op) <= 512 && sizeof(Op) % 4 == 0, "Op must fit its shared-memory copy");

struct InterpParams {
    const Op* ops;         // device op table, or nullptr: use `one`
    int nops;
    int nwc;               // consumer warps per CTA (blockDim.x = 32 * (nwc + 1))
    int ring_bytes;        // bytes of the weight ring (cut into slots per op, see Op::slot_bytes)
    int meta_bytes;        // scale/zero buffer 0 (even INT4 ops of the table)
    int meta1_bytes;       // scale/zero buffer 1 (odd INT4 ops)
    int xs_bytes;          // activation staging area (aliased with the attention scratch)
    int write_token;       // overrides Op::write_token of OP_ARGMAX when >= 0
    unsigned seq_base;     // launch counter * nops: makes the activation tags of this launch unique
    int rank, world;       // tensor parallel: this GPU's rank; world == 1: single GPU
    uint32_t* peers[8];    // base of every rank's tagged-activation buffer as mapped into this process (peers[rank] = ours)
    unsigned* sync;        // [2] grid barrier counter, exit counter (zero between launches)
    const int* pPos;       // device position
    const int* tokens;     // pinned-host token ring of the step (SharedData::tokens): read ONCE per launch, at its start
    unsigned long long* trace;   // optional [nops + 1] timestamps (ns): CTA 0 at the start of each op, and at the end
    int trace_op;                // op whose phases every CTA records at trace[2048 + cta * 8 + k]
    int nomath;                  // development aid: consumers wait for and release their weights but skip the arithmetic (data-path-only timing)
    Op one;                // inline single op (operator API)
};

constexpr int kMaxConsumerWarps = 11;   // 12 warps = 3 per SM sub-partition: 168 registers per thread (a 4th warp on a sub-partition would cap them at 128)
constexpr int kMaxSlots = 64;
constexpr int kBarAll = 13;        // named barrier: all consumer warps
constexpr int kCtrlBytes = 4096;   // mbarriers (first 2 KB) + reduction scratch (at 3 KB)
constexpr int kRedOffset = 3072;
constexpr unsigned long long kWaitLimitNs = 5000000000ull;   // a wait longer than 5 s is a protocol bug (or a dead peer): trap
constexpr int kLapOffset = 2048;   // per-slot release counters (consumers), 64 words
constexpr int kFillBaseOffset = 2304;   // per-slot fills before the current ring epoch (consumers), 64 words
constexpr int kProdFillOffset = 3328;   // per-slot fills so far (producer lane only), 64 words
constexpr int kOpOffset = 2560;    // the current op, copied from the table (512 bytes)
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
// Bounded spin: a protocol bug -- or, under tensor parallelism, a peer that died -- must abort the launch, never hang the device.
// Before the trap the thread leaves a record in pinned host memory (g_fault, set by the host at start-up), so that the host's
// next stream query / synchronize can say WHAT timed out instead of a bare "unspecified launch failure":
//   code 1 weights (a bulk copy never completed)   2 activations (the previous op's tagged output never arrived: on a
//   tensor-parallel run, a dead or stalled peer)   3 grid barrier   4 ring slot never released   5 split-attention flags
__device__ int* g_fault = nullptr;      // [4] pinned host: code, CTA, thread, rank-local SM id
__device__ __noinline__ void protocol_timeout(int code) {
    int* f = g_fault;
    if (f != nullptr) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        f[1] = (int)blockIdx.x; f[2] = (int)threadIdx.x; f[3] = (int)smid;
        __threadfence_system();
        f[0] = code;
        __threadfence_system();
    }
    asm volatile("trap;");
}
// Bounded spin, continued: every wait below gives up after kWaitLimitNs.
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const unsigned long long t0 = global_ns();
    while (!mbar_try_wait(bar, parity))
        if (global_ns() - t0 > kWaitLimitNs) protocol_timeout(1);
}
// The producer's waits are for ring slots the consumers have yet to release.  When the consumers are themselves stuck (waiting
// for activations that never come) the producer would time out too, and first: it gets twice the limit, so that the record the
// host sees names the real cause.
__device__ __forceinline__ void mbar_wait_release(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const unsigned long long t0 = global_ns();
    while (!mbar_try_wait(bar, parity))
        if (global_ns() - t0 > 2 * kWaitLimitNs) protocol_timeout(4);
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
__device__ __forceinline__ uint32_t ld_cg_u16(const void* p) {
    uint16_t r;
    asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(r) : "l"(p));
    return (uint32_t)r;
}
// ------------------------------------------------------------------------------------------------
// Flag-in-data activations.  Inside one launch an activation vector is written by many CTAs and read by all of
// them.  Instead of a grid barrier between writer and readers, every element travels as one 32-bit word
// (tag << 16 | fp16 bits), the tag naming the op that wrote it in this launch; a reader simply re-reads a word
// until its tag is the expected one.  A 32-bit store is indivisible, so no ordering between words is needed.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_tagged(uint32_t* p, uint32_t tag, uint32_t hbits) {
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"((tag << 16) | (hbits & 0xFFFFu)) : "memory");
}
// Tensor parallel: the same word goes to the same offset of every rank's buffer, over NVLink peer mappings.  Each rank then
// polls only its own memory; no cross-GPU barrier exists anywhere.
__device__ __forceinline__ void st_tagged_all(const InterpParams& P, uint32_t* p, uint32_t tag, uint32_t hbits) {
    const size_t off = (size_t)(p - P.peers[P.rank]);
    for (int r = 0; r < P.world; r++) st_tagged(P.peers[r] + off, tag, hbits);
}
template <int F>
__device__ __forceinline__ void st_tagged_maybe_all(const InterpParams& P, bool bcast, uint32_t* p, uint32_t tag, uint32_t hbits) {
    if ((F & 2) && bcast && P.world > 1) st_tagged_all(P, p, tag, hbits);
    else st_tagged(p, tag, hbits);
}
// Two adjacent elements in one 8-byte store (every word carries its own tag, so readers need no atomicity across words;
// under tensor parallelism this halves the NVLink store packets of an exchange).  p must be 8-byte aligned.
__device__ __forceinline__ void st_tagged2(uint32_t* p, uint32_t tag, uint32_t h0, uint32_t h1) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"((tag << 16) | (h0 & 0xFFFFu)), "r"((tag << 16) | (h1 & 0xFFFFu)) : "memory");
}
template <int F>
__device__ __forceinline__ void st_tagged2_maybe_all(const InterpParams& P, bool bcast, uint32_t* p, uint32_t tag, uint32_t h0, uint32_t h1) {
    if ((F & 2) && bcast && P.world > 1) {
        const size_t off = (size_t)(p - P.peers[P.rank]);
        for (int r = 0; r < P.world; r++) st_tagged2(P.peers[r] + off, tag, h0, h1);
    } else {
        st_tagged2(p, tag, h0, h1);
    }
}
__device__ __forceinline__ uint32_t ld_tagged_any(const uint32_t* p) {      // current word, whatever its tag (residual read by its only writer)
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t poll1(const uint32_t* p, uint32_t tag) {   // fp16 bits of one element
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if ((v >> 16) != tag) {
        const unsigned long long t0 = global_ns();
        do {
            asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
            if ((v >> 16) != tag && global_ns() - t0 > kWaitLimitNs) protocol_timeout(2);
        } while ((v >> 16) != tag);
    }
    return v & 0xFFFFu;
}
__device__ __forceinline__ bool tags_ok(const uint4& v, uint32_t tag) {
    return ((v.x >> 16) == tag) & ((v.y >> 16) == tag) & ((v.z >> 16) == tag) & ((v.w >> 16) == tag);
}
__device__ __forceinline__ uint4 ld_vol_v4(const uint32_t* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
// eight consecutive elements as four words of packed fp16 pairs (the layout an 8-half vector load returns)
__device__ __forceinline__ uint4 poll8(const uint32_t* p, uint32_t tag) {
    uint4 a = ld_vol_v4(p), b = ld_vol_v4(p + 4);
    if (!(tags_ok(a, tag) && tags_ok(b, tag))) {
        const unsigned long long t0 = global_ns();
        do {
            __nanosleep(40);        // idle CTAs poll for a long time (e.g. the 116 without a head during attention): stay off the L2
            a = ld_vol_v4(p); b = ld_vol_v4(p + 4);
            if (global_ns() - t0 > kWaitLimitNs) protocol_timeout(2);
        } while (!(tags_ok(a, tag) && tags_ok(b, tag)));
    }
    return make_uint4((a.x & 0xFFFFu) | (a.y << 16), (a.z & 0xFFFFu) | (a.w << 16), (b.x & 0xFFFFu) | (b.y << 16), (b.z & 0xFFFFu) | (b.w << 16));
}

__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}
__device__ __forceinline__ void sts_v2_u32(uint32_t addr, uint2 v) {
    asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
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
#ifndef LQ4_USE_FFMA2
#define LQ4_USE_FFMA2 1
#endif
__device__ __forceinline__ void ffma2_pk(unsigned long long& acc, float w0, float w1, unsigned long long x) {
#if LQ4_USE_FFMA2
    asm("{ .reg .b64 rw;\n\t"
        "mov.b64 rw, {%1, %2};\n\t"
        "fma.rn.f32x2 %0, rw, %3, %0; }"
        : "+l"(acc) : "f"(w0), "f"(w1), "l"(x));
#else
    asm("{ .reg .f32 a0, a1, x0, x1;\n\t"
        "mov.b64 {a0, a1}, %0;\n\t"
        "mov.b64 {x0, x1}, %3;\n\t"
        "fma.rn.f32 a0, %1, x0, a0;\n\t"
        "fma.rn.f32 a1, %2, x1, a1;\n\t"
        "mov.b64 %0, {a0, a1}; }"
        : "+l"(acc) : "f"(w0), "f"(w1), "l"(x));
#endif
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
// arrive: one thread, after the CTA's named barrier (which orders the other threads' stores before this release)
__device__ __forceinline__ void grid_arrive(unsigned* counter) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(counter), "r"(1u) : "memory");
}
// wait: acquire polls by one thread; the named barrier that follows hands the ordering on to the CTA
__device__ __forceinline__ void grid_wait(unsigned* counter, unsigned target) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    if (v < target) {
        const unsigned long long t0 = global_ns();
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
            if (v < target && global_ns() - t0 > kWaitLimitNs) protocol_timeout(3);
        } while (v < target);
    }
}

__device__ __forceinline__ void lds_v2_b64(uint32_t addr, unsigned long long& a, unsigned long long& b) {
    asm volatile("ld.shared.v2.b64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}
__device__ __forceinline__ void sts_v4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void sts_v4_u32(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ------------------------------------------------------------------------------------------------
// Geometry shared by the producer, the consumers and the host planner
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int q4_groups(int K) { return (K + 127) >> 7; }
__host__ __device__ __forceinline__ int q4_zh(int K) { return (q4_groups(K) + 7) >> 3; }
__host__ __device__ __forceinline__ int q4_col_bytes(int K) { return K >> 1; }

// CTA b of nb owns warp-tasks [t0, t1) of the op
__device__ __forceinline__ void cta_task_range(const Op& op, int b, int nb, int& t0, int& t1) {
    const int au = op.au > 0 ? op.au : 1;
    const long long units = op.ntasks / au;       // the host guarantees ntasks % au == 0
    t0 = (int)((units * b) / nb) * au;
    t1 = (int)((units * (b + 1)) / nb) * au;
}
// GEMV: column c of the concatenated column space -> (matrix, column inside it)
__device__ __forceinline__ void gemv_locate(const Op& op, int c, int& seg, int& col) {
    int s = 0;
    while (s + 1 < op.nseg && c >= op.seg[s].ncols) { c -= op.seg[s].ncols; s++; }
    seg = s; col = c;
}

struct Smem {            // shared-memory map (shared-window addresses)
    uint32_t bars;       // full[64], empty[64], mfull[2], mempty[2]
    uint32_t laps;       // [64] releases of each slot so far (consumers only)
    uint32_t xs;         // activation staging / attention scratch
    uint32_t meta;       // two buffers (meta_bytes, then meta1 bytes) of scales and zero points of this CTA's columns, ping-pong per INT4 op
    int meta_bytes;
    uint32_t ring;       // [S][slot_bytes] in the current epoch
    int S, slot_bytes;   // geometry of the current ring epoch (0: none yet)
    __device__ __forceinline__ uint32_t full(int s) const { return bars + (uint32_t)s * 8; }
    __device__ __forceinline__ uint32_t empty(int s) const { return bars + (uint32_t)(kMaxSlots + s) * 8; }
    __device__ __forceinline__ uint32_t slot(int s) const { return ring + (uint32_t)s * slot_bytes; }
    __device__ __forceinline__ uint32_t mfull(int b) const { return bars + (uint32_t)(2 * kMaxSlots + b) * 8; }
    __device__ __forceinline__ uint32_t mempty(int b) const { return bars + (uint32_t)(2 * kMaxSlots + 2 + b) * 8; }
    __device__ __forceinline__ uint32_t mbuf(int b) const { return meta + (uint32_t)b * meta_bytes; }
};

// ------------------------------------------------------------------------------------------------
// Producer: one lane issues, in op order, every bulk copy this CTA will consume.  The ring is a FIFO of
// fixed-size slots; a warp-task takes `spt` consecutive slots, each filled by one copy (two for gate/up) of a
// run of whole columns -- adjacent columns are contiguous in the reference layout, so a slot is one or two
// multi-KB bulk copies (the TMA unit sustains ~1 copy per 70 cycles per SM: copies must be large).
// ------------------------------------------------------------------------------------------------
// Lookahead of the producer: walks the same (op, task, slot) sequence as the issue loop, kPrefetchAhead slots further on, and
// asks the L2 for each slot's bytes (cp.async.bulk.prefetch.L2).  A slot that could not be requested into shared memory yet --
// the ring is full, or it is draining for a new slot size -- then arrives from the L2 instead of from HBM once its turn comes.
// MEASURED AND LEFT OFF (default 0): with 24 or 48 slots of lookahead the 7B step went from 2.03 to 2.25 ms on the same build --
// the extra L2 traffic of the prefetches competes with the bulk copies that are already in flight.
#ifndef LQ4_PREFETCH_AHEAD
#define LQ4_PREFETCH_AHEAD 0
#endif
constexpr int kPrefetchAhead = LQ4_PREFETCH_AHEAD;
__device__ __forceinline__ void prefetch_l2(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
struct ProdAhead {
    int o, task, i, t1;                 // op, warp-task, slot of the task, end of this CTA's task range
    int kind, cps, spt, rpt, nseg, nc0, nc1, rows_total, row_stride, K;
    size_t colb;
    const uint8_t* w[3];
};
__device__ __forceinline__ bool ahead_open_op(const InterpParams& P, const Op* ops, ProdAhead& f) {      // first op at or after f.o with work for this CTA
    for (; f.o < P.nops; f.o++) {
        const Op* op = ops + f.o;
        if (op->kind > OP_CLS) continue;
        int t0, t1;
        cta_task_range(*op, blockIdx.x, gridDim.x, t0, t1);
        if (t1 <= t0) continue;
        f.task = t0; f.t1 = t1; f.i = 0;
        f.kind = op->kind; f.cps = op->cps; f.spt = op->spt; f.rpt = op->rpt; f.nseg = op->nseg; f.K = op->K;
        f.nc0 = op->seg[0].ncols; f.nc1 = op->seg[1].ncols; f.rows_total = op->seg[0].ncols; f.row_stride = op->row_stride;
        f.colb = (op->kind == OP_CLS) ? (size_t)op->K * 2 : (size_t)q4_col_bytes(op->K);
        f.w[0] = (const uint8_t*)op->seg[0].w; f.w[1] = (const uint8_t*)op->seg[1].w; f.w[2] = (const uint8_t*)op->seg[2].w;
        return true;
    }
    return false;
}
__device__ __forceinline__ void ahead_step(const InterpParams& P, const Op* ops, ProdAhead& f) {
    if (f.o >= P.nops) return;
    const uint32_t bytes = (uint32_t)(f.cps * f.colb);
    if (f.kind == OP_GEMV) {
        int col = f.task * 4, seg = 0;
        if (f.nseg > 1 && col >= f.nc0) { col -= f.nc0; seg = 1; if (f.nseg > 2 && col >= f.nc1) { col -= f.nc1; seg = 2; } }
        prefetch_l2(f.w[seg] + (size_t)(col + f.i * f.cps) * f.colb, bytes);
    } else if (f.kind == OP_FFN) {
        const size_t off = (size_t)(f.task * 2 + f.i * f.cps) * f.colb;
        prefetch_l2(f.w[0] + off, bytes);
        prefetch_l2(f.w[1] + off, bytes);
    } else if (f.row_stride == f.K) {
        const int row = f.task * f.rpt + f.i * f.cps;
        int rows = f.rows_total - row;
        if (rows > f.cps) rows = f.cps;
        if (rows > 0) prefetch_l2(f.w[0] + (size_t)row * f.colb, (uint32_t)(rows * f.colb));
    }
    if (++f.i == f.spt) {
        f.i = 0;
        if (++f.task == f.t1) { f.o++; ahead_open_op(P, ops, f); }
    }
}

__device__ void producer_loop(const InterpParams& P, const Op* ops, const Smem& sm) {
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    ProdAhead ahead;
    if constexpr (kPrefetchAhead > 0) {
        ahead.o = 0;
        if (ahead_open_op(P, ops, ahead))
            for (int n = 0; n < kPrefetchAhead; n++) ahead_step(P, ops, ahead);
    }
    // Ring epochs: consecutive ops with the same slot size share one; when the size changes the producer first waits for
    // every slot to be released (the consumers are then inside their hand-over / staging, which hides the refill) and
    // starts again at slot 0 with the new geometry.  pfill[i] = fills of slot i so far (parity of its barriers).
    int slot = 0, S = 0, sbytes = 0;
    const uint32_t pfill = sm.bars + kProdFillOffset;
    auto ld_fill = [&](int i) { unsigned v; asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(pfill + i * 4) : "memory"); return v; };
    unsigned mcount = 0, issued = 0;
    for (int o = 0; o < P.nops; o++) {
        if (ops[o].kind > OP_CLS) continue;
        const Op op = ops[o];         // private copy: the fields are read once, not once per copy
        int t0, t1;
        cta_task_range(op, blockIdx.x, gridDim.x, t0, t1);
        const int cps = op.cps, spt = op.spt;
        if (op.slot_bytes != sbytes) {
            for (int i = 0; i < S; i++) mbar_wait_release(sm.empty(i), (ld_fill(i) & 1) ^ 1);     // drain: last fill of every slot released
            S = op.nslots; sbytes = op.slot_bytes; slot = 0;
        }

        if (op.kind != OP_CLS) {      // scales and zero points of the CTA's columns (layout: see stage_meta_layout)
            const int b = mcount & 1;
            mbar_wait_release(sm.mempty(b), ((mcount >> 1) & 1) ^ 1);
            const uint32_t dst = sm.mbuf(b), bar = sm.mfull(b);
            const int G = q4_groups(op.K), zh = q4_zh(op.K);
            if (t1 > t0) {
                if (op.kind == OP_FFN) {
                    const int o0 = t0 * 2, nout = (t1 - t0) * 2;
                    const uint32_t sb = nout * G * 2, zb = nout * zh * 4;
                    mbar_arrive_expect_tx(bar, 2 * (sb + zb));
                    for (int m = 0; m < 2; m++) {
                        bulk_g2s(dst + m * sb, op.seg[m].s + (size_t)o0 * G, sb, bar, policy);
                        bulk_g2s(dst + 2 * sb + m * zb, op.seg[m].z + (size_t)o0 * zh, zb, bar, policy);
                    }
                } else {
                    const int c0 = t0 * 4, ncol = (t1 - t0) * 4;
                    const uint32_t zoff = ncol * G * 2;
                    mbar_arrive_expect_tx(bar, (uint32_t)ncol * (G * 2 + zh * 4));
                    int done = 0;
                    while (done < ncol) {                                 // one piece per matrix the range touches
                        int seg, col;
                        gemv_locate(op, c0 + done, seg, col);
                        int n = op.seg[seg].ncols - col;
                        if (n > ncol - done) n = ncol - done;
                        bulk_g2s(dst + done * G * 2, op.seg[seg].s + (size_t)col * G, n * G * 2, bar, policy);
                        bulk_g2s(dst + zoff + done * zh * 4, op.seg[seg].z + (size_t)col * zh, n * zh * 4, bar, policy);
                        done += n;
                    }
                }
            } else {
                mbar_arrive(bar);
            }
            mcount++;
        }
        const size_t colb = (op.kind == OP_CLS) ? (size_t)op.K * 2 : (size_t)q4_col_bytes(op.K);
        for (int task = t0; task < t1; task++) {
            int seg = 0, col = 0;
            if (op.kind == OP_GEMV) gemv_locate(op, task * 4, seg, col);
            for (int i = 0; i < spt; i++) {
                const unsigned pf = ld_fill(slot);
                mbar_wait_release(sm.empty(slot), (pf & 1) ^ 1);
                const uint32_t dst = sm.ring + (uint32_t)slot * sbytes, bar = sm.full(slot);
                if (op.kind == OP_GEMV) {
                    const uint32_t bytes = (uint32_t)(cps * colb);
                    mbar_arrive_expect_tx(bar, bytes);
                    bulk_g2s(dst, (const uint8_t*)op.seg[seg].w + (size_t)(col + i * cps) * colb, bytes, bar, policy);
                } else if (op.kind == OP_FFN) {      // slot: [gate cps columns][up cps columns]
                    const uint32_t bytes = (uint32_t)(cps * colb);
                    mbar_arrive_expect_tx(bar, 2 * bytes);
                    const size_t off = (size_t)(task * 2 + i * cps) * colb;
                    bulk_g2s(dst, (const uint8_t*)op.seg[0].w + off, bytes, bar, policy);
                    bulk_g2s(dst + bytes, (const uint8_t*)op.seg[1].w + off, bytes, bar, policy);
                } else {
                    const int row = task * op.rpt + i * cps;
                    int rows = op.seg[0].ncols - row;
                    if (rows > cps) rows = cps;
                    if (rows < 0) rows = 0;
                    mbar_arrive_expect_tx(bar, (uint32_t)(rows * colb));
                    const half* w = (const half*)op.seg[0].w + (size_t)row * op.row_stride;
                    if (op.row_stride == op.K) {
                        if (rows > 0) bulk_g2s(dst, w, (uint32_t)(rows * colb), bar, policy);
                    } else {
                        for (int r = 0; r < rows; r++) bulk_g2s(dst + r * colb, w + (size_t)r * op.row_stride, (uint32_t)colb, bar, policy);
                    }
                }
                asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(pfill + slot * 4), "r"(pf + 1) : "memory");
#ifdef LQ4_PROD_TRACE      // diagnosis builds only: when the producer issued each chunk of the traced op (SM clock)
                if (P.trace != nullptr && o == P.trace_op) { const int k = (task - t0) * spt + i; if (k < 32) P.trace[27136 + blockIdx.x * 32 + k] = (unsigned long long)clock64(); }
#endif
                if constexpr (kPrefetchAhead > 0) ahead_step(P, ops, ahead);      // keep the L2 lookahead kPrefetchAhead slots in front
                if (++slot == S) slot = 0;
                issued++;
                asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(sm.bars + 3584), "r"(issued) : "memory");
            }
        }

    }
}

// ------------------------------------------------------------------------------------------------
// Consumer context
// ------------------------------------------------------------------------------------------------
struct Ctx {
    const InterpParams* P;
    Smem sm;
    float* red;               // 64 floats of reduction scratch
    uint8_t* scratch;         // generic pointer to the xs area (attention scratch)
    int nwc, nthreads;        // consumer warps / threads
    int warp, lane, ctid;
    int pos;
    int token;                // tokens[pos], fetched over PCIe at the start of the launch
    unsigned qbase;           // ring chunks of all ops before the current one in this ring epoch (this CTA)
    unsigned qtotal;          // ... since the start of the launch (development aid)
    unsigned mcount;          // INT4 ops so far (scale/zero buffer = mcount & 1)
    int meta_pending;         // scale/zero buffer to hand back to the producer once every warp has left the op, or -1
    unsigned nsync;           // grid barriers taken so far
    uint32_t tag_in, tag_out; // activation tags: of the op whose output this op reads, and of this op
    unsigned op_seq;          // launch counter * nops + op index + 1: unique per (launch, op) for 2^30 ops (split-attention flags)
    unsigned long long* tr;   // detailed phase trace of the current op (this CTA's 8 entries) or nullptr
};
// Kernel instances.  The persistent kernel is compiled several times, each instance holding only the code its launches can
// reach: merely carrying the long-context attention cost the 7B step 8 % (register allocation of the staging and task loops),
// the tensor-parallel stores and the development aids (phase trace, LQ4_NOMATH) another 3 %.
constexpr int kSplit = 1;      // several CTAs per head for long contexts (run_attn_split); launched once the position passes kAttnSplitFrom
constexpr int kTP = 2;         // tagged outputs are broadcast to every rank's buffer (world > 1)
constexpr int kDev = 4;        // phase trace and the no-arithmetic timing mode
template <int F>
struct CtxT : Ctx {};

template <int F>
__device__ __forceinline__ void cyc_mark(const CtxT<F>& c, int k) {      // SM clock, warp 0 lane 0: sub-microsecond phases
    if constexpr (!(F & kDev)) return;
    if (c.tr != nullptr && c.lane == 0 && c.warp == 0) (c.tr - blockIdx.x * 8 + 148 * 8 + blockIdx.x * 16)[k] = (unsigned long long)clock64();
}
// development aid: a raw value (not a clock) in slot k (8..15) of the CTA's cycle record
template <int F>
__device__ __forceinline__ void val_mark(const CtxT<F>& c, int k, unsigned long long v) {
    if constexpr (!(F & kDev)) return;
    if (c.tr != nullptr && c.lane == 0 && c.warp == 0) (c.tr - blockIdx.x * 8 + 148 * 8 + blockIdx.x * 16)[k] = v;
}
__device__ __forceinline__ unsigned producer_issued(const Ctx& c) {
    unsigned v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(c.sm.bars + 3584) : "memory");
    return v;
}
// development aid, SM clock: per consumer warp, when its first task's weights were there (k = 0), when its 1st..5th task was done
// (k = 1..5) and when it left the op (k = 7).  Record of warp w of CTA b: trace[8192 + b * 128 + w * 8 + k].
template <int F>
__device__ __forceinline__ void warp_mark(const CtxT<F>& c, int k) {
    if constexpr (!(F & kDev)) return;
    if (c.tr != nullptr && c.lane == 0 && k < 8) (c.tr - blockIdx.x * 8 - 2048 + 8192 + blockIdx.x * 128 + c.warp * 8)[k] = (unsigned long long)clock64();
}
template <int F>
__device__ __forceinline__ void trace_mark(const CtxT<F>& c, int k) {
    if constexpr (!(F & kDev)) return;
    if (c.tr != nullptr && c.lane == 0 && (c.warp == 0 || k >= 8)) c.tr[k & 7] = global_ns();
}

// fp16 bits of the (optionally normalised) activation element: half(x * (scale * w)), gpu_kernels.h:100-102
__device__ __forceinline__ uint32_t norm_h(uint32_t xh, uint32_t wh, float scale, bool norm) {
    if (!norm) return xh;
    const float v = __fmul_rn(h2f_bits(xh), __fmul_rn(scale, h2f_bits(wh)));
    return f2h_bits(v);
}
__device__ __forceinline__ uint32_t word_of(const uint4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
__device__ __forceinline__ uint32_t pack_duo(const uint2& v) { return (v.x & 0xFFFFu) | (v.y << 16); }
__device__ __forceinline__ bool tags_ok2(const uint2& v, uint32_t tag) { return ((v.x >> 16) == tag) & ((v.y >> 16) == tag); }
__device__ __forceinline__ uint2 ld_vol_v2(const uint32_t* p) {
    uint2 v;
    asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_cg_u32(const void* p) {
    uint32_t r;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// Activation staging.  Shared-memory layout read by the INT4 trips: thread (half-warp lane j) of a consumer warp owns
// reference lanes A = 2j + sw and B = 2j + 1 - sw, sw = (j>>2)&1 (the swap keeps its two 16-byte weight loads bank-conflict
// free).  Pair (trip t, j, i) = (x[t*1024 + A*32 + i], x[t*1024 + B*32 + i]) lives at byte t*kTripBytes + (i>>1)*kRowBytes +
// j*16 + (i&1)*8: one 16-byte load per thread fetches two pairs and the 16 threads of a half-warp read 256 contiguous bytes.
// The unit of staging is that 16-byte SLOT (t, i2, j): two consecutive elements of lane A and of lane B, i.e. two 8-byte
// tagged vectors in, one 16-byte vector of pairs out.  Sixteen threads with consecutive i2 read one whole 128-byte line of
// tagged words (polling by sectors multiplies the L2 requests of 148 CTAs); the row padding lets the same sixteen threads
// write their slots to sixteen rows without bank conflicts.  All tagged vectors a thread needs are requested before the
// first tag is examined, so a hand-over costs one trip to L2 once the data is there.
// ------------------------------------------------------------------------------------------------
constexpr int kRowBytes = 272;      // 16 slots of 16 bytes + 16 bytes of padding: rows start one bank group apart
constexpr int kTripBytes = 16 * kRowBytes;
constexpr int kNormMaxT = 8;        // fused RMSNorm: K <= 8192 (the planner refuses larger)
constexpr int kNormThreads = 256;   // 16 j x 16 slots of a lane pair
#ifndef LQ4_NORM_CHUNK
#define LQ4_NORM_CHUNK 4
#endif
#ifndef LQ4_PLAIN_CHUNK
#define LQ4_PLAIN_CHUNK 8
#endif
constexpr int kNormChunk = LQ4_NORM_CHUNK;      // trips of a slot column whose tagged vectors are in flight together (2 vectors each)
constexpr int kPollChunk = LQ4_PLAIN_CHUNK;     // slots of the plain staging in flight together (2 vectors each)

struct SlotMap {       // slot -> (j, i2) and the element offsets of its two lanes inside a trip
    int j, i2, offA, offB;
};
// Slot r (0..255) of a trip: i2 fastest, so a half-warp covers one lane pair (butterfly over i2 inside the warp)
__device__ __forceinline__ SlotMap slot_map(int r) {
    SlotMap m;
    m.j = r >> 4;
    m.i2 = r & 15;
    const int sw = (m.j >> 2) & 1;
    m.offA = (2 * m.j + sw) * 32 + m.i2 * 2;
    m.offB = (2 * m.j + 1 - sw) * 32 + m.i2 * 2;
    return m;
}

// Poll the tagged vectors at addr(i) (i in `want`) until every one carries `tag`; only the missing ones are requested again.
template <int N, typename Addr>
__device__ __forceinline__ void poll_vecs(uint2 (&v)[N], Addr addr, uint32_t want, uint32_t tag) {
#pragma unroll
    for (int i = 0; i < N; i++)
        if ((want >> i) & 1u) v[i] = ld_vol_v2(addr(i));
    uint32_t pend = 0;
#pragma unroll
    for (int i = 0; i < N; i++)
        if (((want >> i) & 1u) && !tags_ok2(v[i], tag)) pend |= 1u << i;
    if (pend) {
        const unsigned long long t0 = global_ns();
        do {
            __nanosleep(40);        // idle CTAs poll for a long time (e.g. the 116 without a head during attention): stay off the L2
#pragma unroll
            for (int i = 0; i < N; i++)
                if ((pend >> i) & 1u) v[i] = ld_vol_v2(addr(i));
#pragma unroll
            for (int i = 0; i < N; i++)
                if (((pend >> i) & 1u) && tags_ok2(v[i], tag)) pend &= ~(1u << i);
            if (global_ns() - t0 > kWaitLimitNs) protocol_timeout(2);
        } while (pend);
    }
}

// Fused RMSNorm + staging (rmsnorm_kernel, gpu_kernels.h:73-105).  The reference runs 1024 threads: thread vt chains
// x[vt + 1024 i]^2, each warp is tree-summed (cub), thread 0 adds the 32 warp aggregates in order.  In trip coordinates
// vt = lane*32 + i, i.e. reference lane L of the GEMV IS virtual warp L of the norm and element i its virtual lane.  A thread
// owning slot i2 of lanes A and B over every trip holds two complete chains of each: tree level 1 is a local add, levels 2-5
// a butterfly over the 16 threads of the lane pair.  256 threads do the whole vector; x is read from L2 once.  Until the scale is known a thread parks its raw elements (and, by asynchronous copy,
// the norm weights) in the very slot it will overwrite with the result (private: no barrier, no bank conflicts).  Output: fp32
// pairs (INT4 ops) or plain fp16 (classifier, whose weights are parked behind the vector).  Ends with a named barrier.
template <bool PAIRS>
__device__ __forceinline__ void norm_take(const Op& op, const SlotMap& m, uint32_t slot0, int t, uint32_t xa, uint32_t xb, float (&sa)[2], float (&sb)[2]) {
    const float a0 = h2f_bits(xa & 0xFFFFu), a1 = h2f_bits(xa >> 16), b0 = h2f_bits(xb & 0xFFFFu), b1 = h2f_bits(xb >> 16);
    sa[0] = __fmaf_rn(a0, a0, sa[0]); sa[1] = __fmaf_rn(a1, a1, sa[1]);
    sb[0] = __fmaf_rn(b0, b0, sb[0]); sb[1] = __fmaf_rn(b1, b1, sb[1]);
    if (PAIRS) {
        sts_v2_u32(slot0 + t * kTripBytes, make_uint2(xa, xb));
    } else {
        sts_u32(slot0 + (t * 1024 + m.offA) * 2, xa); sts_u32(slot0 + (t * 1024 + m.offB) * 2, xb);
    }
    if (op.emb != nullptr && blockIdx.x == 0 && op.x_copy != nullptr) {      // layer 0: the embedding row becomes the residual stream
        *reinterpret_cast<uint32_t*>(op.x_copy + t * 1024 + m.offA) = xa;
        *reinterpret_cast<uint32_t*>(op.x_copy + t * 1024 + m.offB) = xb;
    }
}

template <bool PAIRS, int F>
__device__ __forceinline__ void stage_norm(CtxT<F>& c, const Op& op, const half* xin) {
    const int K = op.K, T = (K + 1023) >> 10;
    const uint32_t red = c.sm.bars + kRedOffset;
    const bool worker = c.ctid < kNormThreads;
    const SlotMap m = slot_map(c.ctid & 255);
    const uint32_t slot0 = PAIRS ? c.sm.xs + m.i2 * kRowBytes + m.j * 16 : c.sm.xs;      // + t*4096: the thread's slot
    const uint32_t wpark = c.sm.xs + ((K * 2 + 127) & ~127);                          // classifier: parked norm weights
    if (worker) {
        // the norm weights do not depend on the hand-over: asynchronous copies straight to their parking place
#pragma unroll 1
        for (int t = 0; t < T; t++) {
            if (t * 1024 + m.offA < K) {          // K % 64 == 0: lanes A and B are live or dead together
                cp_async4(PAIRS ? slot0 + t * kTripBytes + 8 : wpark + (t * 1024 + m.offA) * 2, op.norm_w + t * 1024 + m.offA);
                cp_async4(PAIRS ? slot0 + t * kTripBytes + 12 : wpark + (t * 1024 + m.offB) * 2, op.norm_w + t * 1024 + m.offB);
            }
        }
        float sa[2] = {0.0f, 0.0f}, sb[2] = {0.0f, 0.0f};
        if (op.xt != nullptr) {
#pragma unroll
            for (int t0 = 0; t0 < kNormMaxT; t0 += kNormChunk) {
                if (t0 < T) {
                    uint2 v[2 * kNormChunk];
                    uint32_t want = 0;
#pragma unroll
                    for (int q = 0; q < kNormChunk; q++)
                        if (t0 + q < T && (t0 + q) * 1024 + m.offA < K) want |= 3u << (2 * q);
                    const uint32_t* xt = op.xt;
                    poll_vecs(v, [&](int i) { return xt + (t0 + (i >> 1)) * 1024 + ((i & 1) ? m.offB : m.offA); }, want, c.tag_in);
#pragma unroll
                    for (int q = 0; q < kNormChunk; q++)
                        if ((want >> (2 * q)) & 1u) norm_take<PAIRS>(op, m, slot0, t0 + q, pack_duo(v[2 * q]), pack_duo(v[2 * q + 1]), sa, sb);
                }
            }
        } else {      // plain fp16 input (embedding row of layer 0, stand-alone ops): nothing to wait for
#pragma unroll 1
            for (int t = 0; t < T; t++)
                if (t * 1024 + m.offA < K)
                    norm_take<PAIRS>(op, m, slot0, t, ld_cg_u32(xin + t * 1024 + m.offA), ld_cg_u32(xin + t * 1024 + m.offB), sa, sb);
        }
        float va = sa[0] + sa[1], vb = sb[0] + sb[1];
#pragma unroll
        for (int d = 1; d < 16; d <<= 1) {
            va = va + __shfl_xor_sync(0xffffffffu, va, d);
            vb = vb + __shfl_xor_sync(0xffffffffu, vb, d);
        }
        if (m.i2 == 0) {
            const int sw = (m.j >> 2) & 1;
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(red + (2 * m.j + sw) * 4), "f"(va) : "memory");
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(red + (2 * m.j + 1 - sw) * 4), "f"(vb) : "memory");
        }
    }
    named_bar(kBarAll, c.nthreads);
    cyc_mark(c, 1);                               // warp aggregates of x^2 in shared memory
    if (worker) {
        float tot;
        {
            float r[32];
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const uint4 t = lds_v4(red + q * 16);
                r[4 * q] = __uint_as_float(t.x); r[4 * q + 1] = __uint_as_float(t.y); r[4 * q + 2] = __uint_as_float(t.z); r[4 * q + 3] = __uint_as_float(t.w);
            }
            tot = r[0];
#pragma unroll
            for (int w = 1; w < 32; w++) tot = tot + r[w];       // thread 0 of the reference adds the warp aggregates in order
        }
        tot = ((K & (K - 1)) == 0) ? __fmul_rn(tot, 1.0f / (float)K) : __fdiv_rn(tot, (float)K);   // exact either way
        tot = tot + 1e-5f;
        const float scale = __fdiv_rn(1.0f, __fsqrt_rn(tot));
        cyc_mark(c, 2);                           // scale known
        cp_async_wait_all();                      // this thread's parked norm weights
#pragma unroll 2
        for (int t = 0; t < T; t++) {
            if (t * 1024 + m.offA < K) {
                uint32_t xa, xb, ga, gb;
                if (PAIRS) {
                    const uint4 r = lds_v4(slot0 + t * kTripBytes);
                    xa = r.x; xb = r.y; ga = r.z; gb = r.w;
                } else {
                    xa = lds_u32(slot0 + (t * 1024 + m.offA) * 2); xb = lds_u32(slot0 + (t * 1024 + m.offB) * 2);
                    ga = lds_u32(wpark + (t * 1024 + m.offA) * 2); gb = lds_u32(wpark + (t * 1024 + m.offB) * 2);
                }
                const uint32_t a0 = norm_h(xa & 0xFFFFu, ga & 0xFFFFu, scale, true), a1 = norm_h(xa >> 16, ga >> 16, scale, true);
                const uint32_t b0 = norm_h(xb & 0xFFFFu, gb & 0xFFFFu, scale, true), b1 = norm_h(xb >> 16, gb >> 16, scale, true);
                if (PAIRS) {
                    sts_v4(slot0 + t * kTripBytes, h2f_bits(a0), h2f_bits(b0), h2f_bits(a1), h2f_bits(b1));
                } else {
                    sts_u32(slot0 + (t * 1024 + m.offA) * 2, a0 | (a1 << 16)); sts_u32(slot0 + (t * 1024 + m.offB) * 2, b0 | (b1 << 16));
                }
            }
        }
    }
    named_bar(kBarAll, c.nthreads);
}

// Staging for the INT4 ops: x (fp16: tagged words, plain global, or an embedding row) -> fp32 pairs in shared memory.
// The caller has already passed a named barrier: nobody reads the staging area any more.
template <int F>
__device__ void stage_x_pairs(CtxT<F>& c, const Op& op) {
    const int K = op.K;
    const half* xin = op.x;
    if (op.emb != nullptr) {
        const int token = (op.tokens == c.P->tokens) ? c.token : op.tokens[c.pos];
        xin = op.emb + (size_t)token * K;
    }
    if (op.norm_w != nullptr) {
        stage_norm<true>(c, op, xin);
        return;
    }
    if (op.emb != nullptr && blockIdx.x == 0 && op.x_copy != nullptr)
        for (int k = c.ctid; k < K; k += c.nthreads) op.x_copy[k] = xin[k];
    const int units = op.T * 256;     // slots, i2 fastest; the block size is a multiple of 16, so 16 consecutive threads always share a lane pair
    const int nth = c.nthreads;
    // first element of lane A (which = 0) or B (1) of slot u, and the slot's byte offset in the pair area
    auto elem = [](int u, int which) {
        const SlotMap m = slot_map(u & 255);
        return (u >> 8) * 1024 + (which ? m.offB : m.offA);
    };
    auto slot_off = [](int u) {
        const SlotMap m = slot_map(u & 255);
        return (u >> 8) * kTripBytes + m.i2 * kRowBytes + m.j * 16;
    };
    if (op.xt != nullptr) {
#pragma unroll 1
        for (int ub = c.ctid; ub < units; ub += kPollChunk * nth) {
            uint2 v[2 * kPollChunk];
            uint32_t want = 0;
            const uint32_t* xt = op.xt;
#pragma unroll
            for (int q = 0; q < kPollChunk; q++)
                if (ub + q * nth < units && elem(ub + q * nth, 0) < K) want |= 3u << (2 * q);      // K % 64 == 0: both lanes are live or dead together
            poll_vecs(v, [&](int i) { return xt + elem(ub + (i >> 1) * nth, i & 1); }, want, c.tag_in);
#pragma unroll
            for (int q = 0; q < kPollChunk; q++) {
                if ((want >> (2 * q)) & 1u) {
                    const int u = ub + q * nth;
                    const uint32_t a = pack_duo(v[2 * q]), b = pack_duo(v[2 * q + 1]);
                    sts_v4(c.sm.xs + slot_off(u), h2f_bits(a & 0xFFFFu), h2f_bits(b & 0xFFFFu), h2f_bits(a >> 16), h2f_bits(b >> 16));
                }
            }
        }
    } else {
#pragma unroll 1
        for (int u = c.ctid; u < units; u += nth) {
            if (elem(u, 0) < K) {
                const uint32_t a = ld_cg_u32(xin + elem(u, 0)), b = ld_cg_u32(xin + elem(u, 1));
                sts_v4(c.sm.xs + slot_off(u), h2f_bits(a & 0xFFFFu), h2f_bits(b & 0xFFFFu), h2f_bits(a >> 16), h2f_bits(b >> 16));
            }
        }
    }
    named_bar(kBarAll, c.nthreads);
}

// ------------------------------------------------------------------------------------------------
// INT4 GEMV / FFN consumer.  One trip of one warp-task: every thread advances the two lane chains of its two
// columns by 32 k each (128 weights).  acc.lo = chain of reference lane A, acc.hi = chain of lane B.
// ------------------------------------------------------------------------------------------------
struct ColMeta {
    uint32_t s16;      // fp16 scale bits of this trip's group
    float nlo, nhi;    // -(1024+z)*s and -(64+z)*s, exact
};
__device__ __forceinline__ ColMeta col_meta(uint32_t scol, uint32_t zcol, int t, int j) {
    ColMeta m;
    m.s16 = lds_u16(scol + (t * 8 + (j >> 1)) * 2);
    const uint32_t zw = lds_u32(zcol + t * 4);
    const uint32_t zp = ((zw >> ((j >> 1) * 4)) & 0xFu) | 0x6400u;             // fp16 bits of 1024 + z
    m.nlo = fhfma_lo(zp, m.s16 ^ 0x8000u, 0.0f);                               // -(1024+z)*s, exact
    m.nhi = fhfma_lo(0x6380u, m.s16, m.nlo);                                   // -(64+z)*s = 960*s + nlo, exact
    return m;
}

__device__ __forceinline__ void q4_trip2(unsigned long long& acc0, unsigned long long& acc1, uint32_t xaddr, uint32_t w0, uint32_t w1,
                                         const ColMeta& m0, const ColMeta& m1) {
    const uint4 wa0 = lds_v4(w0), wb0 = lds_v4(w0 ^ 16), wa1 = lds_v4(w1), wb1 = lds_v4(w1 ^ 16);
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
            ffma2_pk(acc0, da0[2 * e2], db0[2 * e2], xp0);
            ffma2_pk(acc1, da1[2 * e2], db1[2 * e2], xp0);
            ffma2_pk(acc0, da0[2 * e2 + 1], db0[2 * e2 + 1], xp1);
            ffma2_pk(acc1, da1[2 * e2 + 1], db1[2 * e2 + 1], xp1);
        }
    }
}

// cub::WarpReduce order over the 32 reference lanes of a column held by a half-warp: level 1 pairs lanes
// (2j, 2j+1) = this thread's two chains, levels 2..5 are a butterfly over the 16 threads.
__device__ __forceinline__ float halfwarp_total(unsigned long long acc) {
    float a0, a1;
    unpack_f2(acc, a0, a1);
    float v = a0 + a1;
    v = v + __shfl_xor_sync(0xffffffffu, v, 1);
    v = v + __shfl_xor_sync(0xffffffffu, v, 2);
    v = v + __shfl_xor_sync(0xffffffffu, v, 4);
    v = v + __shfl_xor_sync(0xffffffffu, v, 8);
    return v;
}

// first ring slot / phase of a task, and stepping to the next slot
struct RingPos {
    int slot;
    unsigned lap;
};
__device__ __forceinline__ RingPos ring_pos(const Ctx& c, unsigned q) {
    RingPos r;
    r.lap = q / (unsigned)c.sm.S;
    r.slot = (int)(q - r.lap * c.sm.S);
    return r;
}
__device__ __forceinline__ void ring_next(const Ctx& c, RingPos& r) {
    if (++r.slot == c.sm.S) { r.slot = 0; r.lap++; }
}
// Wait until chunk (slot, lap of the current epoch) has landed.  A parity wait cannot tell one fill of a slot from the
// one two fills later, and a warp may be several laps ahead of the ring (tasks are dealt round-robin), so first wait until
// the slot has been released as often as it was filled before this chunk: from then on the only phase of its full barrier
// that can still complete is this chunk's.  fill_base[slot] = fills before the current epoch.
__device__ __forceinline__ unsigned ring_need(const Ctx& c, const RingPos& r) {
    unsigned base;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(base) : "r"(c.sm.bars + kFillBaseOffset + r.slot * 4) : "memory");
    return base + r.lap;
}
__device__ __forceinline__ void ring_wait(const Ctx& c, const RingPos& r) {
    const unsigned need = ring_need(c, r);
    unsigned done;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(done) : "r"(c.sm.laps + r.slot * 4) : "memory");
    if (done < need) {
        const unsigned long long t0 = global_ns();
        do {
            asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(done) : "r"(c.sm.laps + r.slot * 4) : "memory");
            if (global_ns() - t0 > kWaitLimitNs) protocol_timeout(4);
        } while (done < need);
    }
    mbar_wait(c.sm.full(r.slot), need & 1);
}
// lane 0 of the consuming warp, after the whole warp is done with the slot
__device__ __forceinline__ void ring_release(const Ctx& c, const RingPos& r) {
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(c.sm.laps + r.slot * 4), "r"(ring_need(c, r) + 1) : "memory");
    mbar_arrive(c.sm.empty(r.slot));
}
// A new ring epoch starts when the op's slot size differs from the current one.  Called by all consumer threads between the
// op's two opening named barriers: thread i folds the fills of slot i in the finished epoch into fill_base[i].
__device__ __forceinline__ void ring_epoch(Ctx& c, int slot_bytes, int nslots) {
    if (slot_bytes == c.sm.slot_bytes) return;
    if (c.ctid < c.sm.S) {
        const unsigned n = c.qbase, S = (unsigned)c.sm.S, i = (unsigned)c.ctid;
        const unsigned fills = (n > i) ? (n - i + S - 1) / S : 0u;
        const uint32_t a = c.sm.bars + kFillBaseOffset + i * 4;
        unsigned v;
        asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
        asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(a), "r"(v + fills) : "memory");
    }
    c.qbase = 0;
    c.sm.S = nslots;
    c.sm.slot_bytes = slot_bytes;
}

template <int F>
__device__ void run_q4(CtxT<F>& c, const Op& op) {
    int t0, t1;
    cta_task_range(op, blockIdx.x, gridDim.x, t0, t1);
    // everything that does not depend on the activations first: the hand-over below is where this warp waits anyway
    const int mb = c.mcount & 1;
    const uint32_t meta = c.sm.mbuf(mb);
    const int K = op.K, T = op.T, G = q4_groups(K), zh = q4_zh(K), colb = q4_col_bytes(K);
    const int cps = op.cps, spt = op.spt, csh = (cps == 4) ? 2 : (cps == 2) ? 1 : 0;
    const bool dual = (op.kind == OP_FFN);
    const int h = c.lane >> 4, j = c.lane & 15, sw = (j >> 2) & 1;
    const int nseg = op.nseg, ncols0 = op.seg[0].ncols, ncols1 = op.seg[1].ncols;     // read once, not once per task
    RingPos rp = ring_pos(c, c.qbase + (unsigned)c.warp * spt);                        // this warp's first task
    stage_x_pairs(c, op);
    trace_mark(c, 2);                                    // activations staged
    cyc_mark(c, 3);
    if ((F & kDev) && c.tr != nullptr) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        val_mark(c, 8, producer_issued(c)); val_mark(c, 10, c.qtotal); val_mark(c, 11, smid);
    }

    // (Waiting for the first task's weights BEFORE the staging was measured and dropped: at a ring-epoch change the producer is
    // still draining, and the polls of the hand-over then start late: 1.91 -> 2.05 ms per 7B token.)
    mbar_wait(c.sm.mfull(mb), (c.mcount >> 1) & 1);      // scales / zero points of this op have landed
    cyc_mark(c, 4);
#pragma unroll 1
    for (int task = t0 + c.warp; task < t1; task += c.nwc) {
        // ---- my two columns: where their weights, scales and zeros are ----
        int seg = 0, col = 0;
        uint32_t scol0, scol1, zcol0, zcol1, w0, w1;
        // wait for every slot of the task
        {
            RingPos r = rp;
            for (int i = 0; i < spt; i++) { ring_wait(c, r); ring_next(c, r); }
        }
        if (dual) {
            const int nout = (t1 - t0) * 2, lo = (task - t0) * 2 + h;      // local output index
            col = task * 2 + h;
            scol0 = meta + lo * G * 2;                         scol1 = scol0 + nout * G * 2;
            zcol0 = meta + 2 * nout * G * 2 + lo * zh * 4;     zcol1 = zcol0 + nout * zh * 4;
            // slot: [gate cps columns][up cps columns]; output h of the task is column h % cps of slot h / cps
            RingPos r = rp;
            if (h >= cps) ring_next(c, r);                            // cps == 1: second output sits in the second slot
            w0 = c.sm.slot(r.slot) + (h & (cps - 1)) * colb;
            w1 = w0 + cps * colb;
        } else {
            const int ncol = (t1 - t0) * 4, lc = (task - t0) * 4 + 2 * h;  // local index of my first column
            col = task * 4;
            if (nseg > 1 && col >= ncols0) { col -= ncols0; seg = 1; if (nseg > 2 && col >= ncols1) { col -= ncols1; seg = 2; } }
            col += 2 * h;
            scol0 = meta + lc * G * 2;                         scol1 = scol0 + G * 2;
            zcol0 = meta + ncol * G * 2 + lc * zh * 4;         zcol1 = zcol0 + zh * 4;
            // column i (0..3) of the task is column i % cps of slot i / cps
            RingPos r = rp;
            const int i0 = 2 * h, i1 = 2 * h + 1;
            for (int s = 0; s < (i0 >> csh); s++) ring_next(c, r);
            w0 = c.sm.slot(r.slot) + (i0 & (cps - 1)) * colb;
            if ((i1 >> csh) != (i0 >> csh)) ring_next(c, r);
            w1 = c.sm.slot(r.slot) + (i1 & (cps - 1)) * colb;
        }
        if (task == t0) { trace_mark(c, 3); cyc_mark(c, 5); if ((F & kDev) && c.tr != nullptr) val_mark(c, 9, producer_issued(c)); }   // warp 0: first task's weights are in shared memory
        if (task == t0 + c.warp) warp_mark(c, 0);
        const int nth_task = (task - t0) / c.nwc;
        w0 += j * 32 + sw * 16;
        w1 += j * 32 + sw * 16;
        unsigned long long acc0 = 0ull, acc1 = 0ull;
        const int tlive = (K - j * 64 + 1023) >> 10;      // trips in which this thread's lanes hold data
#pragma unroll 1
        for (int t = 0; t < T; t++) {
            if (t < tlive && !((F & kDev) && c.P->nomath)) {
                const ColMeta m0 = col_meta(scol0, zcol0, t, j), m1 = col_meta(scol1, zcol1, t, j);
                q4_trip2(acc0, acc1, c.sm.xs + t * kTripBytes + j * 16, w0 + t * 512, w1 + t * 512, m0, m1);
            }
        }
        __syncwarp();
        if (c.lane == 0) {
            RingPos r = rp;
            for (int i = 0; i < spt; i++) { ring_release(c, r); ring_next(c, r); }
        }
        {   // ring position of this warp's next task: nwc tasks further on
            rp.slot += c.nwc * spt;
            while (rp.slot >= c.sm.S) { rp.slot -= c.sm.S; rp.lap++; }
        }
        warp_mark(c, 1 + nth_task < 6 ? 1 + nth_task : 8);
        if (task == t0) { trace_mark(c, 6); cyc_mark(c, 6); }   // warp 0: first task done
        else if (task == t0 + c.nwc) trace_mark(c, 7);    // warp 0: second task done
        // ---- epilogue ----
        const float v0 = halfwarp_total(acc0), v1 = halfwarp_total(acc1);
        if (j == 0) {
            if (dual) {                                          // gpu_kernels.h:269-273
                float val = v0;
                val = __fmul_rn(val, __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-val))));
                val = __fmul_rn(val, v1);
                const uint32_t hb = f2h_bits(val);
                if (op.seg[0].out32 != nullptr) st_tagged_maybe_all<F>(*c.P, op.seg[0].bcast != 0, op.seg[0].out32 + col, c.tag_out, hb);
                else op.seg[0].out[col] = __ushort_as_half((unsigned short)hb);
            } else {
                const Seg& sg = op.seg[seg];
                half* dst = sg.out;
                if (dst != nullptr && sg.pos_stride != 0) dst += sg.loff + (size_t)c.pos * sg.pos_stride;
                float s0 = v0, s1 = v1;
                if (op.accum) {
                    uint32_t o0, o1;
                    if (op.res_emb != nullptr) {                 // first layer: the residual is the embedding row itself
                        const half* e = op.res_emb + (size_t)((op.tokens == c.P->tokens) ? c.token : op.tokens[c.pos]) * op.res_stride + col;
                        o0 = ldg_stream_u16(e); o1 = ldg_stream_u16(e + 1);
                    } else if (sg.out32 != nullptr) {
                        o0 = ld_tagged_any(sg.out32 + col) & 0xFFFFu; o1 = ld_tagged_any(sg.out32 + col + 1) & 0xFFFFu;
                    } else {
                        o0 = ld_cg_u16(dst + col); o1 = ld_cg_u16(dst + col + 1);
                    }
                    s0 = s0 + h2f_bits(o0);
                    s1 = s1 + h2f_bits(o1);
                }
                const uint32_t h0 = f2h_bits(s0), h1 = f2h_bits(s1);
                if (sg.out32 != nullptr) st_tagged2_maybe_all<F>(*c.P, sg.bcast != 0, sg.out32 + col, c.tag_out, h0, h1);   // col is even
                if (dst != nullptr) { dst[col] = __ushort_as_half((unsigned short)h0); dst[col + 1] = __ushort_as_half((unsigned short)h1); }
            }
        }
    }
    trace_mark(c, 4);                                    // warp 0 has finished its tasks
    warp_mark(c, 7);
    c.qbase += (unsigned)(t1 - t0) * spt;
    c.qtotal += (unsigned)(t1 - t0) * spt;
    c.meta_pending = mb;
    c.mcount++;
}

// ------------------------------------------------------------------------------------------------
// fp16 classifier consumer (mat_vec_kernel, gpu_kernels.h:109-139).  Reference lane L chains
// k = (trip*32 + L)*8 + el over trips of 256 k; one thread is one reference lane of four rows.
// ------------------------------------------------------------------------------------------------
template <int F>
__device__ void run_cls(CtxT<F>& c, const Op& op) {
    const int n = op.K, T = op.T, lane = c.lane, cps = op.cps, spt = op.spt;
    // ---- stage x as fp16 (through the fused RMSNorm); the caller has passed a named barrier ----
    if (op.norm_w != nullptr) {
        stage_norm<false>(c, op, op.x);
    } else {
#pragma unroll 1
        for (int u = c.ctid; u * 8 < n; u += c.nthreads) {
            const uint4 xv = (op.xt != nullptr) ? poll8(op.xt + u * 8, c.tag_in) : ld_cg_v4(op.x + u * 8);
            sts_v4_u32(c.sm.xs + u * 16, xv);
        }
        named_bar(kBarAll, c.nthreads);
    }

    int t0, t1;
    cta_task_range(op, blockIdx.x, gridDim.x, t0, t1);
#pragma unroll 1
    for (int task = t0 + c.warp; task < t1; task += c.nwc) {
        int rows = op.seg[0].ncols - task * op.rpt;
        if (rows > op.rpt) rows = op.rpt;
        const RingPos rp = ring_pos(c, c.qbase + (unsigned)(task - t0) * spt);
        uint32_t wrow[4];
        {
            RingPos r = rp;
            for (int i = 0; i < spt; i++) {
                ring_wait(c, r);
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (q / cps == i) wrow[q] = c.sm.slot(r.slot) + (q % cps) * n * 2 + lane * 16;
                ring_next(c, r);
            }
        }
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        for (int t = 0; t < T; t++) {
            const int jx = (t * 32 + lane) * 8;
            if (jx < n && !((F & kDev) && c.P->nomath)) {
                const uint4 xv = lds_v4(c.sm.xs + jx * 2);
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    if (r < rows) {
                        const uint4 w = lds_v4(wrow[r] + t * 512);
                        float a = acc[r];
                        a = fhfma_ll(w.x, xv.x, a); a = fhfma_hh(w.x, xv.x, a);
                        a = fhfma_ll(w.y, xv.y, a); a = fhfma_hh(w.y, xv.y, a);
                        a = fhfma_ll(w.z, xv.z, a); a = fhfma_hh(w.z, xv.z, a);
                        a = fhfma_ll(w.w, xv.w, a); a = fhfma_hh(w.w, xv.w, a);
                        acc[r] = a;
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) {
            RingPos r = rp;
            for (int i = 0; i < spt; i++) { ring_release(c, r); ring_next(c, r); }
        }
        const float tot = warp_tree_sum4(acc[0], acc[1], acc[2], acc[3], lane);   // lanes 0,1,2,3 hold rows 0,2,1,3
        if (lane < 4) {
            const int r = 2 * (lane & 1) + ((lane >> 1) & 1);
            if (r < rows) op.seg[0].out[task * op.rpt + r] = __float2half_rn(__fmul_rn(tot, op.alpha));
        }
    }
    c.qbase += (unsigned)(t1 - t0) * spt;
    c.qtotal += (unsigned)(t1 - t0) * spt;
}

// ------------------------------------------------------------------------------------------------
// Attention for one head per CTA: [RoPE on q and on the new k row] -> QK^T -> softmax -> PV
// (RoPERotation_kernel :332-355, mat_vec_kernel_simple :142-168, softmax_kernel :357-401,
//  vec_mat_kernel :279-329).  The reference's 1024-thread reductions are replayed with virtual threads.
//
// K and V rows of earlier positions stream through shared memory in tiles of kAttnTile rows, four tile buffers deep, with
// asynchronous 16-byte copies (cp.async: no registers, three tiles in flight while one is consumed, one named barrier per
// tile); the first three K tiles are requested BEFORE the hand-over that precedes the op, since rows t < pos were written by
// earlier launches.
// Scratch layout (floats unless noted): qs[hs] | krow[hs] | vrow[hs] | att[max_seq] | 4 tile buffers (fp16; the first two later hold `part`)
// ------------------------------------------------------------------------------------------------
constexpr int kAttnTile = 32;
constexpr int kAttnAhead = 3;      // tiles requested ahead of the one being consumed (kAttnAhead + 1 buffers)
__host__ __device__ __forceinline__ int attn_buf_bytes(int hs) {      // two tile buffers, or the 32 x hs fp32 partial sums of the PV tree
    const int tile = 2 * kAttnTile * hs * 2, part = 32 * hs * 4;
    return ((tile > part ? tile : part) + 127) & ~127;
}
__host__ __device__ __forceinline__ int attn_fixed_bytes(int hs, int max_seq) { return ((3 * hs + ((max_seq + 3) & ~3)) * 4 + 127) & ~127; }
__device__ __forceinline__ uint32_t attn_tile_buf(const Ctx& c, int hs, int max_seq, int ti) {
    return c.sm.xs + attn_fixed_bytes(hs, max_seq) + (uint32_t)(ti & kAttnAhead) * (kAttnTile * hs * 2);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Request tile ti (rows [ti*kAttnTile, ...) of one head: hs halfs each, kv_stride apart) into its buffer and close a copy group.
// A tile past the end is an empty group, so that every caller's group count stays the same.
__device__ __forceinline__ void attn_tile_async(const Ctx& c, const half* base, int kv_stride, int hs, int max_seq, int ti, int rows_total) {
    const int vsh = (hs == 128) ? 4 : (hs == 64) ? 3 : 2;  // log2(16-byte vectors per row)
    const int row0 = ti * kAttnTile;
    int nrows = rows_total - row0;
    if (nrows > kAttnTile) nrows = kAttnTile;
    if (nrows > 0) {
        const int total = nrows << vsh;
        const uint32_t dst = attn_tile_buf(c, hs, max_seq, ti);
        const half* src = base + (size_t)row0 * kv_stride;
        for (int idx = c.ctid; idx < total; idx += c.nthreads)
            cp_async16(dst + idx * 16, src + (size_t)(idx >> vsh) * kv_stride + (idx & ((1 << vsh) - 1)) * 8);
    }
    cp_async_commit();
}
__device__ __forceinline__ int attn_parts(const Op& op, int pos, int grid);
__device__ __forceinline__ void attn_rows_async(const Ctx& c, uint32_t dst, const half* src, int stride_elems, int row_bytes, int nrows);
// SPLIT: the kernel instance that carries the several-CTAs-per-head attention (interp_kernel<true>, launched by the host once the
// position passes kAttnSplitFrom).  The short-context instance holds none of that code: merely compiling it in cost 8 % of the
// 7B step (1.88 -> 2.03 ms; register allocation of the staging and task loops), so it lives in its own instantiation.
template <int F>
__device__ void attn_prefetch(const CtxT<F>& c, const Op& op) {
    const int hs = op.head_size;
    const int S = (F & kSplit) ? attn_parts(op, c.pos, (int)gridDim.x) : 1;
    if (S > 1) {                                    // split mode: my K tiles are part_id, part_id + S, ... of head blockIdx.x / S
        const int h = blockIdx.x / S, part_id = blockIdx.x - h * S;
        if (h >= op.n_heads) return;
        const half* kb = op.kcache + (size_t)(h / op.kv_mul) * hs;
        const int ntiles = (c.pos + kAttnTile - 1) / kAttnTile;
        const uint32_t bufs = c.sm.xs + attn_fixed_bytes(hs, op.max_seq);
        for (int m = 0; m < kAttnAhead; m++) {
            const int ti = part_id + m * S;
            int nrows = (ti < ntiles) ? c.pos - ti * kAttnTile : 0;
            if (nrows > kAttnTile) nrows = kAttnTile;
            attn_rows_async(c, bufs + (uint32_t)m * (kAttnTile * hs * 2), kb + (size_t)ti * kAttnTile * op.kv_stride, op.kv_stride, hs * 2, nrows);
        }
        return;
    }
    const int h = blockIdx.x;
    if (h >= op.n_heads) return;
    const half* kb = op.kcache + (size_t)(h / op.kv_mul) * hs;
    for (int ti = 0; ti < kAttnAhead; ti++) attn_tile_async(c, kb, op.kv_stride, hs, op.max_seq, ti, c.pos);
}

template <int NSER, int F>
__device__ void run_attn_t(CtxT<F>& c, const Op& op, bool prefetched) {
    constexpr int NS = NSER;                               // hs / 32
    const int hs = op.head_size, nt = c.nthreads, tid = c.ctid, lane = c.lane, warp = c.warp;
    float* qs = reinterpret_cast<float*>(c.scratch);       // hs
    float* krow = qs + hs;                                  // hs: rotated k row of this step
    float* vrow = krow + hs;                                // hs: v row of this step
    float* att = vrow + hs;                                 // max_seq
    float* part = reinterpret_cast<float*>(c.scratch + attn_fixed_bytes(hs, op.max_seq));   // 32 * hs, aliases the first two tile buffers
    float* red = c.red;
    const int pos = c.pos, size = pos + 1;
#pragma unroll 1
    for (int h = blockIdx.x; h < op.n_heads; h += gridDim.x) {
        const int kvh = h / op.kv_mul;
        half* kbase = op.kcache + (size_t)kvh * hs;
        const half* vbase = op.vcache + (size_t)kvh * hs;
        const int ntiles = (pos + kAttnTile - 1) / kAttnTile;
        if (!(prefetched && h == (int)blockIdx.x))          // the first K tiles of the CTA's first head are already on their way
            for (int ti = 0; ti < kAttnAhead; ti++) attn_tile_async(c, kbase, op.kv_stride, hs, op.max_seq, ti, pos);
        // ---- q, the new k row (rotated here when a table is given) and the new v row -> shared memory ----
        if (op.rope_tab != nullptr) {
            for (int i = tid; i < hs / 2; i += nt) {
                const float2 cs = op.rope_tab[(size_t)pos * (hs / 2) + i];
                half* q = op.q + (size_t)h * hs;
                const half* kr = op.kraw + (size_t)kvh * hs;
                uint32_t q0b, q1b, k0b, k1b;
                if (op.qt != nullptr) {      // fused path: q and the un-rotated k row arrive as tagged words from the q|k|v op
                    const uint32_t* pq = op.qt + (size_t)h * hs + i;
                    const uint32_t* pk = op.krawt + (size_t)kvh * hs + i;
                    const uint32_t tag = c.tag_in;
                    const unsigned long long t0 = global_ns();
                    for (;;) {               // the four words are fetched together: one round trip per attempt, not four
                        q0b = ld_tagged_any(pq); q1b = ld_tagged_any(pq + hs / 2); k0b = ld_tagged_any(pk); k1b = ld_tagged_any(pk + hs / 2);
                        if (((q0b >> 16) == tag) & ((q1b >> 16) == tag) & ((k0b >> 16) == tag) & ((k1b >> 16) == tag)) break;
                        if (global_ns() - t0 > kWaitLimitNs) protocol_timeout(2);
                    }
                    q0b &= 0xFFFFu; q1b &= 0xFFFFu; k0b &= 0xFFFFu; k1b &= 0xFFFFu;
                } else {
                    q0b = ld_cg_u16(q + i); q1b = ld_cg_u16(q + i + hs / 2); k0b = ld_cg_u16(kr + i); k1b = ld_cg_u16(kr + i + hs / 2);
                }
                const float q0 = h2f_bits(q0b), q1 = h2f_bits(q1b), k0 = h2f_bits(k0b), k1 = h2f_bits(k1b);
                const half o0 = __float2half_rn(__fmaf_rn(q0, cs.x, -__fmul_rn(q1, cs.y)));
                const half o1 = __float2half_rn(__fmaf_rn(q1, cs.x, __fmul_rn(q0, cs.y)));
                if (op.qt == nullptr) { q[i] = o0; q[i + hs / 2] = o1; }     // the reference rotates q in place; nobody reads it in the fused path
                qs[i] = __half2float(o0); qs[i + hs / 2] = __half2float(o1);
                const half r0 = __float2half_rn(__fmaf_rn(k0, cs.x, -__fmul_rn(k1, cs.y)));
                const half r1 = __float2half_rn(__fmaf_rn(k0, cs.y, __fmul_rn(k1, cs.x)));
                krow[i] = __half2float(r0); krow[i + hs / 2] = __half2float(r1);
                if (h == kvh * op.kv_mul) {   // one head per kv group stores the rotated row into the cache
                    half* kd = kbase + (size_t)pos * op.kv_stride;
                    kd[i] = r0; kd[i + hs / 2] = r1;
                }
            }
            for (int i = tid - 64; i >= 0 && i < hs; i += nt)
                vrow[i] = h2f_bits(op.vrawt != nullptr ? poll1(op.vrawt + (size_t)kvh * hs + i, c.tag_in) : ld_cg_u16(vbase + (size_t)pos * op.kv_stride + i));
        } else {
            for (int i = tid; i < hs; i += nt) {
                qs[i] = h2f_bits(ld_cg_u16(op.q + (size_t)h * hs + i));
                krow[i] = h2f_bits(ld_cg_u16(kbase + (size_t)pos * op.kv_stride + i));
                vrow[i] = h2f_bits(ld_cg_u16(vbase + (size_t)pos * op.kv_stride + i));
            }
        }
        if (tid < 64 && nt <= 64)
            for (int i = tid; i < hs; i += nt)
                vrow[i] = h2f_bits(op.vrawt != nullptr ? poll1(op.vrawt + (size_t)kvh * hs + i, c.tag_in) : ld_cg_u16(vbase + (size_t)pos * op.kv_stride + i));
        // ---- scores (lane chain over j = 32 i + lane, gpu_kernels.h:154-159), K tile by tile ----
        // Copy groups: kAttnAhead are open when the loop starts and every iteration closes one more (tile ti + kAttnAhead, or an
        // empty one), so "at most kAttnAhead - 1 groups pending" always means that this thread's part of tile ti has landed.
#pragma unroll 1
        for (int ti = 0; ti < ntiles; ti++) {
            const int tile0 = ti * kAttnTile, nrows = (pos - tile0 < kAttnTile) ? pos - tile0 : kAttnTile;
            cp_async_wait<kAttnAhead - 1>();
            named_bar(kBarAll, nt);                        // tile ti is complete for everyone (first time round also qs / krow / vrow); tile ti-1 is no longer read
            attn_tile_async(c, kbase, op.kv_stride, hs, op.max_seq, ti + kAttnAhead, pos);     // into the buffer tile ti-1 just left
            const uint32_t kbuf = attn_tile_buf(c, hs, op.max_seq, ti);
            // four rows per warp and pass: their lane chains are independent, and one 6-shuffle transposing tree (the cub
            // association per row, see warp_tree_sum4) replaces four 5-shuffle trees
#pragma unroll 1
            for (int r0 = warp * 4; r0 < nrows; r0 += c.nwc * 4) {
                float s4[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint32_t row = kbuf + (uint32_t)min(r0 + q, nrows - 1) * hs * 2;
                    float sum = 0.0f;
#pragma unroll
                    for (int i = 0; i < NS; i++) sum = __fmaf_rn(h2f_bits(lds_u16(row + (i * 32 + lane) * 2)), qs[i * 32 + lane], sum);
                    s4[q] = sum;
                }
                const float tot = __fmul_rn(warp_tree_sum4(s4[0], s4[1], s4[2], s4[3], lane), op.att_alpha);
                const int rq = r0 + 2 * (lane & 1) + ((lane >> 1) & 1);      // lanes 0,1,2,3 hold rows r0 + 0,2,1,3
                if (lane < 4 && rq < nrows) att[tile0 + rq] = __half2float(__float2half_rn(tot));
            }
        }
        named_bar(kBarAll, nt);                            // all tile reads done; krow visible even when no tile loop ran (pos == 0)
        if (warp == 0) {                                   // the row of this step
            float sum = 0.0f;
#pragma unroll
            for (int i = 0; i < NS; i++) sum = __fmaf_rn(krow[i * 32 + lane], qs[i * 32 + lane], sum);
            sum = warp_tree_sum(sum);
            sum = __fmul_rn(sum, op.att_alpha);
            if (lane == 0) att[pos] = __half2float(__float2half_rn(sum));
        }
        // the first V tiles travel while the softmax runs: K is dead from here on (every warp is past the barrier above)
        cp_async_wait<0>();
        for (int ti = 0; ti < kAttnAhead; ti++) attn_tile_async(c, vbase, op.kv_stride, hs, op.max_seq, ti, pos);
        named_bar(kBarAll, nt);
        trace_mark(c, 3);
        // ---- softmax (idle reference threads seed the max with 0, gpu_kernels.h:374).  Beyond 8192 positions the reference
        // switches to softmax_kernel_no_smem (llama2_q4.cu:276-279, gpu_kernels.h:403-446), which keeps exp() in the fp16 score
        // buffer: the sum still adds the unrounded values, but the quotient is formed from the fp16-rounded one. ----
        const bool exp16 = size > op.exp16_from;
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
                att[i] = exp16 ? __half2float(__float2half_rn(e)) : e;
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
        trace_mark(c, 4);
        // ---- PV: reference lane tx chains t = 32 e + tx (e ascending); then the cub tree over tx.  A warp keeps the
        // chains of up to four tx (tx = warp + k * nwc); a lane owns hs/32 consecutive outputs. ----
        float a[4][NS];
#pragma unroll
        for (int k = 0; k < 4; k++)
#pragma unroll
            for (int q = 0; q < NS; q++) a[k][q] = 0.0f;
#pragma unroll 1
        for (int ti = 0; ti < ntiles; ti++) {
            const int tile0 = ti * kAttnTile, nrows = (pos - tile0 < kAttnTile) ? pos - tile0 : kAttnTile;
            cp_async_wait<kAttnAhead - 1>();
            named_bar(kBarAll, nt);
            attn_tile_async(c, vbase, op.kv_stride, hs, op.max_seq, ti + kAttnAhead, pos);
            const uint32_t buf = attn_tile_buf(c, hs, op.max_seq, ti);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int tx = warp + k * c.nwc;
                if (tx < nrows) {                              // kAttnTile == 32: row tx of the tile is position t = tile0 + tx, t % 32 == tx
                    const float pt = att[tile0 + tx];
                    const uint32_t row = buf + (uint32_t)tx * hs * 2 + lane * NS * 2;
#pragma unroll
                    for (int q = 0; q < NS; q++) a[k][q] = __fmaf_rn(h2f_bits(lds_u16(row + q * 2)), pt, a[k][q]);
                }
            }
        }
        cp_async_wait<0>();
        named_bar(kBarAll, nt);                            // tiles are dead: `part` may overwrite the tile buffers
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int tx = warp + k * c.nwc;
            if (tx < 32) {
                if ((pos & 31) == tx) {                    // the row of this step closes its chain
                    const float pt = att[pos];
#pragma unroll
                    for (int q = 0; q < NS; q++) a[k][q] = __fmaf_rn(vrow[lane * NS + q], pt, a[k][q]);
                }
#pragma unroll
                for (int q = 0; q < NS; q++) part[tx * hs + lane * NS + q] = a[k][q];
            }
        }
        named_bar(kBarAll, nt);
        trace_mark(c, 6);
        for (int i = tid; i < hs; i += nt) {
            float v[32];
#pragma unroll
            for (int w = 0; w < 32; w++) v[w] = part[w * hs + i];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
#pragma unroll
                for (int w = 0; w < 32; w += 2 * o) v[w] = v[w] + v[w + o];
            if (op.attn_out32 != nullptr) st_tagged_maybe_all<F>(*c.P, op.attn_bcast != 0, op.attn_out32 + (size_t)h * hs + i, c.tag_out, f2h_bits(v[0]));
            else op.attn_out[(size_t)h * hs + i] = __float2half_rn(v[0]);
        }
        named_bar(kBarAll, nt);
        trace_mark(c, 7);
    }
}

// ------------------------------------------------------------------------------------------------
// Long contexts: S = 2 or 4 CTAs per head ("parts").  One CTA per head keeps 24 KB of K / V tiles in flight, which bounds a
// 1024-position head at ~35 us; S parts stream S times as much.
//   scores   part c takes the K tiles ti = c, c + S, ... (and part 0 the row of this step), rounds each score to fp16 like
//            the reference and writes it to att_sc[h][t]; a release store of Ctx::op_seq to att_flags[h][c] publishes them
//   softmax  every part waits for all S flags, reads ALL scores of the head and runs the same softmax (identical bits)
//   PV       part c owns the output dimensions [c * hs/S, (c+1) * hs/S): it streams that slice of every V row (S * 32 rows
//            per tile buffer) and keeps the reference's chains (lane tx over t = 32 e + tx, e ascending) and its 32-way tree
// Everything a part computes is a chain or tree of the reference in the reference's order: results are bit-identical to
// the one-CTA path.
// ------------------------------------------------------------------------------------------------
constexpr int kAttnSplitFrom = 384;     // positions from which the split pays for its extra exchange (~2 us)
__device__ __forceinline__ int attn_parts(const Op& op, int pos, int grid) {
    if (pos < kAttnSplitFrom || op.att_sc == nullptr || (op.rope_tab != nullptr && op.qt == nullptr)) return 1;
    int S = op.att_split;
    while (S > 1 && (op.n_heads * S > grid || (op.head_size / 32) % S)) S >>= 1;
    return S;
}
// rows [row0, row0 + nrows) x `row_bytes` (a multiple of 16) of a strided global matrix -> contiguous shared memory; closes a copy group
__device__ __forceinline__ void attn_rows_async(const Ctx& c, uint32_t dst, const half* src, int stride_elems, int row_bytes, int nrows) {
    if (nrows > 0) {
        const int vpr = row_bytes >> 4, total = nrows * vpr;
        for (int idx = c.ctid; idx < total; idx += c.nthreads) {
            const int r = idx / vpr, v = idx - r * vpr;
            cp_async16(dst + r * row_bytes + v * 16, src + (size_t)r * stride_elems + v * 8);
        }
    }
    cp_async_commit();
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <int NSER, int S, int F>
__device__ void run_attn_split(CtxT<F>& c, const Op& op, bool prefetched) {
    constexpr int NS = NSER;                               // hs / 32
    constexpr int NSP = NS / S;                            // output dimensions per lane in the PV phase
    const int hs = op.head_size, nt = c.nthreads, tid = c.ctid, lane = c.lane, warp = c.warp;
    const int h = blockIdx.x / S, part_id = blockIdx.x - h * S;
    if (h >= op.n_heads) return;
    float* qs = reinterpret_cast<float*>(c.scratch);
    float* krow = qs + hs;
    float* vrow = krow + hs;
    float* att = vrow + hs;
    float* part = reinterpret_cast<float*>(c.scratch + attn_fixed_bytes(hs, op.max_seq));
    float* red = c.red;
    const int pos = c.pos, size = pos + 1;
    const int kvh = h / op.kv_mul;
    half* kbase = op.kcache + (size_t)kvh * hs;
    const half* vbase = op.vcache + (size_t)kvh * hs;
    const int ntiles = (pos + kAttnTile - 1) / kAttnTile;
    const int mtiles = (ntiles > part_id) ? (ntiles - part_id + S - 1) / S : 0;      // my K tiles: part_id, part_id + S, ...
    const uint32_t bufs = c.sm.xs + attn_fixed_bytes(hs, op.max_seq);
    const int tile_bytes = kAttnTile * hs * 2;
    auto k_tile_async = [&](int m) {
        const int ti = part_id + m * S;
        int nrows = (m < mtiles) ? pos - ti * kAttnTile : 0;
        if (nrows > kAttnTile) nrows = kAttnTile;
        attn_rows_async(c, bufs + (uint32_t)(m & kAttnAhead) * tile_bytes, kbase + (size_t)ti * kAttnTile * op.kv_stride, op.kv_stride, hs * 2, nrows);
    };
    if (!prefetched)
        for (int m = 0; m < kAttnAhead; m++) k_tile_async(m);
    // ---- q, the new k row (rotated) and the new v row -> shared memory (every part; part 0 of the first head of a kv group stores k) ----
    if (op.rope_tab != nullptr) {
        for (int i = tid; i < hs / 2; i += nt) {
            const float2 cs = op.rope_tab[(size_t)pos * (hs / 2) + i];
            half* q = op.q + (size_t)h * hs;
            const half* kr = op.kraw + (size_t)kvh * hs;
            uint32_t q0b, q1b, k0b, k1b;
            if (op.qt != nullptr) {
                const uint32_t* pq = op.qt + (size_t)h * hs + i;
                const uint32_t* pk = op.krawt + (size_t)kvh * hs + i;
                const uint32_t tag = c.tag_in;
                const unsigned long long t0 = global_ns();
                for (;;) {
                    q0b = ld_tagged_any(pq); q1b = ld_tagged_any(pq + hs / 2); k0b = ld_tagged_any(pk); k1b = ld_tagged_any(pk + hs / 2);
                    if (((q0b >> 16) == tag) & ((q1b >> 16) == tag) & ((k0b >> 16) == tag) & ((k1b >> 16) == tag)) break;
                    if (global_ns() - t0 > kWaitLimitNs) protocol_timeout(2);
                }
                q0b &= 0xFFFFu; q1b &= 0xFFFFu; k0b &= 0xFFFFu; k1b &= 0xFFFFu;
            } else {
                q0b = ld_cg_u16(q + i); q1b = ld_cg_u16(q + i + hs / 2); k0b = ld_cg_u16(kr + i); k1b = ld_cg_u16(kr + i + hs / 2);
            }
            const float q0 = h2f_bits(q0b), q1 = h2f_bits(q1b), k0 = h2f_bits(k0b), k1 = h2f_bits(k1b);
            const half o0 = __float2half_rn(__fmaf_rn(q0, cs.x, -__fmul_rn(q1, cs.y)));
            const half o1 = __float2half_rn(__fmaf_rn(q1, cs.x, __fmul_rn(q0, cs.y)));
            (void)q;      // the in-place rotation of q (operator-API semantics) is left to the one-CTA path: other parts still read q here
            qs[i] = __half2float(o0); qs[i + hs / 2] = __half2float(o1);
            const half r0 = __float2half_rn(__fmaf_rn(k0, cs.x, -__fmul_rn(k1, cs.y)));
            const half r1 = __float2half_rn(__fmaf_rn(k0, cs.y, __fmul_rn(k1, cs.x)));
            krow[i] = __half2float(r0); krow[i + hs / 2] = __half2float(r1);
            if (h == kvh * op.kv_mul && part_id == 0) {
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
    for (int i = tid; i < hs; i += nt)
        vrow[i] = h2f_bits(op.vrawt != nullptr ? poll1(op.vrawt + (size_t)kvh * hs + i, c.tag_in) : ld_cg_u16(vbase + (size_t)pos * op.kv_stride + i));
    // ---- my share of the scores ----
    uint16_t* sc = op.att_sc + (size_t)h * op.att_sc_stride;
#pragma unroll 1
    for (int m = 0; m < mtiles; m++) {
        const int tile0 = (part_id + m * S) * kAttnTile, nrows = (pos - tile0 < kAttnTile) ? pos - tile0 : kAttnTile;
        cp_async_wait<kAttnAhead - 1>();
        named_bar(kBarAll, nt);
        k_tile_async(m + kAttnAhead);
        const uint32_t kbuf = bufs + (uint32_t)(m & kAttnAhead) * tile_bytes;
#pragma unroll 1
        for (int r0 = warp * 4; r0 < nrows; r0 += c.nwc * 4) {
            float s4[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const uint32_t row = kbuf + (uint32_t)min(r0 + q, nrows - 1) * hs * 2;
                float sum = 0.0f;
#pragma unroll
                for (int i = 0; i < NS; i++) sum = __fmaf_rn(h2f_bits(lds_u16(row + (i * 32 + lane) * 2)), qs[i * 32 + lane], sum);
                s4[q] = sum;
            }
            const float tot = __fmul_rn(warp_tree_sum4(s4[0], s4[1], s4[2], s4[3], lane), op.att_alpha);
            const int rq = r0 + 2 * (lane & 1) + ((lane >> 1) & 1);
            if (lane < 4 && rq < nrows) sc[tile0 + rq] = (uint16_t)f2h_bits(tot);
        }
    }
    named_bar(kBarAll, nt);                                // qs / krow visible even when this part had no tile; all tile reads done
    if (part_id == 0 && warp == 0) {                       // the row of this step
        float sum = 0.0f;
#pragma unroll
        for (int i = 0; i < NS; i++) sum = __fmaf_rn(krow[i * 32 + lane], qs[i * 32 + lane], sum);
        sum = warp_tree_sum(sum);
        sum = __fmul_rn(sum, op.att_alpha);
        if (lane == 0) sc[pos] = (uint16_t)f2h_bits(sum);
    }
    // the first V slices travel while the scores are exchanged: K is dead from here on
    cp_async_wait<0>();
    const int dsub = hs / S, vrow_bytes = dsub * 2, rows_per_buf = kAttnTile * S;      // S * 32 rows x (hs / S) dims = one tile buffer
    const int nst = (pos + rows_per_buf - 1) / rows_per_buf;
    auto v_tile_async = [&](int sti) {
        int nrows = (sti < nst) ? pos - sti * rows_per_buf : 0;
        if (nrows > rows_per_buf) nrows = rows_per_buf;
        attn_rows_async(c, bufs + (uint32_t)(sti & kAttnAhead) * tile_bytes, vbase + (size_t)sti * rows_per_buf * op.kv_stride + part_id * dsub, op.kv_stride,
                        vrow_bytes, nrows);
    };
    named_bar(kBarAll, nt);                                // every thread's score stores precede the release below; K buffers are free
    for (int sti = 0; sti < kAttnAhead; sti++) v_tile_async(sti);
    if (tid == 0) st_release_u32(op.att_flags + h * 4 + part_id, c.op_seq);
    if (tid < S) {
        const unsigned* f = op.att_flags + h * 4 + tid;
        if (ld_acquire_u32(f) != c.op_seq) {
            const unsigned long long t0 = global_ns();
            while (ld_acquire_u32(f) != c.op_seq)
                if (global_ns() - t0 > kWaitLimitNs) protocol_timeout(5);
        }
    }
    named_bar(kBarAll, nt);
    trace_mark(c, 3);
    for (int i = tid; i < size; i += nt) att[i] = h2f_bits(ld_cg_u16(sc + i));
    named_bar(kBarAll, nt);
    // ---- softmax: the same code path as the one-CTA kernel, run by every part ----
    const bool exp16 = size > op.exp16_from;
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
            att[i] = exp16 ? __half2float(__float2half_rn(e)) : e;
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
        if (op.att_out != nullptr && part_id == 0) op.att_out[(size_t)h * size + i] = pr;
    }
    named_bar(kBarAll, nt);
    trace_mark(c, 4);
    // ---- PV over my output dimensions: lane owns NS / S consecutive ones; chains as in the one-CTA kernel ----
    float a[4][NSP];
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
        for (int q = 0; q < NSP; q++) a[k][q] = 0.0f;
#pragma unroll 1
    for (int sti = 0; sti < nst; sti++) {
        const int tile0 = sti * rows_per_buf, nrows = (pos - tile0 < rows_per_buf) ? pos - tile0 : rows_per_buf;
        cp_async_wait<kAttnAhead - 1>();
        named_bar(kBarAll, nt);
        v_tile_async(sti + kAttnAhead);
        const uint32_t buf = bufs + (uint32_t)(sti & kAttnAhead) * tile_bytes;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int tx = warp + k * c.nwc;
            if (tx < 32) {
#pragma unroll
                for (int es = 0; es < S; es++) {
                    const int r = es * 32 + tx;
                    if (r < nrows) {
                        const float pt = att[tile0 + r];
                        const uint32_t row = buf + (uint32_t)r * vrow_bytes + lane * NSP * 2;
#pragma unroll
                        for (int q = 0; q < NSP; q++) a[k][q] = __fmaf_rn(h2f_bits(lds_u16(row + q * 2)), pt, a[k][q]);
                    }
                }
            }
        }
    }
    cp_async_wait<0>();
    named_bar(kBarAll, nt);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int tx = warp + k * c.nwc;
        if (tx < 32) {
            if ((pos & 31) == tx) {
                const float pt = att[pos];
#pragma unroll
                for (int q = 0; q < NSP; q++) a[k][q] = __fmaf_rn(vrow[part_id * dsub + lane * NSP + q], pt, a[k][q]);
            }
#pragma unroll
            for (int q = 0; q < NSP; q++) part[tx * dsub + lane * NSP + q] = a[k][q];
        }
    }
    named_bar(kBarAll, nt);
    trace_mark(c, 6);
    for (int i = tid; i < dsub; i += nt) {
        float v[32];
#pragma unroll
        for (int w = 0; w < 32; w++) v[w] = part[w * dsub + i];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
#pragma unroll
            for (int w = 0; w < 32; w += 2 * o) v[w] = v[w] + v[w + o];
        const size_t e = (size_t)h * hs + part_id * dsub + i;
        if (op.attn_out32 != nullptr) st_tagged_maybe_all<F>(*c.P, op.attn_bcast != 0, op.attn_out32 + e, c.tag_out, f2h_bits(v[0]));
        else op.attn_out[e] = __float2half_rn(v[0]);
    }
    named_bar(kBarAll, nt);
    trace_mark(c, 7);
}

template <int F>
__device__ void run_attn(CtxT<F>& c, const Op& op, bool prefetched) {
    constexpr bool SPLIT = (F & kSplit) != 0;
    const int S = SPLIT ? attn_parts(op, c.pos, (int)gridDim.x) : 1;
    if (SPLIT && S > 1) {      // attn_parts admits S > 1 only when (hs / 32) % S == 0
        if (op.head_size == 128 && S == 4) run_attn_split<4, 4>(c, op, prefetched);
        else if (op.head_size == 128) run_attn_split<4, 2>(c, op, prefetched);
        else run_attn_split<2, 2>(c, op, prefetched);
        return;
    }
    switch (op.head_size) {            // the host accepts only these head sizes for the persistent kernel
        case 128: run_attn_t<4>(c, op, prefetched); break;
        case 64: run_attn_t<2>(c, op, prefetched); break;
        default: run_attn_t<1>(c, op, prefetched); break;
    }
}

// ------------------------------------------------------------------------------------------------
// Greedy sampler on CTA 0 (argmax_kernel, gpu_kernels.h:448-493).  Equal maxima: lowest index.
// ------------------------------------------------------------------------------------------------
template <int F>
__device__ void run_argmax(CtxT<F>& c, const Op& op, int write_token) {
    if (blockIdx.x != 0) return;
    float* smax = c.red;
    int* sidx = reinterpret_cast<int*>(c.red + 32);
    float max_val = -INFINITY;
    int max_pos = 0x7fffffff;
    // 8 logits per 16-byte load, four loads in flight per thread; ascending index order per thread keeps "first maximum"
    const int nvec = op.vocab >> 3;
#pragma unroll 1
    for (int vb = c.ctid; vb < nvec; vb += 4 * c.nthreads) {
        uint4 lv[4];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int vi = vb + r * c.nthreads;
            lv[r] = (vi < nvec) ? ld_cg_v4(op.logits + (size_t)vi * 8) : make_uint4(0xFC00FC00u, 0xFC00FC00u, 0xFC00FC00u, 0xFC00FC00u);
        }
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int vi = vb + r * c.nthreads;
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const uint32_t word = word_of(lv[r], e >> 1);
                const float v = h2f_bits((e & 1) ? (word >> 16) : (word & 0xFFFFu));
                if (vi < nvec && v > max_val) { max_val = v; max_pos = vi * 8 + e; }
            }
        }
    }
    for (int i = (nvec << 3) + c.ctid; i < op.vocab; i += c.nthreads) {
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
        if ((F & kTP) && op.cand != nullptr && c.P->world > 1) {
            // tensor parallel: every rank publishes the best of its vocabulary slice to all ranks, then picks the global
            // winner itself (largest value, lowest index on ties): all ranks write the same token without a host round trip
            const int rank = c.P->rank, world = c.P->world;
            const uint32_t gidx = (uint32_t)(max_pos + op.vocab0);      // a tagged word carries 16 bits: the index travels as two of them
            st_tagged_all(*c.P, op.cand + 3 * rank, c.tag_out, f2h_bits(max_val));
            st_tagged_all(*c.P, op.cand + 3 * rank + 1, c.tag_out, gidx & 0xFFFFu);
            st_tagged_all(*c.P, op.cand + 3 * rank + 2, c.tag_out, gidx >> 16);
            max_val = -INFINITY; max_pos = 0x7fffffff;
            for (int r = 0; r < world; r++) {
                const float v = h2f_bits(poll1(op.cand + 3 * r, c.tag_out));
                const int idx = (int)(poll1(op.cand + 3 * r + 1, c.tag_out) | (poll1(op.cand + 3 * r + 2, c.tag_out) << 16));
                if (v > max_val || (v == max_val && idx < max_pos)) { max_val = v; max_pos = idx; }
            }
        }
        int token_pos = c.pos + 1;     // SharedData::pos and RunState::pos move together (gpu_kernels.h:486-491); no PCIe read
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
template <int F>
__global__ void __launch_bounds__(32 * (kMaxConsumerWarps + 1), 1) interp_kernel(const __grid_constant__ InterpParams P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const Op* ops = (P.ops != nullptr) ? P.ops : &P.one;
    Smem sm;
    sm.S = 0;
    sm.slot_bytes = 0;
    sm.bars = smem_u32(smem);
    sm.laps = sm.bars + kLapOffset;
    sm.xs = sm.bars + kCtrlBytes;
    sm.meta = sm.xs + P.xs_bytes;
    sm.meta_bytes = P.meta_bytes;
    sm.ring = sm.meta + P.meta_bytes + P.meta1_bytes;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kMaxSlots; s++) {
            mbar_init(sm.full(s), 1);
            mbar_init(sm.empty(s), 1);
            reinterpret_cast<volatile unsigned*>(smem + kLapOffset)[s] = 0u;
            reinterpret_cast<volatile unsigned*>(smem + kFillBaseOffset)[s] = 0u;
            reinterpret_cast<volatile unsigned*>(smem + kProdFillOffset)[s] = 0u;
        }
        for (int b = 0; b < 2; b++) { mbar_init(sm.mfull(b), 1); mbar_init(sm.mempty(b), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (warp == P.nwc) {
        if (lane == 0) producer_loop(P, ops, sm);
        return;
    }

    CtxT<F> c;
    c.P = &P; c.sm = sm;
    c.red = reinterpret_cast<float*>(smem + kRedOffset);
    c.scratch = smem + kCtrlBytes;
    c.nwc = P.nwc; c.nthreads = P.nwc * 32; c.warp = warp; c.lane = lane; c.ctid = threadIdx.x;
    c.pos = (P.pPos != nullptr) ? *P.pPos : 0;
    c.token = (P.tokens != nullptr) ? P.tokens[c.pos] : 0;
    c.qbase = 0; c.qtotal = 0; c.mcount = 0; c.meta_pending = -1; c.nsync = 0; c.tr = nullptr; c.tag_in = c.tag_out = 0; c.op_seq = 0;

    const Op& op = *reinterpret_cast<const Op*>(smem + kOpOffset);
    for (int o = 0; o < P.nops; o++) {
        // every warp has left the previous op: its staging areas, scale/zero buffer and op copy are free
        named_bar(kBarAll, c.nthreads);
        if (c.ctid < (int)(sizeof(Op) / 4))        // the op goes to shared memory: its fields are read many times per task
            reinterpret_cast<uint32_t*>(smem + kOpOffset)[c.ctid] = reinterpret_cast<const uint32_t*>(ops + o)[c.ctid];
        const int sync_before = ops[o].sync_before;
        if (ops[o].kind <= OP_CLS) ring_epoch(c, ops[o].slot_bytes, ops[o].nslots);
        trace_mark(c, 5);                          // previous op: every warp of this CTA is done
        c.tr = ((F & kDev) && P.trace != nullptr && o == P.trace_op) ? P.trace + 2048 + blockIdx.x * 8 : nullptr;
        trace_mark(c, 0);                          // this CTA arrives at the op
        cyc_mark(c, 7);
        if (c.ctid == 0) {
            if (c.meta_pending >= 0) mbar_arrive(sm.mempty(c.meta_pending));
            if (sync_before) grid_arrive(P.sync);
        }
        c.meta_pending = -1;
        c.op_seq = ((P.seq_base + (unsigned)o) & 0x3FFFFFFFu) + 1u;
        c.tag_out = ((P.seq_base + (unsigned)o + 1u) & 0x7FFFu) | 0x8000u;     // never 0: a zeroed buffer is never "fresh"
        c.tag_in = ((P.seq_base + (unsigned)o) & 0x7FFFu) | 0x8000u;           // the previous op's
        const bool attn_pref = (ops[o].kind == OP_ATTN);
        if (attn_pref) attn_prefetch(c, ops[o]);   // K rows of earlier positions do not depend on this launch at all
        if (sync_before) {
            c.nsync++;
            if (c.ctid == 0) grid_wait(P.sync, c.nsync * gridDim.x);
        }
        named_bar(kBarAll, c.nthreads);
        if ((F & kDev) && P.trace != nullptr && blockIdx.x == 0 && c.ctid == 0) P.trace[o] = global_ns();
        trace_mark(c, 1);                          // grid barrier passed
        cyc_mark(c, 0);
        switch (op.kind) {
            case OP_GEMV:
            case OP_FFN: run_q4(c, op); break;
            case OP_CLS: run_cls(c, op); break;
            case OP_ATTN: run_attn(c, op, attn_pref); break;
            case OP_ARGMAX: run_argmax(c, op, (P.write_token >= 0) ? P.write_token : op.write_token); break;
            default: break;
        }
    }
    if ((F & kDev) && P.trace != nullptr && blockIdx.x == 0 && c.ctid == 0) P.trace[P.nops] = global_ns();
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

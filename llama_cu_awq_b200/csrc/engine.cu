// engine.cu -- host side of the C ABI declared in include/llama_q4_b200.h.
//
// Mirrors the reference's host wrappers (llama2_q4.cu:207-432) one for one.  The decode step is ONE
// launch of the persistent op-interpreter kernel (interp_sm100.cuh) over a device op table; every
// operator of the API is the same kernel run over a single inline op.  Shapes the persistent kernel does
// not take (K % 64 != 0 and the like) go to the generic kernels in kernels_sm100.cuh.  No CPU fallback
// exists: without a CUDA device the calls fail loudly.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include <float.h>

#include <cub/cub.cuh>

#include "lq4_types.h"
#include "kernels_sm100.cuh"
#include "interp_sm100.cuh"
#include "prefill_sm100.cuh"
#include "synth.h"
#include "../../include/llama_q4_b200.h"

using namespace lq4;

namespace {

struct RopeKey {
    float theta; int head_size; int seq_len;
    bool operator<(const RopeKey& o) const {
        if (theta != o.theta) return theta < o.theta;
        if (head_size != o.head_size) return head_size < o.head_size;
        return seq_len < o.seq_len;
    }
};

// launch geometry of the persistent kernel (shared-memory map of interp_sm100.cuh)
struct Plan {
    int nwc = 0, ring_bytes = 0, meta_bytes = 0, meta1_bytes = 0, xs_bytes = 0;
    size_t smem = 0;
};

// cached per RunState: the device op table of a whole decode step
struct NetPlan {
    Plan plan;
    Op* d_ops = nullptr;
    int nops = 0;              // including the trailing OP_ARGMAX
    uint32_t* tagged = nullptr;   // flag-in-data activation vectors of the fused step: x | xb | hb | q | k (un-rotated) | v | argmax candidates
    uint16_t* att_sc = nullptr;   // split attention: [heads of this rank][seq_len] fp16 scores
    unsigned* att_flags = nullptr;
    uint32_t* peers[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // every rank's `tagged`, mapped here
    int world = 1, rank = 0;
    unsigned seq_base = 0;        // activation-tag base of the last launch of this plan
    const int* tokens = nullptr;  // SharedData::tokens of the RunState the plan was built for
    std::vector<char> key;     // Config + pointers the table was built from
    bool ok = false;           // false: some shape is not supported by the persistent kernel
};

struct Engine {
    bool inited = false;
    int device = 0;
    int sm_count = 0;
    int max_smem = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int opt_pdl = 1;        // generic per-op kernels only
    int opt_opk = 0;        // operator API: 1 = stand-alone kernels chained by programmatic dependent launch instead of one-op launches of the persistent kernel
    int opt_fused = 1;      // run_llama_network: one persistent launch per token (1) or op-by-op like the reference (0)
    int opt_nwc = 0;        // consumer warps per CTA; 0 = choose per model
    int opt_tp_repl_o = -1; // tensor parallel: every rank computes the whole o projection (no exchange of x after it); -1 = from 8 ranks on
                            // (measured, 7B: 851 instead of 830 tok/s at 8 ranks, 830 instead of 858 at 4, 732 instead of 734 at 2)
    int opt_cls_rpt = 0;    // classifier rows per warp-task (1, 2, 4); 0 = choose by row length (development aid)
    int opt_nslots = 0;     // cap on ring slots; 0 = as many as fit
    int opt_slot_bytes = 0; // ring slot size; 0 = the largest minimum chunk of the model's ops
    std::map<RopeKey, float2*> rope_tabs;
    std::map<const void*, NetPlan> nets;
    unsigned* sync = nullptr;
    int tp_rank = 0, tp_world = 1;       // tensor parallel: one process per GPU, this one's rank
    int opt_trace = 0;      // record per-op timestamps of fused steps (development aid)
    int opt_trace_op = -1;  // op whose phases are recorded per CTA
    int opt_nomath = 0;     // development aid (LQ4_NOMATH=1): data-path-only timing, results are garbage
    unsigned long long* trace = nullptr;
    int trace_n = 0;
    std::map<const void*, void*> arenas;   // Transformer* -> device arena (loader)
    uint16_t* att_sc = nullptr;            // operator API: score exchange of the split attention (grown on demand)
    size_t att_sc_elems = 0;
    unsigned* att_flags = nullptr;         // [heads][4], zero-initialised once
    int att_flag_heads = 0;
    unsigned op_seq = 0;                   // single-op launches: a fresh Ctx::op_seq base per launch
    int* fault = nullptr;                  // [4] pinned host record a timed-out device wait leaves before it traps (interp_sm100.cuh, g_fault)
    char err[512] = {0};
};
Engine g;

// The persistent kernel's instances (interp_sm100.cuh, kSplit | kTP | kDev): production launches take the leanest instance that
// covers them; anything with the development aids switched on takes the full one.
typedef void (*InterpFn)(const InterpParams);
constexpr int kNumInterpInstances = 8;
InterpFn interp_instance(int f) {      // f = kSplit | kTP | kDev
    switch (f) {
        case 0: return interp_kernel<0>;
        case 1: return interp_kernel<1>;
        case 2: return interp_kernel<2>;
        case 3: return interp_kernel<3>;
        case 4: return interp_kernel<4>;
        case 5: return interp_kernel<5>;
        case 6: return interp_kernel<6>;
        default: return interp_kernel<7>;
    }
}
InterpFn interp_pick(bool long_ctx, bool tp, bool dev) { return interp_instance((long_ctx ? kSplit : 0) | (tp ? kTP : 0) | (dev ? kDev : 0)); }

void set_err(const char* what, cudaError_t e) {
    int n = snprintf(g.err, sizeof g.err, "%s: %s", what, cudaGetErrorString(e));
    if (g.fault != nullptr && g.fault[0] != 0 && n > 0 && n < (int)sizeof g.err) {      // the kernel said why it gave up
        static const char* const what_timed_out[] = {"?", "weights (a bulk copy never completed)",
            "activations (the previous op's output never arrived; under tensor parallelism: a dead or stalled peer)", "the grid barrier",
            "a ring slot that was never released", "the split-attention score flags"};
        const int code = g.fault[0];
        snprintf(g.err + n, sizeof g.err - n, " -- device protocol time-out after 5 s waiting for %s on CTA %d (thread %d, SM %d, rank %d of %d)",
                 what_timed_out[(code >= 1 && code <= 5) ? code : 0], g.fault[1], g.fault[2], g.fault[3], g.tp_rank, g.tp_world);
    }
    fprintf(stderr, "lq4: %s\n", g.err);
}

#define LQ4_CHECK(call)                                                   \
    do {                                                                  \
        cudaError_t e_ = (call);                                          \
        if (e_ != cudaSuccess) { set_err(#call, e_); exit(EXIT_FAILURE); } \
    } while (0)

void ensure_init() {
    if (g.inited) return;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        fprintf(stderr, "lq4: no CUDA device available (%s); this engine has no CPU path\n",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        exit(EXIT_FAILURE);
    }
    LQ4_CHECK(cudaGetDevice(&g.device));
    cudaDeviceProp prop;
    LQ4_CHECK(cudaGetDeviceProperties(&prop, g.device));
    if (prop.major != 10) {
        fprintf(stderr, "lq4: device '%s' is sm_%d%d; this library holds sm_100a code only\n", prop.name, prop.major,
                prop.minor);
        exit(EXIT_FAILURE);
    }
    g.sm_count = prop.multiProcessorCount;
    g.max_smem = (int)prop.sharedMemPerBlockOptin;
    if (!g.stream) { LQ4_CHECK(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking)); g.own_stream = true; }
    const char* env;
    if ((env = getenv("LQ4_PDL"))) g.opt_pdl = atoi(env);
    if ((env = getenv("LQ4_OPK"))) g.opt_opk = atoi(env);
    if ((env = getenv("LQ4_FUSED"))) g.opt_fused = atoi(env);
    if ((env = getenv("LQ4_NWC"))) { g.opt_nwc = atoi(env); if (g.opt_nwc != 0 && g.opt_nwc < 8) g.opt_nwc = 8; }
    if ((env = getenv("LQ4_NOMATH"))) g.opt_nomath = atoi(env);
    if ((env = getenv("LQ4_NSLOTS"))) g.opt_nslots = atoi(env);
    if ((env = getenv("LQ4_TP_REPL_O"))) g.opt_tp_repl_o = atoi(env);
    if ((env = getenv("LQ4_CLS_RPT"))) { const int v = atoi(env); if (v == 1 || v == 2 || v == 4) g.opt_cls_rpt = v; }
    if ((env = getenv("LQ4_SLOT_BYTES"))) g.opt_slot_bytes = atoi(env);
    LQ4_CHECK(cudaMallocHost((void**)&g.fault, 4 * sizeof(int)));
    memset(g.fault, 0, 4 * sizeof(int));
    LQ4_CHECK(cudaMemcpyToSymbol(lq4::g_fault, &g.fault, sizeof(int*)));
    LQ4_CHECK(cudaMalloc((void**)&g.sync, 2 * sizeof(unsigned)));
    LQ4_CHECK(cudaMemset(g.sync, 0, 2 * sizeof(unsigned)));
    for (int f = 0; f < kNumInterpInstances; f++)
        LQ4_CHECK(cudaFuncSetAttribute(interp_instance(f), cudaFuncAttributeMaxDynamicSharedMemorySize, g.max_smem));
    g.inited = true;
}

[[noreturn]] void unsupported() {
    printf("\nUnsupported matmul size. Exiting\n");   // llama2_q4.cu:215,225,236,251
    exit(EXIT_FAILURE);
}

QW qw_view(const QWeight* w) { return QW{w->weight, w->zeros, reinterpret_cast<const uint16_t*>(w->scales)}; }

template <typename... KArgs, typename... Args>
void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, bool pdl, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = g.stream;
    cudaLaunchAttribute attr[1];
    if (pdl) {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
    if (e != cudaSuccess) { set_err("kernel launch", e); exit(EXIT_FAILURE); }
}

template <typename K>
void allow_smem(K kernel, size_t bytes) {
    static std::map<const void*, size_t> granted;
    size_t& cur = granted[(const void*)kernel];
    if (bytes > 48 * 1024 && bytes > cur) {
        LQ4_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        cur = bytes;
    }
}

// ------------------------------------------------------------------------------- generic kernels (odd shapes)
size_t gemv_smem(int K) { return (size_t)(((K + 1023) & ~1023) + 32) * sizeof(float); }

template <int KIND>
void launch_gemv(const GemvParams& p, int ntasks, bool pdl) {
    const size_t smem = gemv_smem(p.K);
    allow_smem(gemv_q4_kernel<KIND>, smem);
    int blocks = (ntasks + kGemvWarps - 1) / kGemvWarps;
    blocks = std::min(blocks, 2 * g.sm_count);
    launch(gemv_q4_kernel<KIND>, dim3(blocks), dim3(kGemvThreads), smem, pdl, p);
}

void check_q4_shape(int K, int N) {
    if ((K & 7) || (N & 7)) unsupported();     // the reference's own check
    if (K & 31) unsupported();                 // packed height is padded to 32 (llama2_q4.cu:82-88); K%32 != 0
                                               // makes the packer's and the runtime's layouts disagree
}

float2* rope_table(float theta, int head_size, int seq_len) {
    RopeKey key{theta, head_size, seq_len};
    auto it = g.rope_tabs.find(key);
    if (it != g.rope_tabs.end()) return it->second;
    float2* tab = nullptr;
    const int n = seq_len * (head_size / 2);
    LQ4_CHECK(cudaMalloc(&tab, sizeof(float2) * (size_t)n));
    rope_table_kernel<<<(n + 255) / 256, 256, 0, g.stream>>>(tab, seq_len, head_size, theta);
    LQ4_CHECK(cudaStreamSynchronize(g.stream));
    g.rope_tabs[key] = tab;
    return tab;
}

// wall clock of the host loops in ns (the reference's time_in_ms, llama2_q4.cu:400-405, rounds to 1 ms: too coarse for a
// 20-step run of 2 ms steps)
long long time_in_ns() {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (long long)t.tv_sec * 1000000000ll + t.tv_nsec;
}

// ------------------------------------------------------------------------------- persistent kernel: plans and ops
int attn_scratch_bytes(int head_size, int max_seq) { return attn_fixed_bytes(head_size, max_seq) + 2 * attn_buf_bytes(head_size); }

// Shared-memory plan for `nwc` consumer warps: a staging area of `xs` bytes (activations / attention scratch),
// two buffers of scales / zero points (`meta` bytes for the even INT4 ops of the table, `meta1` for the odd ones: they are
// used alternately) and as many ring slots of `slot` bytes as fit.
bool make_plan(Plan& pl, int nwc, int xs, int meta, int meta1) {
    pl.nwc = nwc;
    pl.xs_bytes = (xs + 127) & ~127;
    pl.meta_bytes = (meta + 127) & ~127;
    pl.meta1_bytes = (meta1 + 127) & ~127;
    const int fixed = kCtrlBytes + pl.xs_bytes + pl.meta_bytes + pl.meta1_bytes;
    const int ring = (g.max_smem - fixed) & ~127;
    if (ring < 4 * 128) return false;
    pl.ring_bytes = ring;
    pl.smem = (size_t)fixed + (size_t)ring;
    return true;
}

bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

// fills the shape-derived fields of a q4 op; false when the persistent kernel cannot take it
bool q4_op_shape(Op& op, int K, const int* ncols, int nseg, bool dual) {
    if (K % 64 != 0 || K < 64) return false;
    op.K = K;
    op.T = (K + 1023) / 1024;
    op.nseg = nseg;
    int total = 0;
    for (int s = 0; s < nseg; s++) {
        if (ncols[s] % 8) return false;
        if (!aligned16(op.seg[s].w) || !aligned16(op.seg[s].z) || !aligned16(op.seg[s].s)) return false;
        total += ncols[s];
    }
    if (dual && ncols[0] != ncols[1]) return false;
    op.ntasks = dual ? ncols[0] / 2 : total / 4;
    // CTA ranges start on multiples of `au` tasks so that the scale / zero rows of a range are whole 16-byte units
    const int G = q4_groups(K), zh = q4_zh(K);
    if (dual) op.au = (G % 4 == 0 && zh % 2 == 0) ? 1 : (G % 2 == 0) ? 2 : 4;
    else op.au = (G % 2 == 0) ? 1 : 2;
    if (op.ntasks % op.au) return false;
    return true;
}
constexpr int kMaxChunk = 16384;     // largest ring slot
bool cls_op_shape(Op& op, int n, int d, int row_stride) {
    if ((n & 7) || (row_stride & 7) || (d & 3) || !aligned16(op.seg[0].w)) return false;
    op.K = n;
    op.T = (n + 255) / 256;
    op.nseg = 1;
    op.seg[0].ncols = d;
    // rows per warp-task: short rows go four to a slot; long rows two to a task, so that one bulk copy brings 16 KB (the producer
    // lane issues a copy per ~0.2 us whatever its size: with one 8 KB row per copy the 7B classifier streamed at 4.2 TB/s)
    op.rpt = (g.opt_cls_rpt > 0) ? g.opt_cls_rpt : (n * 2 < 4096) ? 4 : (n * 4 <= kMaxChunk) ? 2 : 1;      // 13B (10 KB rows): two rows would take two slots, measured slower
    op.ntasks = (d + op.rpt - 1) / op.rpt;
    op.au = 1;
    op.row_stride = row_stride;
    return true;
}
// bytes of the smallest ring chunk the op can be cut into (one column | one gate/up pair | one row)
int op_min_chunk(const Op& op) {
    if (op.kind == OP_CLS) return op.K * 2;
    return (op.kind == OP_FFN ? 2 : 1) * q4_col_bytes(op.K);
}
// bytes of a whole warp-task (4 columns | 2 gate/up pairs | 4 rows)
int op_task_chunks(const Op& op) { return op.kind == OP_FFN ? 2 : op.kind == OP_CLS ? op.rpt : 4; }
// Cut the op's tasks into ring slots: the largest piece of a task (whole, half, quarter) of at most kMaxChunk bytes is one
// slot, and the ring holds as many of those as fit.  Ops with the same slot size share a ring epoch (interp_sm100.cuh).
bool op_set_chunking(Op& op, int ring_bytes) {
    const int per = op_task_chunks(op);             // chunks of minimum size per task
    const int limit = g.opt_slot_bytes > 0 ? std::max(g.opt_slot_bytes, op_min_chunk(op)) : kMaxChunk;
    int cps = per;
    while (cps > 1 && cps * op_min_chunk(op) > limit) cps >>= 1;
    const int slot = (cps * op_min_chunk(op) + 127) & ~127;
    int n = std::min(ring_bytes / slot, kMaxSlots);
    if (g.opt_nslots > 0) n = std::min(n, g.opt_nslots);
    op.cps = cps;
    op.spt = per / cps;
    op.slot_bytes = slot;
    op.nslots = n;
    return n >= op.spt && n >= 2;
}
// staging area: fp32 pairs (INT4 ops) or fp16 x (classifier, plus its parked norm weights when RMSNorm is fused)
int op_xs_bytes(const Op& op) {
    return op.kind == OP_CLS ? ((op.K * 2 + 127) & ~127) * (op.norm_w != nullptr ? 2 : 1) : op.T * kTripBytes;
}
int op_meta_bytes(const Op& op, int grid) {
    if (op.kind == OP_CLS) return 0;
    const int units = op.ntasks / op.au;
    const int maxtasks = ((units + grid - 1) / grid) * op.au;
    const int cols = maxtasks * 4;                   // 4 columns, or 2 gate + 2 up columns, per task
    return cols * (q4_groups(op.K) * 2 + q4_zh(op.K) * 4) + 64;
}

void set_seg(Seg& sg, const QWeight* w, half* out, int ncols, int loff, int pos_stride, uint32_t* out32 = nullptr, int bcast = 0) {
    sg.w = w->weight; sg.z = w->zeros; sg.s = reinterpret_cast<const uint16_t*>(w->scales);
    sg.out = out; sg.out32 = out32; sg.ncols = ncols; sg.loff = loff; sg.pos_stride = pos_stride; sg.bcast = bcast;
}
// columns [c0, c0 + n) of a QWeight(K, N): a contiguous byte range of each of its three arrays (tensor-parallel slices)
QWeight slice_cols(const QWeight& w, int K, int c0) {
    QWeight r;
    r.weight = w.weight + (size_t)c0 * (size_t)(divUp(K, 32) * 4);
    r.zeros = w.zeros + (size_t)c0 * (size_t)q4_zh(K);
    r.scales = w.scales + (size_t)c0 * (size_t)q4_groups(K);
    return r;
}

// long_ctx: launch the kernel instance that can split a head's attention over several CTAs (positions >= kAttnSplitFrom).  Both
// instances give bit-identical results at every position, so a host that only knows a bound of the position may pick either.
void launch_interp(const Plan& pl, const Op* d_ops, int nops, const Op* one, const int* pPos, int write_token,
                   bool cooperative, int grid, const NetPlan* tp_or_plan = nullptr, bool long_ctx = false) {
    InterpParams P;
    memset(&P, 0, sizeof P);
    P.ops = d_ops; P.nops = nops;
    P.nwc = pl.nwc; P.ring_bytes = pl.ring_bytes; P.meta_bytes = pl.meta_bytes; P.meta1_bytes = pl.meta1_bytes; P.xs_bytes = pl.xs_bytes;
    P.write_token = write_token;
    P.sync = g.sync; P.pPos = pPos;
    // Activation tags: consecutive launches over the same buffers must never share a tag.  Each plan advances its own base by
    // its op count per launch, so two consecutive launches differ by nops (mod 2^15), whatever else ran in between; all ranks
    // of a tensor-parallel group launch in lockstep and therefore agree on it.
    g.op_seq = (g.op_seq + 1u) & 0x3FFFFFFFu;
    P.seq_base = g.op_seq;        // single-op launches: only Ctx::op_seq (split-attention flags) depends on it
    if (tp_or_plan != nullptr) {
        NetPlan* np = const_cast<NetPlan*>(tp_or_plan);
        np->seq_base = (np->seq_base + (unsigned)np->nops) & 0x3FFFFFFFu;
        P.seq_base = np->seq_base;
        P.tokens = np->tokens;
    }
    const NetPlan* tp = (tp_or_plan != nullptr && tp_or_plan->world > 1) ? tp_or_plan : nullptr;
    P.rank = 0; P.world = 1;
    if (tp != nullptr) {
        P.rank = tp->rank; P.world = tp->world;
        for (int r = 0; r < tp->world; r++) {
            if (tp->peers[r] == nullptr) { fprintf(stderr, "lq4: tensor-parallel peer %d has not been imported (lq4_tp_import)\n", r); exit(EXIT_FAILURE); }
            P.peers[r] = tp->peers[r];
        }
    }
    if (g.opt_trace && d_ops != nullptr && nops + 1 <= 2048) {
        if (!g.trace) { LQ4_CHECK(cudaMalloc((void**)&g.trace, 32768 * sizeof(unsigned long long))); LQ4_CHECK(cudaMemset(g.trace, 0, 32768 * sizeof(unsigned long long))); }
        P.trace = g.trace;
        P.trace_op = g.opt_trace_op;
        g.trace_n = nops + 1;
    }
    P.nomath = g.opt_nomath;
    if (one) P.one = *one;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(32 * (pl.nwc + 1));
    cfg.dynamicSmemBytes = pl.smem;
    cfg.stream = g.stream;
    cudaLaunchAttribute attr[1];
    if (cooperative) {
        attr[0].id = cudaLaunchAttributeCooperative;
        attr[0].val.cooperative = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    const bool dev = (P.trace != nullptr) || P.nomath;
    cudaError_t e = cudaLaunchKernelEx(&cfg, interp_pick(long_ctx, tp != nullptr, dev), P);
    if (e != cudaSuccess) { set_err("interp_kernel launch", e); exit(EXIT_FAILURE); }
}

int default_nwc() { return g.opt_nwc > 0 ? std::min(g.opt_nwc, kMaxConsumerWarps) : kMaxConsumerWarps; }

// one op through the persistent kernel (operator API)
void run_single(Op& op, const int* pPos) {
    Plan pl;
    int grid = g.sm_count, xs = 0, meta = 0;
    if (op.kind <= OP_CLS) {
        grid = std::max(1, std::min(g.sm_count, op.ntasks));
        xs = op_xs_bytes(op);
        meta = op_meta_bytes(op, grid);
    } else if (op.kind == OP_ATTN) {
        // long contexts run 2 or 4 CTAs per head (interp_sm100.cuh, run_attn_split): scratch for the score exchange, grown on demand
        const size_t need = (size_t)op.n_heads * op.max_seq;
        if (g.att_sc_elems < need || g.att_flag_heads < op.n_heads) {
            LQ4_CHECK(cudaStreamSynchronize(g.stream));
            cudaFree(g.att_sc); cudaFree(g.att_flags);
            g.att_sc_elems = std::max(need, g.att_sc_elems); g.att_flag_heads = std::max(op.n_heads, g.att_flag_heads);
            LQ4_CHECK(cudaMalloc((void**)&g.att_sc, g.att_sc_elems * sizeof(uint16_t)));
            LQ4_CHECK(cudaMalloc((void**)&g.att_flags, (size_t)g.att_flag_heads * 4 * sizeof(unsigned)));
            LQ4_CHECK(cudaMemset(g.att_flags, 0, (size_t)g.att_flag_heads * 4 * sizeof(unsigned)));
        }
        op.att_split = 4; op.att_sc = g.att_sc; op.att_sc_stride = op.max_seq; op.att_flags = g.att_flags;
        grid = std::min(g.sm_count, op.n_heads * op.att_split);
        xs = attn_scratch_bytes(op.head_size, op.max_seq);
    } else {
        grid = 1;
    }
    // the attention op keeps four PV chains per warp (tx = warp + k * nwc, k < 4): it needs at least 8 consumer warps
    const int nwc = (op.kind == OP_ATTN) ? std::max(default_nwc(), 8) : default_nwc();
    if (!make_plan(pl, nwc, xs, meta, 0)) unsupported();      // a single op only ever uses buffer 0
    if (op.kind <= OP_CLS && !op_set_chunking(op, pl.ring_bytes)) unsupported();
    launch_interp(pl, nullptr, 1, &op, pPos, -1, false, grid, nullptr, op.kind == OP_ATTN);      // the position is on the device: the attention op always gets the instance that can split
}

// ------------------------------------------------------------------------------- temperature / top-p sampler (scope row f3)
// Off the hot path (the metric is greedy).  sampler.h:51-81 is a five-step pipeline over fp16 storage: temperature + softmax,
// radix sort by probability, inclusive prefix sum, threshold search.  Two of those steps are toolkit library calls
// (cub::DeviceRadixSort / cub::DeviceScan) and STAY library calls here, for a parity reason that is worth spelling out:
//   * the prefix sum runs IN FP16 (accumulator type = the half input type).  Its value at every position depends on the
//     association of the additions, which is cub's tile / warp-scan / look-back structure for this toolkit and architecture --
//     with ~32000 probabilities of ~3e-5 each the partial sums sit where one fp16 ulp is as large as several addends, so any
//     other association (a sequential scan, a different tiling) moves the position where the sum crosses `coin * topp` and
//     therefore the sampled token.  Token parity with the reference for a fixed seed is only defined through the same library.
//   * the sort only has to be stable and descending on the fp16 key bits (any stable sort gives the same permutation), but
//     its output feeds the scan above in place, so it is kept beside it.
// What is ours are the two kernels around the library calls, written for one 1024-thread CTA with the reductions of
// kernels_sm100.cuh: temperature_softmax_kernel fuses the temperature division, the softmax (fp16 round trips and fp32
// summation order of gpu_kernels.h:499-550: strided per-thread partial sums, warp trees, warp aggregates added in order) and
// the index initialisation; threshold_pick_kernel finds the first position whose prefix sum reaches the threshold and
// publishes token and position like the greedy sampler does.
constexpr int kSampThreads = 1024;

__device__ __forceinline__ float block_sum_in_order(float v, float* red) {      // cub::BlockReduce<float,1024>::Sum association
    v = warp_tree_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float tot = red[0];
#pragma unroll
    for (int w = 1; w < kSampThreads / 32; w++) tot = tot + red[w];
    __syncthreads();
    return tot;
}

__global__ void __launch_bounds__(kSampThreads) temperature_softmax_kernel(half* probs, int n, float temperature, int* indices) {
    __shared__ float red[kSampThreads / 32];
    const int tid = threadIdx.x;
    // pass 1: logits / temperature, stored back in fp16 (the reference's round trip), running maximum of the rounded values
    float mx = -FLT_MAX;
    for (int i = tid; i < n; i += kSampThreads) {
        const half h = __float2half_rn(__fdiv_rn(__half2float(probs[i]), temperature));
        probs[i] = h;
        indices[i] = i;
        mx = fmaxf(mx, __half2float(h));
    }
    mx = warp_max(mx);
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int w = 1; w < kSampThreads / 32; w++) mx = fmaxf(mx, red[w]);
    __syncthreads();
    // pass 2: exp(x - max) parked in fp16, the sum adds the unrounded values (per-thread strided order, then the cub tree)
    float part = 0.0f;
    for (int i = tid; i < n; i += kSampThreads) {
        const float e = expf(__fsub_rn(__half2float(probs[i]), mx));
        probs[i] = __float2half_rn(e);
        part = __fadd_rn(part, e);
    }
    const float total = block_sum_in_order(part, red);
    // pass 3: normalise the fp16-rounded exponentials
    for (int i = tid; i < n; i += kSampThreads) probs[i] = __float2half_rn(__fdiv_rn(__half2float(probs[i]), total));
}

__global__ void __launch_bounds__(kSampThreads) threshold_pick_kernel(const half* prefix, const int* indices, int n, float threshold, int* tokens,
                                                                      volatile int* pos_host, int* pos_dev) {
    __shared__ int first[kSampThreads / 32];
    const int tid = threadIdx.x;
    int best = n - 1;                                   // nothing reaches the threshold: the last entry (gpu_kernels.h:559)
    for (int i = tid; i < n && i < best; i += kSampThreads)
        if (__half2float(prefix[i]) >= threshold) { best = i; break; }      // a thread's positions ascend: its first hit is its minimum
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((tid & 31) == 0) first[tid >> 5] = best;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < kSampThreads / 32; w++) best = min(best, first[w]);
        const int next_pos = *pos_host + 1;
        tokens[next_pos] = indices[best];
        *pos_host = next_pos;
        *pos_dev = next_pos;
    }
}

}  // namespace

// =====================================================================================================
extern "C" {

int lq4_init(int device) {
    if (!g.inited) {
        cudaError_t e = cudaSetDevice(device);
        if (e != cudaSuccess) { set_err("cudaSetDevice", e); return 1; }
    }
    ensure_init();
    return 0;
}
void* lq4_get_stream(void) { ensure_init(); return (void*)g.stream; }
void lq4_set_stream(void* s) {
    ensure_init();
    if (g.own_stream && g.stream) cudaStreamDestroy(g.stream);
    g.stream = (cudaStream_t)s;
    g.own_stream = false;
}
int lq4_stream_synchronize(void) {
    ensure_init();
    cudaError_t e = cudaStreamSynchronize(g.stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { set_err("stream synchronize", e); return 1; }
    return 0;
}
int lq4_stream_query(void) {
    ensure_init();
    cudaError_t e = cudaStreamQuery(g.stream);
    if (e == cudaSuccess) return 0;
    if (e == cudaErrorNotReady) return 1;
    set_err("stream query", e);
    return 2;
}
const char* lq4_last_error(void) { return g.err; }
int lq4_sm_count(void) { ensure_init(); return g.sm_count; }

static void drop_net_plans() {
    for (auto& kv : g.nets) { cudaFree(kv.second.d_ops); cudaFree(kv.second.tagged); cudaFree(kv.second.att_sc); cudaFree(kv.second.att_flags); }
    g.nets.clear();
}

void lq4_set_option(const char* name, int value) {
    ensure_init();
    if (!strcmp(name, "pdl")) g.opt_pdl = value;
    else if (!strcmp(name, "opk")) g.opt_opk = value;
    else if (!strcmp(name, "fused")) g.opt_fused = value;
    else if (!strcmp(name, "nwc")) {
        if (value != 0 && value < 8) { fprintf(stderr, "lq4: nwc must be 0 (default) or >= 8 (RMSNorm staging and the attention op need 8 consumer warps); ignored\n"); return; }
        g.opt_nwc = value; cudaStreamSynchronize(g.stream); drop_net_plans();
    }
    else if (!strcmp(name, "slot_bytes")) { g.opt_slot_bytes = value; cudaStreamSynchronize(g.stream); drop_net_plans(); }
    else if (!strcmp(name, "nslots")) { g.opt_nslots = value; cudaStreamSynchronize(g.stream); drop_net_plans(); }
    else if (!strcmp(name, "trace")) g.opt_trace = value;
    else if (!strcmp(name, "trace_op")) g.opt_trace_op = value;
    else if (!strcmp(name, "graphs")) { /* the decode step is a single launch: nothing to capture */ }
}

// development aid: timestamps (ns) of the last fused step, one per op start plus the end; returns the count
int lq4_debug_trace(unsigned long long* out, int* kinds, int max) {
    ensure_init();
    if (!g.trace || g.trace_n == 0) return 0;
    cudaStreamSynchronize(g.stream);
    const int n = std::min(max, g.trace_n);
    cudaMemcpy(out, g.trace, sizeof(unsigned long long) * n, cudaMemcpyDeviceToHost);
    if (max >= 8192) cudaMemcpy(out + 2048, g.trace + 2048, sizeof(unsigned long long) * (max >= 32768 ? 30720 : 6144), cudaMemcpyDeviceToHost);
    if (kinds && !g.nets.empty()) {
        NetPlan& np = g.nets.begin()->second;
        std::vector<Op> ops(np.nops);
        cudaMemcpy(ops.data(), np.d_ops, sizeof(Op) * np.nops, cudaMemcpyDeviceToHost);
        for (int i = 0; i < n && i < np.nops; i++) kinds[i] = ops[i].kind * 100000 + ops[i].K;
    }
    return n;
}

// ---------------------------------------------------------------------------------- operator API
void lq4_rmsnorm(half* o, half* x, half* weight, int size) {
    ensure_init();
    launch(rmsnorm_kernel, dim3(1), dim3(1024), 0, g.opt_opk && g.opt_pdl, o, (const half*)x, (const half*)weight, size);
}

void lq4_matmul_fp16(half* xout, half* x, half* w, int n, int d, int batch, int x_stride, int w_stride, int op_stride,
                     int w_row_stride, float alpha) {
    ensure_init();
    if ((n & 7) || (d & 7)) unsupported();
    if (w_row_stride == -1) w_row_stride = n;
    for (int b = 0; b < batch; b++) {   // the reference batches through blockIdx.y; the only caller uses batch 1
        Op op;
        memset(&op, 0, sizeof op);
        op.kind = OP_CLS;
        op.seg[0].w = reinterpret_cast<const uint32_t*>(w + (size_t)b * w_stride);
        op.seg[0].out = xout + (size_t)b * op_stride;
        op.x = x + (size_t)b * x_stride;
        op.alpha = alpha;
        if (!cls_op_shape(op, n, d, w_row_stride)) unsupported();
        run_single(op, nullptr);
    }
}

void lq4_matmul_q4(half* xout, half* x, const QWeight* w, int inpSize, int opSize, int accum, int loff, int* pPos) {
    ensure_init();
    check_q4_shape(inpSize, opSize);
    Op op;
    memset(&op, 0, sizeof op);
    op.kind = OP_GEMV;
    op.x = x; op.accum = accum;
    const bool cache_row = (loff != -1);
    set_seg(op.seg[0], w, xout, opSize, cache_row ? loff : 0, cache_row ? opSize : 0);
    if (!g.opt_opk && q4_op_shape(op, inpSize, &opSize, 1, false)) {
        run_single(op, cache_row ? pPos : nullptr);
        return;
    }
    GemvParams p = {};
    p.x = x; p.K = inpSize;
    p.m[0] = qw_view(w); p.n[0] = opSize; p.out[0] = xout;
    p.accum = accum; p.loff = loff; p.pPos = cache_row ? pPos : nullptr;
    launch_gemv<GEMV_PLAIN>(p, opSize / 4, g.opt_opk && g.opt_pdl);
}

void lq4_qkv_matvec(half* q, half* key_cache, half* value_cache, half* x, const QWeight* qw, const QWeight* kw,
                    const QWeight* vw, int inpSize, int opSize, int loff, int* pPos) {
    ensure_init();
    check_q4_shape(inpSize, opSize);
    Op op;
    memset(&op, 0, sizeof op);
    op.kind = OP_GEMV;
    op.x = x;
    set_seg(op.seg[0], qw, q, opSize, 0, 0);
    set_seg(op.seg[1], kw, key_cache, opSize, loff, opSize);
    set_seg(op.seg[2], vw, value_cache, opSize, loff, opSize);
    const int nc[3] = {opSize, opSize, opSize};
    if (!g.opt_opk && q4_op_shape(op, inpSize, nc, 3, false)) {
        run_single(op, pPos);
        return;
    }
    GemvParams p = {};
    p.x = x; p.K = inpSize;
    p.m[0] = qw_view(qw); p.m[1] = qw_view(kw); p.m[2] = qw_view(vw);
    p.n[0] = p.n[1] = p.n[2] = opSize;
    p.out[0] = q; p.out[1] = key_cache; p.out[2] = value_cache;
    p.loff = loff; p.pPos = pPos;
    launch_gemv<GEMV_QKV>(p, 3 * opSize / 4, g.opt_opk && g.opt_pdl);
}

void lq4_ffn_matvec_silu(half* xout, half* x, const QWeight* gate_w, const QWeight* up_w, int inpSize, int opSize) {
    ensure_init();
    check_q4_shape(inpSize, opSize);
    Op op;
    memset(&op, 0, sizeof op);
    op.kind = OP_FFN;
    op.x = x;
    set_seg(op.seg[0], gate_w, xout, opSize, 0, 0);
    set_seg(op.seg[1], up_w, xout, opSize, 0, 0);
    const int nc[2] = {opSize, opSize};
    if (!g.opt_opk && q4_op_shape(op, inpSize, nc, 2, true)) {
        run_single(op, nullptr);
        return;
    }
    GemvParams p = {};
    p.x = x; p.K = inpSize;
    p.m[0] = qw_view(gate_w); p.m[1] = qw_view(up_w);
    p.n[0] = opSize; p.out[0] = xout;
    p.loff = -1;
    launch_gemv<GEMV_FFN>(p, opSize / 2, g.opt_opk && g.opt_pdl);
}

void lq4_rope_rotation(half* q, half* k, int num_heads, int num_kv_heads, int head_size, int* pPos, int loff,
                       float rope_theta) {
    ensure_init();
    launch(rope_kernel, dim3(num_heads), dim3(head_size / 2), 0, g.opt_opk && g.opt_pdl, q, k, num_kv_heads, head_size, (const int*)pPos, loff, rope_theta);
}

static void fill_attn_op(Op& op, half* output, half* q, half* key_cache, half* value_cache, half* att, int num_heads,
                         int head_size, int kv_mul, int max_seq_len) {
    if (head_size != 32 && head_size != 64 && head_size != 128) unsupported();
    op.kind = OP_ATTN;
    // MultiHeadAttention picks the softmax kernel by its max_seq_len argument (llama2_q4.cu:276-279): beyond 8192 it is
    // softmax_kernel_no_smem, whose exp() values are rounded to fp16 before the division.  The fused step overrides this with
    // the position-dependent rule the reference's graph bins amount to (see get_net_plan).
    op.exp16_from = (max_seq_len > MAX_SEQ_LEN_SMEM_KERNEL) ? 0 : 0x7fffffff;
    op.q = q; op.kcache = key_cache; op.vcache = value_cache; op.att_out = att; op.attn_out = output;
    op.n_heads = num_heads; op.head_size = head_size; op.kv_mul = kv_mul;
    op.kv_stride = (num_heads * head_size) / kv_mul;
    op.max_seq = max_seq_len;
    op.att_alpha = (float)(1.0 / sqrt((double)head_size));     // llama2_q4.cu:273
}

void lq4_multi_head_attention(half* output, half* q, half* key_cache, half* value_cache, half* att, int num_heads,
                              int head_size, int kv_mul, int max_seq_len, int* pPos) {
    ensure_init();
    if ((g.opt_opk || (head_size != 32 && head_size != 64 && head_size != 128)) && head_size % 32 == 0 && head_size <= 256) {
        // other head sizes: the stand-alone kernel (same arithmetic, one block of 1024 threads per head)
        AttnParams ap = {};
        ap.out = output; ap.q = q; ap.kcache = key_cache; ap.vcache = value_cache; ap.att_out = att;
        ap.head_size = head_size; ap.kv_mul = kv_mul; ap.kv_stride = (num_heads * head_size) / kv_mul;
        ap.pPos = pPos; ap.alpha = (float)(1.0 / sqrt((double)head_size)); ap.max_seq = max_seq_len;
        ap.exp16 = max_seq_len > MAX_SEQ_LEN_SMEM_KERNEL;
        const size_t smem = sizeof(float) * (size_t)(head_size + 32 + 4 + ((max_seq_len + 3) & ~3) + 32 * head_size);
        allow_smem(attention_kernel, smem);
        launch(attention_kernel, dim3(num_heads), dim3(kAttnThreads), smem, g.opt_opk && g.opt_pdl, ap);
        return;
    }
    Op op;
    memset(&op, 0, sizeof op);
    fill_attn_op(op, output, q, key_cache, value_cache, att, num_heads, head_size, kv_mul, max_seq_len);
    run_single(op, pPos);
}

// ---------------------------------------------------------------------------------- forward pass
static void run_network_unfused(int* pPos, Config* p, RunState* s, TransformerWeights* w, int seq_len_bin) {
    // op-by-op, exactly the sequence of llama2_q4.cu:286-340
    half* x = s->x;
    const int dim = p->dim, hidden_dim = p->hidden_dim;
    const int head_size = dim / p->n_heads;
    const int kv_dim = (p->dim * p->n_kv_heads) / p->n_heads;
    const int kv_mul = p->n_heads / p->n_kv_heads;
    launch(copy_embedding_kernel, dim3(divUp(dim, 256)), dim3(256), 0, g.opt_opk && g.opt_pdl, x, (const half*)w->token_embedding_table, dim,
           (const int*)s->shared_data->tokens, (const int*)pPos);
    for (int l = 0; l < p->n_layers; l++) {
        PerLayerWeight& L = w->layers[l];
        lq4_rmsnorm(s->xb, x, L.rms_att_weight, dim);
        const int loff = l * p->seq_len * kv_dim;
        if (dim == kv_dim) {
            lq4_qkv_matvec(s->q, s->key_cache, s->value_cache, s->xb, &L.wq_q, &L.wq_k, &L.wq_v, dim, dim, loff, pPos);
        } else {
            lq4_matmul_q4(s->q, s->xb, &L.wq_q, dim, dim, 0, -1, nullptr);
            lq4_matmul_q4(s->key_cache, s->xb, &L.wq_k, dim, kv_dim, 0, loff, pPos);
            lq4_matmul_q4(s->value_cache, s->xb, &L.wq_v, dim, kv_dim, 0, loff, pPos);
        }
        lq4_rope_rotation(s->q, s->key_cache, p->n_heads, p->n_kv_heads, head_size, pPos, loff, p->rope_theta);
        lq4_multi_head_attention(s->xb, s->q, s->key_cache + loff, s->value_cache + loff, s->att, p->n_heads,
                                 head_size, kv_mul, seq_len_bin, pPos);
        lq4_matmul_q4(s->x, s->xb, &L.wq_o, dim, dim, 1, -1, nullptr);
        lq4_rmsnorm(s->xb, x, L.rms_ffn_weight, dim);
        lq4_ffn_matvec_silu(s->hb, s->xb, &L.wq_gate, &L.wq_up, dim, hidden_dim);
        lq4_matmul_q4(s->x, s->hb, &L.wq_down, hidden_dim, dim, 1, -1, nullptr);
    }
    lq4_rmsnorm(x, x, w->rms_final_weight, dim);
    lq4_matmul_fp16(s->logits, x, w->wcls, p->dim, p->vocab_size, 1, 0, 0, 0, -1, 1.0f);
}

// Builds (once per RunState) the op table of a whole decode step:
//   per layer  [embed+]RMSNorm+q|k|v  ->  RoPE+attention  ->  o+residual  ->  RMSNorm+gate/up+SiLU  ->  down+residual
//   then       RMSNorm+classifier  ->  argmax
static NetPlan& get_net_plan(Config* p, RunState* s, TransformerWeights* w) {
    std::vector<char> key(sizeof(Config) + sizeof(RunState) + sizeof(TransformerWeights) + 2 * sizeof(int));
    memcpy(key.data() + sizeof(Config) + sizeof(RunState) + sizeof(TransformerWeights), &g.tp_rank, sizeof(int));
    memcpy(key.data() + sizeof(Config) + sizeof(RunState) + sizeof(TransformerWeights) + sizeof(int), &g.tp_world, sizeof(int));
    memcpy(key.data(), p, sizeof(Config));
    memcpy(key.data() + sizeof(Config), s, sizeof(RunState));
    memcpy(key.data() + sizeof(Config) + sizeof(RunState), w, sizeof(TransformerWeights));
    NetPlan& np = g.nets[(const void*)s];
    if (np.key == key) return np;
    if (np.d_ops) { LQ4_CHECK(cudaStreamSynchronize(g.stream)); cudaFree(np.d_ops); cudaFree(np.tagged); cudaFree(np.att_sc); cudaFree(np.att_flags); np.d_ops = nullptr; np.tagged = nullptr; np.att_sc = nullptr; np.att_flags = nullptr; }
    np.key = key;
    np.ok = false;

    const int dim = p->dim, hidden = p->hidden_dim;
    const int head_size = dim / p->n_heads;
    const int kv_dim = (p->dim * p->n_kv_heads) / p->n_heads;
    const int kv_mul = p->n_heads / p->n_kv_heads;
    if (head_size != 32 && head_size != 64 && head_size != 128) return np;
    if (dim > 1024 * kNormMaxT || dim % 64) return np;      // fused RMSNorm staging (interp_sm100.cuh, stage_norm)
    // Tensor parallel (T ranks, one process per GPU): every matrix is split by output columns, rank r owning the r-th
    // contiguous slice (heads r*H/T.. for q|k|v and attention, hidden and dim slices for the FFN and the projections, vocabulary
    // rows for the classifier).  Each column is still computed entirely by one thread in the reference's order, so results are
    // bit-identical to one GPU.  The slices are expressed by offsetting the pointers: the kernel sees smaller matrices.
    const int T = g.tp_world, R = g.tp_rank;
    np.world = T; np.rank = R;
    np.tokens = s->shared_data->tokens;
    if (T > 1 && (p->n_heads % T || p->n_kv_heads % T || (dim / T) % 8 || (hidden / T) % 8 || (kv_dim / T) % 8 || (p->vocab_size / T) % 8 ||
                  p->vocab_size % T || T > 8)) return np;
    const int sdim = dim / T, shid = hidden / T, skv = kv_dim / T, sheads = p->n_heads / T, svoc = p->vocab_size / T;
    const int bc = (T > 1) ? 1 : 0;
    // activations of the fused step travel as tagged 32-bit words (interp_sm100.cuh, "flag-in-data"): no grid barrier
    // between the op that writes a vector and the op that reads it; under tensor parallelism the writer stores into every
    // rank's copy and each rank polls its own
    const size_t ntag = (size_t)3 * dim + hidden + 2 * kv_dim + 32;      // ... + 3 candidate words per rank
    LQ4_CHECK(cudaMalloc((void**)&np.tagged, ntag * sizeof(uint32_t)));
    LQ4_CHECK(cudaMemset(np.tagged, 0, ntag * sizeof(uint32_t)));
    for (int r = 0; r < 8; r++) np.peers[r] = nullptr;
    np.peers[R] = np.tagged;
    uint32_t* xt = np.tagged;
    uint32_t* xbt = xt + dim;
    uint32_t* hbt = xbt + dim;
    uint32_t* qt = hbt + hidden;
    uint32_t* krawt = qt + dim;
    uint32_t* vrawt = krawt + kv_dim;
    uint32_t* cand = vrawt + kv_dim;
    const float2* rope_tab = rope_table(p->rope_theta, head_size, p->seq_len);
    // split attention (long contexts): fp16 score exchange + per-part flags for this rank's heads
    LQ4_CHECK(cudaMalloc((void**)&np.att_sc, (size_t)sheads * p->seq_len * sizeof(uint16_t)));
    LQ4_CHECK(cudaMalloc((void**)&np.att_flags, (size_t)sheads * 4 * sizeof(unsigned)));
    LQ4_CHECK(cudaMemset(np.att_flags, 0, (size_t)sheads * 4 * sizeof(unsigned)));

    std::vector<Op> ops;
    bool ok = true;
    for (int l = 0; l < p->n_layers && ok; l++) {
        PerLayerWeight& L = w->layers[l];
        const int loff = l * p->seq_len * kv_dim;
        const QWeight wq = slice_cols(L.wq_q, dim, R * sdim), wk = slice_cols(L.wq_k, dim, R * skv), wv = slice_cols(L.wq_v, dim, R * skv);
        const QWeight wo = slice_cols(L.wq_o, dim, R * sdim), wg = slice_cols(L.wq_gate, dim, R * shid), wu = slice_cols(L.wq_up, dim, R * shid);
        const QWeight wd = slice_cols(L.wq_down, hidden, R * sdim);
        {   // RMSNorm + q | k | v of this rank's heads (first layer: + embedding gather); consumed by this rank's attention only
            Op op; memset(&op, 0, sizeof op);
            op.kind = OP_GEMV;
            op.x = s->x; op.norm_w = L.rms_att_weight;
            if (l == 0) { op.emb = w->token_embedding_table; op.tokens = s->shared_data->tokens; }
            else op.xt = xt;
            set_seg(op.seg[0], &wq, nullptr, sdim, 0, 0, qt + R * sdim);
            set_seg(op.seg[1], &wk, nullptr, skv, 0, 0, krawt + R * skv);              // rotated into the cache by OP_ATTN
            set_seg(op.seg[2], &wv, s->value_cache + R * skv, skv, loff, kv_dim, vrawt + R * skv);
            const int nc[3] = {sdim, skv, skv};
            ok = ok && q4_op_shape(op, dim, nc, 3, false);
            ops.push_back(op);
        }
        {   // RoPE + attention over this rank's heads; the output slice goes to every rank
            Op op; memset(&op, 0, sizeof op);
            fill_attn_op(op, nullptr, s->q, s->key_cache + loff + R * skv, s->value_cache + loff + R * skv, nullptr, sheads, head_size, kv_mul,
                         p->seq_len);
            op.kv_stride = kv_dim;
            // run_transformer's graph bins (llama2_q4.cu:354-360) hand MultiHeadAttention a max_seq_len above 8192 exactly when
            // pos + 1 > 8192 (bins are 128 ... 8192, then the model's seq_len): the softmax variant depends on the position only
            op.exp16_from = MAX_SEQ_LEN_SMEM_KERNEL;
            op.att_split = 4; op.att_sc = np.att_sc; op.att_sc_stride = p->seq_len; op.att_flags = np.att_flags;
            op.qt = qt + R * sdim; op.krawt = krawt + R * skv; op.vrawt = vrawt + R * skv;
            op.attn_out32 = xbt + R * sdim; op.attn_bcast = bc; op.rope_tab = rope_tab;
            ops.push_back(op);
        }
        if (T > 1 && (g.opt_tp_repl_o < 0 ? T >= 8 : g.opt_tp_repl_o != 0)) {   // o + residual computed by EVERY rank in full: 8.7 MB more to stream per layer, one cross-GPU hand-over less
            Op op; memset(&op, 0, sizeof op);
            op.kind = OP_GEMV; op.x = s->xb; op.xt = xbt; op.accum = 1;
            if (l == 0) { op.res_emb = w->token_embedding_table; op.res_stride = dim; op.tokens = s->shared_data->tokens; }
            set_seg(op.seg[0], &L.wq_o, nullptr, dim, 0, 0, xt, 0);
            ok = ok && q4_op_shape(op, dim, &dim, 1, false);
            ops.push_back(op);
        } else {   // o + residual: this rank's slice of x, broadcast
            Op op; memset(&op, 0, sizeof op);
            op.kind = OP_GEMV; op.x = s->xb; op.xt = xbt; op.accum = 1;
            if (l == 0) { op.res_emb = w->token_embedding_table + R * sdim; op.res_stride = dim; op.tokens = s->shared_data->tokens; }
            set_seg(op.seg[0], &wo, nullptr, sdim, 0, 0, xt + R * sdim, bc);
            ok = ok && q4_op_shape(op, dim, &sdim, 1, false);
            ops.push_back(op);
        }
        {   // RMSNorm + gate/up + SiLU: this rank's slice of hb, broadcast
            Op op; memset(&op, 0, sizeof op);
            op.kind = OP_FFN; op.x = s->x; op.xt = xt; op.norm_w = L.rms_ffn_weight;
            set_seg(op.seg[0], &wg, nullptr, shid, 0, 0, hbt + R * shid, bc);
            set_seg(op.seg[1], &wu, nullptr, shid, 0, 0, hbt + R * shid, bc);
            const int nc[2] = {shid, shid};
            ok = ok && q4_op_shape(op, dim, nc, 2, true);
            ops.push_back(op);
        }
        {   // down + residual
            Op op; memset(&op, 0, sizeof op);
            op.kind = OP_GEMV; op.x = s->hb; op.xt = hbt; op.accum = 1;
            set_seg(op.seg[0], &wd, nullptr, sdim, 0, 0, xt + R * sdim, bc);
            ok = ok && q4_op_shape(op, hidden, &sdim, 1, false);
            ops.push_back(op);
        }
    }
    {   // final RMSNorm + classifier rows of this rank; the logits stay plain fp16 (the API exposes them), so the sampler keeps its barrier
        Op op; memset(&op, 0, sizeof op);
        op.kind = OP_CLS; op.x = s->x; op.xt = xt; op.norm_w = w->rms_final_weight;
        op.seg[0].w = reinterpret_cast<const uint32_t*>(w->wcls + (size_t)R * svoc * dim);
        op.seg[0].out = s->logits + R * svoc;
        op.alpha = 1.0f;
        ok = ok && cls_op_shape(op, dim, svoc, dim);
        ops.push_back(op);
    }
    {
        Op op; memset(&op, 0, sizeof op);
        op.kind = OP_ARGMAX; op.sync_before = 1;
        op.logits = s->logits + R * svoc; op.vocab = svoc; op.vocab0 = R * svoc;
        op.cand = (T > 1) ? cand : nullptr;
        op.tokens_out = &(s->shared_data->tokens[0]);
        op.pos_host = &(s->shared_data->pos);
        op.pos_dev = s->pos;
        op.write_token = 1;
        ops.push_back(op);
    }
    if (!ok) return np;

    // scale/zero buffers alternate per INT4 op: q|k|v and gate/up land in buffer 0, o and down in buffer 1, so each is sized
    // for its own ops only
    int xs = attn_scratch_bytes(head_size, p->seq_len), meta[2] = {0, 0}, nq4 = 0;
    for (auto& op : ops) {
        if (op.kind > OP_CLS) continue;
        xs = std::max(xs, op_xs_bytes(op));
        if (op.kind != OP_CLS) { meta[nq4 & 1] = std::max(meta[nq4 & 1], op_meta_bytes(op, g.sm_count)); nq4++; }
    }
    const int nwc = std::max(default_nwc(), kNormThreads / 32);      // stage_norm needs 128 consumer threads
    if (!make_plan(np.plan, nwc, xs, meta[0], meta[1])) return np;
    for (auto& op : ops)
        if (op.kind <= OP_CLS && !op_set_chunking(op, np.plan.ring_bytes)) return np;
    // the persistent kernel needs one co-resident CTA per SM
    int per_sm = 0;
    for (int f = 0; f < kNumInterpInstances; f++) {
        LQ4_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, interp_instance(f), 32 * (nwc + 1), np.plan.smem));
        if (per_sm < 1) return np;
    }
    np.nops = (int)ops.size();
    LQ4_CHECK(cudaMalloc((void**)&np.d_ops, sizeof(Op) * ops.size()));
    LQ4_CHECK(cudaMemcpy(np.d_ops, ops.data(), sizeof(Op) * ops.size(), cudaMemcpyHostToDevice));
    np.ok = true;
    return np;
}

// one persistent launch: the forward pass, and the greedy sampler too when with_argmax
static bool run_network_fused(int* pPos, Config* p, RunState* s, TransformerWeights* w, bool with_argmax, int write_token, int seq_len) {
    NetPlan& np = get_net_plan(p, s, w);
    if (!np.ok) return false;
    if (pPos != s->pos) {
        fprintf(stderr, "lq4: run_llama_network expects pPos == RunState::pos\n");
        exit(EXIT_FAILURE);
    }
    launch_interp(np.plan, np.d_ops, with_argmax ? np.nops : np.nops - 1, nullptr, pPos, write_token, true, g.sm_count, &np, seq_len > kAttnSplitFrom);
    return true;
}

// Tensor parallel: only the fused greedy step is complete on every rank.  Its classifier writes this rank's vocabulary slice
// of the logits only, and its sampler is the launch's one cross-rank rendezvous (every rank has finished polling the
// activation buffers before any rank leaves it).  Anything that needs the full logits on one rank -- temperature / top-p
// sampling, perplexity (copyLogits), run_llama_network on its own, lq4_step with a logits buffer -- is refused, loudly.
[[noreturn]] static void tp_unsupported(const char* what) {
    fprintf(stderr, "lq4: %s is not available under tensor parallelism (rank %d of %d): only the fused greedy decode step is; "
                    "run it on one GPU\n", what, g.tp_rank, g.tp_world);
    exit(EXIT_FAILURE);
}

void lq4_run_llama_network(int* pPos, Config* p, RunState* s, TransformerWeights* w, int seq_len_bin) {
    ensure_init();
    if (g.tp_world > 1) tp_unsupported("run_llama_network without the fused sampler");
    if (g.opt_fused && run_network_fused(pPos, p, s, w, false, -1, seq_len_bin)) return;      // the bin bounds the position from above
    run_network_unfused(pPos, p, s, w, seq_len_bin);
}

void lq4_build_sampler(Sampler* sampler, int vocab_size, float temperature, float topp, unsigned long long rng_seed) {
    ensure_init();
    memset(sampler, 0, sizeof *sampler);
    sampler->vocab_size = vocab_size;
    sampler->temperature = temperature;
    sampler->topp = topp;
    sampler->rng_state = rng_seed;
    LQ4_CHECK(cudaMalloc((void**)&sampler->indices, vocab_size * sizeof(int)));   // sampler.h:22
}

void lq4_destroy_sampler(Sampler* sampler) {
    cudaFree(sampler->indices);
    cudaFree(sampler->tempStorage_sort);
    cudaFree(sampler->tempStorage_scan);
}

static unsigned int random_u32(unsigned long long* state) {   // sampler.h:31-37 (xorshift*)
    *state ^= *state >> 12;
    *state ^= *state << 25;
    *state ^= *state >> 27;
    return (unsigned int)((*state * 0x2545F4914F6CDD1Dull) >> 32);
}

static float random_f32(unsigned long long* state) { return (random_u32(state) >> 8) / 16777216.0f; }   // sampler.h:38-40

static bool is_greedy(const Sampler* sampler, int gen_token) { return sampler->temperature == 0.0f || !gen_token; }

// the non-greedy branch of sample() (sampler.h:51-80) with an already drawn coin
static void sample_nongreedy(Sampler* sampler, RunState* s, float coin, cudaStream_t st) {
    const int n = sampler->vocab_size;
    temperature_softmax_kernel<<<1, kSampThreads, 0, st>>>(s->logits, n, sampler->temperature, sampler->indices);
    float threshold = 0.0f;
    if (sampler->topp <= 0 || sampler->topp >= 1) {
        threshold = coin;
    } else {
        if (sampler->temp_storage_bytes_sort == 0) {
            cub::DeviceRadixSort::SortPairsDescending(sampler->tempStorage_sort, sampler->temp_storage_bytes_sort, s->logits, s->logits,
                                                      sampler->indices, sampler->indices, n, 0, sizeof(half) * 8, st);
            LQ4_CHECK(cudaMalloc(&sampler->tempStorage_sort, sampler->temp_storage_bytes_sort));
        }
        cub::DeviceRadixSort::SortPairsDescending(sampler->tempStorage_sort, sampler->temp_storage_bytes_sort, s->logits, s->logits,
                                                  sampler->indices, sampler->indices, n, 0, sizeof(half) * 8, st);
        threshold = coin * sampler->topp;
    }
    if (sampler->temp_storage_bytes_scan == 0) {
        cub::DeviceScan::InclusiveSum(sampler->tempStorage_scan, sampler->temp_storage_bytes_scan, s->logits, s->logits, n, st);
        LQ4_CHECK(cudaMalloc(&sampler->tempStorage_scan, sampler->temp_storage_bytes_scan));
    }
    cub::DeviceScan::InclusiveSum(sampler->tempStorage_scan, sampler->temp_storage_bytes_scan, s->logits, s->logits, n, st);
    threshold_pick_kernel<<<1, kSampThreads, 0, st>>>(s->logits, sampler->indices, n, threshold, &(s->shared_data->tokens[0]), &(s->shared_data->pos), s->pos);
}

void lq4_sample(Sampler* sampler, RunState* s, int gen_token, void* cuda_stream) {
    ensure_init();
    if (g.tp_world > 1) tp_unsupported("sample() on the full logits");
    const float coin = random_f32(&sampler->rng_state);   // one draw per step, greedy or not (sampler.h:45)
    if (is_greedy(sampler, gen_token)) {
        argmax_kernel<<<1, 1024, 0, (cudaStream_t)cuda_stream>>>(s->logits, sampler->vocab_size, &(s->shared_data->tokens[0]),
                                                                 &(s->shared_data->pos), s->pos, nullptr, gen_token != 0);
    } else {
        sample_nongreedy(sampler, s, coin, (cudaStream_t)cuda_stream);
    }
}

// run_transformer's length bin for a sequence of seq_len positions (llama2_q4.cu:354-360): 128, 256, ... 8192, else the model's
// seq_len.  It sizes the QK grid there; here it only selects the softmax variant of the op-by-op path (bins above 8192).
static int seq_len_bin_of(const Config* p, int seq_len) {
    int bin = 128;
    for (int i = 0; i < 7; i++, bin *= 2)
        if (seq_len <= bin) return bin;
    return p->seq_len;
}

// forward + sample; seq_len (= pos + 1) only matters to the op-by-op path
static void forward_and_sample(int gen_token, Config* p, RunState* s, TransformerWeights* w, int copyLogits,
                               Sampler* pSampler, int seq_len) {
    const int seq_len_bin = seq_len_bin_of(p, seq_len);
    if (g.opt_fused && !copyLogits && is_greedy(pSampler, gen_token)) {
        if (run_network_fused(s->pos, p, s, w, true, gen_token != 0, seq_len)) {      // greedy sampler = last op of the persistent kernel
            (void)random_u32(&pSampler->rng_state);
            return;
        }
    }
    if (g.tp_world > 1) tp_unsupported(copyLogits ? "perplexity mode (copyLogits)" : is_greedy(pSampler, gen_token) ? "the op-by-op path (fused=0 or an unsupported shape)" : "temperature / top-p sampling");
    lq4_run_llama_network(s->pos, p, s, w, seq_len_bin);
    if (copyLogits) {                                  // llama2_q4.cu:377-382 (perplexity mode)
        float* pOutput = s->logits_array + (size_t)p->vocab_size * s->shared_data->pos;
        convert_fp16_to_fp32_kernel<<<divUp(p->vocab_size, 128), 128, 0, g.stream>>>(pOutput, s->logits, p->vocab_size);
    }
    lq4_sample(pSampler, s, gen_token, g.stream);
}

void lq4_run_transformer(int gen_token, Config* p, RunState* s, TransformerWeights* w, int copyLogits,
                         Sampler* pSampler) {
    ensure_init();
    // the reference picks a CUDA graph by length bin here (llama2_q4.cu:354-372); this engine's step is a
    // single launch whose work depends only on the device-side position, so there is nothing to select
    forward_and_sample(gen_token, p, s, w, copyLogits, pSampler, s->shared_data->pos + 1);   // the host copy of the position, like the reference (:354)
}

// ---------------------------------------------------------------------------------- loader
static size_t qweight_bytes(size_t K, size_t N, size_t* wb, size_t* zb, size_t* sb) {
    const size_t pwh = (size_t)divUp((int)K, 32) * 4, G = (size_t)divUp((int)K, 128), zh = (size_t)divUp((int)G, 8);
    *wb = pwh * N * 4; *zb = zh * N * 4; *sb = G * N * 2;
    return *wb + *zb + *sb;
}

int lq4_build_transformer(Transformer* t, const char* checkpoint_path, int perplexity) {
    ensure_init();
    FILE* file = fopen(checkpoint_path, "rb");
    if (!file) { printf("Couldn't open file %s\n", checkpoint_path); exit(1); }
    if (fread(&t->config, sizeof(Config), 1, file) != 1) { printf("Invalid header size\n"); exit(1); }
    Config* p = &t->config;
    printf("\nModel params:- \ndim: %d \nhidden_dim: %d\nn_heads: %d\nn_kv_heads: %d\nn_layers: %d\nseq_len: %d\nvocab_size: %d\nrope_theta: %g\n",
           p->dim, p->hidden_dim, p->n_heads, p->n_kv_heads, p->n_layers, p->seq_len, p->vocab_size, p->rope_theta);
    const size_t dim = p->dim, hidden = p->hidden_dim, vocab = p->vocab_size;
    const size_t kv_dim = (size_t)(p->dim * p->n_kv_heads) / p->n_heads;

    // The file is one contiguous run of tensors whose sizes are all multiples of 16 bytes, so the whole
    // payload goes into ONE device arena with the file's own layout (B1) and the structs point into it.
    fseek(file, 0, SEEK_END);
    const size_t file_size = (size_t)ftell(file);
    fseek(file, sizeof(Config), SEEK_SET);
    const size_t payload = file_size - sizeof(Config);
    size_t wb, zb, sb;
    size_t expect = vocab * dim * 4 + dim * 2;
    expect += (size_t)p->n_layers * (2 * qweight_bytes(dim, dim, &wb, &zb, &sb) + 2 * qweight_bytes(dim, kv_dim, &wb, &zb, &sb) +
                                     2 * qweight_bytes(dim, hidden, &wb, &zb, &sb) + qweight_bytes(hidden, dim, &wb, &zb, &sb) + dim * 4);
    if (expect != payload) { printf("error reading weights"); exit(EXIT_FAILURE); }   // llama2_q4.cu:158
    uint8_t* arena = nullptr;
    LQ4_CHECK(cudaMalloc((void**)&arena, payload));
    if (!arena) { printf("malloc failed!\n"); exit(EXIT_FAILURE); }

    printf("\nLoading Weights... ");
    fflush(stdout);
    {   // double-buffered pinned staging: fread of chunk i+1 overlaps the H2D copy of chunk i
        const size_t chunk = 64u << 20;
        uint8_t* stage[2];
        cudaEvent_t done[2];
        for (int i = 0; i < 2; i++) { LQ4_CHECK(cudaMallocHost((void**)&stage[i], chunk)); LQ4_CHECK(cudaEventCreate(&done[i])); }
        size_t off = 0;
        int b = 0;
        while (off < payload) {
            const size_t n = std::min(chunk, payload - off);
            LQ4_CHECK(cudaEventSynchronize(done[b]));
            if (fread(stage[b], 1, n, file) != n) { printf("error reading weights"); exit(EXIT_FAILURE); }
            LQ4_CHECK(cudaMemcpyAsync(arena + off, stage[b], n, cudaMemcpyHostToDevice, g.stream));
            LQ4_CHECK(cudaEventRecord(done[b], g.stream));
            off += n;
            b ^= 1;
        }
        LQ4_CHECK(cudaStreamSynchronize(g.stream));
        for (int i = 0; i < 2; i++) { cudaFreeHost(stage[i]); cudaEventDestroy(done[i]); }
    }
    fclose(file);

    // carve (checkpoint_init_weights, llama2_q4.cu:180-197)
    TransformerWeights* w = &t->weights;
    uint8_t* cur = arena;
    auto take = [&](size_t bytes) { uint8_t* r = cur; cur += bytes; return r; };
    auto take_q = [&](QWeight* q, size_t K, size_t N) {
        qweight_bytes(K, N, &wb, &zb, &sb);
        q->weight = (uint32_t*)take(wb);
        q->zeros = (uint32_t*)take(zb);
        q->scales = (half*)take(sb);
    };
    w->token_embedding_table = (half*)take(vocab * dim * 2);
    w->wcls = (half*)take(vocab * dim * 2);
    w->rms_final_weight = (half*)take(dim * 2);
    w->layers = (PerLayerWeight*)malloc(p->n_layers * sizeof(PerLayerWeight));
    w->num_layers = p->n_layers;
    for (int l = 0; l < p->n_layers; l++) {
        PerLayerWeight* L = &w->layers[l];
        take_q(&L->wq_q, dim, dim);
        take_q(&L->wq_k, dim, kv_dim);
        take_q(&L->wq_v, dim, kv_dim);
        take_q(&L->wq_o, dim, dim);
        take_q(&L->wq_up, dim, hidden);      // up before gate in the file (llama2_q4.cu:191-192)
        take_q(&L->wq_gate, dim, hidden);
        take_q(&L->wq_down, hidden, dim);
        L->rms_att_weight = (half*)take(dim * 2);
        L->rms_ffn_weight = (half*)take(dim * 2);
    }
    printf("done!\n");
    g.arenas[(const void*)t] = arena;

    // malloc_run_state (llama2_q4.cu:38-67); att is sized for seq_len rows (the reference's n_heads*dim
    // overflows once seq_len > dim)
    RunState* s = &t->state;
    memset(s, 0, sizeof *s);
    const size_t att_elems = (size_t)p->n_heads * std::max((size_t)p->dim, (size_t)p->seq_len);
    LQ4_CHECK(cudaMalloc((void**)&s->x, dim * sizeof(half)));
    LQ4_CHECK(cudaMalloc((void**)&s->xb, dim * sizeof(half)));
    LQ4_CHECK(cudaMalloc((void**)&s->hb, hidden * sizeof(half)));
    LQ4_CHECK(cudaMalloc((void**)&s->q, dim * sizeof(half)));
    LQ4_CHECK(cudaMalloc((void**)&s->att, att_elems * sizeof(half)));
    LQ4_CHECK(cudaMalloc((void**)&s->logits, vocab * sizeof(half)));
    LQ4_CHECK(cudaMalloc((void**)&s->key_cache, sizeof(half) * p->n_layers * (size_t)p->seq_len * kv_dim));
    LQ4_CHECK(cudaMalloc((void**)&s->value_cache, sizeof(half) * p->n_layers * (size_t)p->seq_len * kv_dim));
    LQ4_CHECK(cudaMalloc((void**)&s->pos, sizeof(int)));
    LQ4_CHECK(cudaMallocHost((void**)&s->shared_data, sizeof(SharedData)));
    LQ4_CHECK(cudaMemset(s->pos, 0, sizeof(int)));
    LQ4_CHECK(cudaMemset(s->key_cache, 0, sizeof(half) * p->n_layers * (size_t)p->seq_len * kv_dim));
    LQ4_CHECK(cudaMemset(s->value_cache, 0, sizeof(half) * p->n_layers * (size_t)p->seq_len * kv_dim));
    s->shared_data->pos = 0;
    if (perplexity) LQ4_CHECK(cudaMalloc((void**)&s->logits_array, sizeof(float) * (size_t)p->seq_len * vocab));
    (void)rope_table(p->rope_theta, p->dim / p->n_heads, p->seq_len);
    LQ4_CHECK(cudaDeviceSynchronize());   // the memsets above ran on the legacy stream
    (void)get_net_plan(p, s, w);          // op table of the decode step
    return 0;
}

void lq4_free_transformer(Transformer* t) {
    RunState* s = &t->state;
    cudaStreamSynchronize(g.stream);
    auto np = g.nets.find((const void*)s);
    if (np != g.nets.end()) {
        for (int r = 0; r < np->second.world; r++)
            if (r != np->second.rank && np->second.peers[r] != nullptr) cudaIpcCloseMemHandle(np->second.peers[r]);
        cudaFree(np->second.d_ops); cudaFree(np->second.tagged); cudaFree(np->second.att_sc); cudaFree(np->second.att_flags); g.nets.erase(np);
    }
    cudaFree(s->x); cudaFree(s->xb); cudaFree(s->pos); cudaFree(s->hb); cudaFree(s->q); cudaFree(s->att);
    cudaFree(s->logits); cudaFree(s->key_cache); cudaFree(s->value_cache); cudaFreeHost(s->shared_data);
    if (s->logits_array) cudaFree(s->logits_array);
    auto it = g.arenas.find((const void*)t);
    if (it != g.arenas.end()) { cudaFree(it->second); g.arenas.erase(it); }
    free(t->weights.layers);
    memset(t, 0, sizeof *t);
}

// ---------------------------------------------------------------------------------- step driver
void lq4_reset(Transformer* t, const int* tokens, int n) {   // llama2_q4.cu:461-463
    ensure_init();
    LQ4_CHECK(cudaMemsetAsync(t->state.pos, 0, sizeof(int), g.stream));
    LQ4_CHECK(cudaStreamSynchronize(g.stream));
    t->state.shared_data->pos = 0;
    memcpy((void*)t->state.shared_data->tokens, tokens, sizeof(int) * n);
}

int lq4_step(Transformer* t, Sampler* sampler, int gen_token, half* logits_out, int* next_token_out) {
    ensure_init();
    if (g.tp_world > 1 && logits_out) tp_unsupported("lq4_step with a logits buffer");
    LQ4_CHECK(cudaStreamSynchronize(g.stream));
    lq4_run_transformer(gen_token, &t->config, &t->state, &t->weights, 0, sampler);
    LQ4_CHECK(cudaStreamSynchronize(g.stream));
    const int pos = t->state.shared_data->pos;
    if (logits_out)
        LQ4_CHECK(cudaMemcpy(logits_out, t->state.logits, sizeof(half) * t->config.vocab_size, cudaMemcpyDeviceToHost));
    if (next_token_out) *next_token_out = t->state.shared_data->tokens[pos];
    return pos;
}

// Enqueue one forward + sample without touching the host copy of the position.  seq_len (= pos+1) is what
// the reference's graph-bin selection needs (llama2_q4.cu:354-360); only the op-by-op path uses it here.
void lq4_enqueue_step(Transformer* t, Sampler* sampler, int seq_len, int gen_token) {
    ensure_init();
    forward_and_sample(gen_token, &t->config, &t->state, &t->weights, 0, sampler, std::max(seq_len, 1));
}

int lq4_generate_tokens(Transformer* t, Sampler* sampler, const int* prompt_tokens, int n_prompt, int steps,
                        int* out_tokens, double* seconds, int pipelined) {
    ensure_init();
    if (n_prompt < 1) { fprintf(stderr, "something is wrong, expected at least 1 prompt token\n"); exit(EXIT_FAILURE); }
    if (steps <= 0 || steps > t->config.seq_len) steps = t->config.seq_len;   // llama2_q4.cu:690
    Config* p = &t->config;
    RunState* s = &t->state;
    const long long start = time_in_ns();
    LQ4_CHECK(cudaMemsetAsync(s->pos, 0, sizeof(int), g.stream));
    LQ4_CHECK(cudaStreamSynchronize(g.stream));
    s->shared_data->pos = 0;
    memcpy((void*)s->shared_data->tokens, prompt_tokens, sizeof(int) * n_prompt);
    int pos = 0;
    if (out_tokens) out_tokens[0] = prompt_tokens[0];
    const int eos = 2;
    if (!pipelined) {
        while (pos < steps) {                        // llama2_q4.cu:465-482
            LQ4_CHECK(cudaStreamSynchronize(g.stream));
            lq4_run_transformer(pos >= n_prompt - 1, p, s, &t->weights, 0, sampler);
            if (pos > 0) {
                int next = s->shared_data->tokens[pos];
                if (next >= p->vocab_size) next = 0;
                if (out_tokens) out_tokens[pos] = next;
                if (next == eos) break;
            }
            pos++;
        }
    } else {
        // The position and the sampled token live on the device (the sampler writes both), so step
        // pos+1 can be enqueued before step pos has finished; the host only trails behind to read
        // tokens.  Events mark the end of each step.
        cudaEvent_t ev[2];
        LQ4_CHECK(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
        LQ4_CHECK(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
        int launched = 0;
        bool stop = false;
        while (pos < steps && !stop) {
            while (launched < steps && launched <= pos + 1) {
                forward_and_sample(launched >= n_prompt - 1, p, s, &t->weights, 0, sampler, launched + 1);
                LQ4_CHECK(cudaEventRecord(ev[launched & 1], g.stream));
                launched++;
            }
            LQ4_CHECK(cudaEventSynchronize(ev[pos & 1]));        // step `pos` finished: tokens[pos+1] valid
            if (pos > 0) {
                int next = s->shared_data->tokens[pos];
                if (next >= p->vocab_size) next = 0;
                if (out_tokens) out_tokens[pos] = next;
                if (next == eos) stop = true;
            }
            if (!stop) pos++;
        }
        LQ4_CHECK(cudaStreamSynchronize(g.stream));
        cudaEventDestroy(ev[0]);
        cudaEventDestroy(ev[1]);
    }
    LQ4_CHECK(cudaStreamSynchronize(g.stream));
    const long long end = time_in_ns();
    if (seconds) *seconds = (double)(end - start) * 1e-9;
    return pos;
}

// device -> host copy for pure-host callers of the C ABI (the CLI's perplexity mode reads RunState::logits_array)
int lq4_memcpy_to_host(void* dst, const void* src_device, size_t bytes) {
    ensure_init();
    cudaError_t e = cudaStreamSynchronize(g.stream);
    if (e == cudaSuccess) e = cudaMemcpy(dst, src_device, bytes, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { set_err("lq4_memcpy_to_host", e); return 1; }
    return 0;
}

// ---------------------------------------------------------------------------------- batched prefill (scope row f1)
// Not in the reference (prompt tokens go one by one through decode, llama2_q4.cu:465-470).  Every projection of a batch of
// token rows is one dense INT4 -> fp16 GEMM on the tcgen05 tensor cores (prefill_sm100.cuh); parity is fp16 tolerance
// against the sequential decode path, not bit-exactness.
namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !ptr) {
            fprintf(stderr, "lq4: cuTensorMapEncodeTiled is not available from this driver\n");
            exit(EXIT_FAILURE);
        }
        fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

struct PrefillWs {          // activations of a batch of token rows (device), grown on demand
    size_t rows = 0;
    Config cfg = {};
    half *x = nullptr, *xn = nullptr, *q = nullptr, *k = nullptr, *v = nullptr, *att = nullptr, *g = nullptr, *u = nullptr, *xl = nullptr, *logits = nullptr;
    int* tokens = nullptr;
    int batch = 0;
} g_pf;

void pf_free() {
    cudaFree(g_pf.x); cudaFree(g_pf.xn); cudaFree(g_pf.q); cudaFree(g_pf.k); cudaFree(g_pf.v); cudaFree(g_pf.att); cudaFree(g_pf.g);
    cudaFree(g_pf.u); cudaFree(g_pf.xl); cudaFree(g_pf.logits); cudaFree(g_pf.tokens);
    g_pf = PrefillWs();
}

// Y[M][N] = X[M][K] . dequant(W)[N][K]^T (+ res), fp16 in / fp32 accumulate in tensor memory / fp16 out
bool gemm_q4_tc(half* y, const half* x, const QWeight* w, int M, int K, int N, const half* res) {
    if (M < 1 || K % lq4pf::kBK || N % lq4pf::kBNmin || (((uintptr_t)x | (uintptr_t)y | (uintptr_t)res) & 15)) return false;
    CUtensorMap tmap;
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M};
    const cuuint64_t strides[1] = {(cuuint64_t)K * sizeof(half)};
    const cuuint32_t box[2] = {(cuuint32_t)lq4pf::kBK, (cuuint32_t)lq4pf::kBM};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode_tiled()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { fprintf(stderr, "lq4: cuTensorMapEncodeTiled failed (%d)\n", (int)r); exit(EXIT_FAILURE); }
    static bool attr = false;
    if (!attr) {
        LQ4_CHECK(cudaFuncSetAttribute(lq4pf::gemm_q4_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lq4pf::Geo<128>::kSmemBytes));
        LQ4_CHECK(cudaFuncSetAttribute(lq4pf::gemm_q4_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lq4pf::Geo<256>::kSmemBytes));
        attr = true;
    }
    lq4pf::GemmParams p;
    p.w = w->weight; p.z = w->zeros; p.s = reinterpret_cast<const uint16_t*>(w->scales);
    p.y = y; p.res = res; p.M = M; p.N = N; p.K = K; p.ldy = N;
    const int num_m = (M + lq4pf::kBM - 1) / lq4pf::kBM;
    // 256-wide tiles halve the L2 traffic of X; the narrow tile serves N % 256 != 0 and problems too small to fill the SMs otherwise
    const bool wide = (N % 256 == 0) && (num_m * (N / 256) >= g.sm_count || getenv("LQ4_GEMM_WIDE"));
    if (wide) {
        const int tiles = num_m * (N / 256);
        lq4pf::gemm_q4_tc_kernel<256><<<std::min(tiles, g.sm_count), lq4pf::kThreads, lq4pf::Geo<256>::kSmemBytes, g.stream>>>(tmap, p);
    } else {
        const int tiles = num_m * (N / 128);
        lq4pf::gemm_q4_tc_kernel<128><<<std::min(tiles, g.sm_count), lq4pf::kThreads, lq4pf::Geo<128>::kSmemBytes, g.stream>>>(tmap, p);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_err("gemm_q4_tc_kernel launch", e); exit(EXIT_FAILURE); }
    return true;
}
}  // namespace

// Dense INT4 GEMM on the tensor cores: y[M][N] = x[M][K] . W^T (+ y when accum), all device pointers, row-major fp16.
// Returns 1 when the shape is not supported (K % 64, N % 128, 16-byte alignment); nothing else computes it then.
int lq4_gemm_q4(half* y, const half* x, const QWeight* w, int M, int K, int N, int accum) {
    ensure_init();
    return gemm_q4_tc(y, x, w, M, K, N, accum ? y : nullptr) ? 0 : 1;
}

// Batched prefill of `batch` sequences of `seq` tokens each (host ids, [batch][seq]).  logits_last (host, [batch][vocab] fp16,
// may be NULL) receives the logits of every sequence's last position.  kv_seq >= 0: that sequence's rotated K and V rows are
// also written to the transformer's KV cache and the device / host positions are set to seq - 1 with the tokens copied into
// SharedData, so that the decode path continues from there (lq4_step / lq4_run_transformer at position seq - 1 recomputes the
// last prompt position and samples).  ms_total / ms_gemm (may be NULL): device time of the whole pass and of its GEMMs.
// Returns 0, or 1 when a shape is not supported by the tensor-core GEMM.
int lq4_prefill(Transformer* t, const int* tokens, int batch, int seq, int kv_seq, half* logits_last, float* ms_total, float* ms_gemm) {
    ensure_init();
    Config* p = &t->config;
    RunState* s = &t->state;
    TransformerWeights* w = &t->weights;
    const int dim = p->dim, hidden = p->hidden_dim, hs = dim / p->n_heads;
    const int kv_dim = (p->dim * p->n_kv_heads) / p->n_heads, kv_mul = p->n_heads / p->n_kv_heads;
    if (batch < 1 || seq < 1 || seq > p->seq_len || kv_seq >= batch) return 1;
    if (dim % 128 || kv_dim % 128 || hidden % 128 || dim % 64 || hidden % 64 || hs % 32 || hs > 256 || 256 % lq4pf::kPfQ || hs % (256 / lq4pf::kPfQ)) return 1;
    const size_t M = (size_t)batch * seq;
    if (g_pf.rows < M || g_pf.batch < batch || memcmp(&g_pf.cfg, p, sizeof(Config))) {
        LQ4_CHECK(cudaStreamSynchronize(g.stream));
        pf_free();
        g_pf.rows = M; g_pf.cfg = *p; g_pf.batch = batch;
        LQ4_CHECK(cudaMalloc((void**)&g_pf.x, M * dim * sizeof(half)));
        LQ4_CHECK(cudaMalloc((void**)&g_pf.xn, M * dim * sizeof(half)));
        LQ4_CHECK(cudaMalloc((void**)&g_pf.q, M * dim * sizeof(half)));
        LQ4_CHECK(cudaMalloc((void**)&g_pf.k, M * kv_dim * sizeof(half)));
        LQ4_CHECK(cudaMalloc((void**)&g_pf.v, M * kv_dim * sizeof(half)));
        LQ4_CHECK(cudaMalloc((void**)&g_pf.att, M * dim * sizeof(half)));
        LQ4_CHECK(cudaMalloc((void**)&g_pf.g, M * hidden * sizeof(half)));
        LQ4_CHECK(cudaMalloc((void**)&g_pf.u, M * hidden * sizeof(half)));
        LQ4_CHECK(cudaMalloc((void**)&g_pf.xl, (size_t)batch * dim * sizeof(half)));
        LQ4_CHECK(cudaMalloc((void**)&g_pf.logits, (size_t)batch * p->vocab_size * sizeof(half)));
        LQ4_CHECK(cudaMalloc((void**)&g_pf.tokens, M * sizeof(int)));
    }
    PrefillWs& W = g_pf;
    const float2* rope_tab = rope_table(p->rope_theta, hs, p->seq_len);
    LQ4_CHECK(cudaMemcpyAsync(W.tokens, tokens, M * sizeof(int), cudaMemcpyHostToDevice, g.stream));
    cudaEvent_t ev[4];
    for (auto& e : ev) LQ4_CHECK(cudaEventCreate(&e));
    float gemm_ms = 0.0f;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> spans;
    auto gemm = [&](half* y, const half* x, const QWeight* qw, int K, int N, const half* res) {
        cudaEvent_t a = nullptr, b = nullptr;
        if (ms_gemm) { LQ4_CHECK(cudaEventCreate(&a)); LQ4_CHECK(cudaEventCreate(&b)); LQ4_CHECK(cudaEventRecord(a, g.stream)); }
        if (!gemm_q4_tc(y, x, qw, (int)M, K, N, res)) { fprintf(stderr, "lq4: prefill GEMM shape K=%d N=%d is not supported\n", K, N); exit(EXIT_FAILURE); }
        if (ms_gemm) { LQ4_CHECK(cudaEventRecord(b, g.stream)); spans.push_back({a, b}); }
    };
    const size_t attn_smem = sizeof(float) * ((size_t)lq4pf::kPfQ * hs + (size_t)lq4pf::kPfQ * seq);
    if (hs != 128 && hs != 64) {
        if (attn_smem > (size_t)g.max_smem) return 1;
        allow_smem(lq4pf::attn_prefill_kernel, attn_smem);
    } else if (hs == 128) {
        allow_smem(lq4pf::attn_prefill_mma_kernel<128>, lq4pf::fa_smem_bytes<128>());
    } else {
        allow_smem(lq4pf::attn_prefill_mma_kernel<64>, lq4pf::fa_smem_bytes<64>());
    }
    LQ4_CHECK(cudaEventRecord(ev[0], g.stream));
    lq4pf::embed_rows_kernel<<<(unsigned)M, 128, 0, g.stream>>>(W.x, w->token_embedding_table, W.tokens, dim);
    for (int l = 0; l < p->n_layers; l++) {
        PerLayerWeight& L = w->layers[l];
        lq4pf::rmsnorm_rows_kernel<<<(unsigned)M, 256, 0, g.stream>>>(W.xn, W.x, L.rms_att_weight, dim);
        gemm(W.q, W.xn, &L.wq_q, dim, dim, nullptr);
        gemm(W.k, W.xn, &L.wq_k, dim, kv_dim, nullptr);
        gemm(W.v, W.xn, &L.wq_v, dim, kv_dim, nullptr);
        lq4pf::rope_rows_kernel<<<(unsigned)M, 256, 0, g.stream>>>(W.q, W.k, rope_tab, p->n_heads, p->n_kv_heads, hs, seq);
        if (kv_seq >= 0) {
            const size_t loff = (size_t)l * p->seq_len * kv_dim, r0 = (size_t)kv_seq * seq * kv_dim;
            lq4pf::kv_store_kernel<<<seq, 128, 0, g.stream>>>(s->key_cache + loff, s->value_cache + loff, W.k + r0, W.v + r0, kv_dim);
        }
        const float att_alpha = (float)(1.0 / sqrt((double)hs));
        const dim3 fa_grid((seq + lq4pf::kFaQ - 1) / lq4pf::kFaQ, p->n_heads, batch);
        if (hs == 128)
            lq4pf::attn_prefill_mma_kernel<128><<<fa_grid, 128, lq4pf::fa_smem_bytes<128>(), g.stream>>>(W.att, W.q, W.k, W.v, seq, p->n_heads, kv_mul, att_alpha);
        else if (hs == 64)
            lq4pf::attn_prefill_mma_kernel<64><<<fa_grid, 128, lq4pf::fa_smem_bytes<64>(), g.stream>>>(W.att, W.q, W.k, W.v, seq, p->n_heads, kv_mul, att_alpha);
        else        // other head sizes: the CUDA-core kernel
            lq4pf::attn_prefill_kernel<<<dim3((seq + lq4pf::kPfQ - 1) / lq4pf::kPfQ, p->n_heads, batch), 256, attn_smem, g.stream>>>(
                W.att, W.q, W.k, W.v, seq, p->n_heads, kv_mul, hs, att_alpha);
        gemm(W.x, W.att, &L.wq_o, dim, dim, W.x);
        lq4pf::rmsnorm_rows_kernel<<<(unsigned)M, 256, 0, g.stream>>>(W.xn, W.x, L.rms_ffn_weight, dim);
        gemm(W.g, W.xn, &L.wq_gate, dim, hidden, nullptr);
        gemm(W.u, W.xn, &L.wq_up, dim, hidden, nullptr);
        const size_t nh = M * hidden;
        lq4pf::silu_mul_kernel<<<(unsigned)((nh / 8 + 255) / 256), 256, 0, g.stream>>>(W.g, W.g, W.u, nh);
        gemm(W.x, W.g, &L.wq_down, hidden, dim, W.x);
    }
    // last position of every sequence: final RMSNorm + classifier (the decode path's fp16 GEMV, one row per sequence)
    for (int b = 0; b < batch; b++) {
        lq4pf::rmsnorm_rows_kernel<<<1, 256, 0, g.stream>>>(W.xl + (size_t)b * dim, W.x + ((size_t)b * seq + seq - 1) * dim, w->rms_final_weight, dim);
        lq4_matmul_fp16(W.logits + (size_t)b * p->vocab_size, W.xl + (size_t)b * dim, w->wcls, dim, p->vocab_size, 1, 0, 0, 0, -1, 1.0f);
    }
    LQ4_CHECK(cudaEventRecord(ev[1], g.stream));
    if (kv_seq >= 0) {        // hand over to the decode path at the last prompt position
        const int pos = seq - 1;
        LQ4_CHECK(cudaMemcpyAsync(s->pos, &pos, sizeof(int), cudaMemcpyHostToDevice, g.stream));
    }
    LQ4_CHECK(cudaStreamSynchronize(g.stream));
    if (kv_seq >= 0) {
        s->shared_data->pos = seq - 1;
        memcpy((void*)s->shared_data->tokens, tokens + (size_t)kv_seq * seq, sizeof(int) * seq);
    }
    if (ms_total) LQ4_CHECK(cudaEventElapsedTime(ms_total, ev[0], ev[1]));
    for (auto& sp : spans) {
        float ms = 0.0f;
        LQ4_CHECK(cudaEventElapsedTime(&ms, sp.first, sp.second));
        gemm_ms += ms;
        cudaEventDestroy(sp.first); cudaEventDestroy(sp.second);
    }
    if (ms_gemm) *ms_gemm = gemm_ms;
    for (auto& e : ev) cudaEventDestroy(e);
    if (logits_last) LQ4_CHECK(cudaMemcpy(logits_last, W.logits, (size_t)batch * p->vocab_size * sizeof(half), cudaMemcpyDeviceToHost));
    return 0;
}

// ---------------------------------------------------------------------------------- tensor parallel (one process per GPU)
// Rank / world of this process; call before lq4_build_transformer.  Every rank loads the whole .bin (a 7B model is 2 % of a
// B200's memory) and streams only its column slices.
int lq4_tp_config(int rank, int world) {
    ensure_init();
    if (world < 1 || world > 8 || rank < 0 || rank >= world) return 1;
    g.tp_rank = rank; g.tp_world = world;
    return 0;
}
static NetPlan* tp_plan(Transformer* t) {
    NetPlan& np = get_net_plan(&t->config, &t->state, &t->weights);
    return np.ok ? &np : nullptr;
}
// 64-byte CUDA IPC handle of this rank's tagged-activation buffer, to be sent to every other rank
int lq4_tp_export(Transformer* t, void* handle64) {
    ensure_init();
    NetPlan* np = tp_plan(t);
    if (!np) return 1;
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, np->tagged) != cudaSuccess) { set_err("cudaIpcGetMemHandle", cudaGetLastError()); return 1; }
    static_assert(sizeof(h) == 64, "CUDA IPC handles are 64 bytes");
    memcpy(handle64, &h, 64);
    return 0;
}
// map rank `peer`'s buffer (its exported handle) into this process; the stores of the broadcast go through this mapping
int lq4_tp_import(Transformer* t, int peer, const void* handle64) {
    ensure_init();
    NetPlan* np = tp_plan(t);
    if (!np || peer < 0 || peer >= np->world) return 1;
    if (peer == np->rank) { np->peers[peer] = np->tagged; return 0; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { set_err("cudaIpcOpenMemHandle", e); return 1; }
    np->peers[peer] = (uint32_t*)ptr;
    return 0;
}

// ---------------------------------------------------------------------------------- synthetic files
size_t lq4_write_synth_model(const char* path, const Config* cfg, unsigned long long seed) {
    synth::Cfg c;
    memcpy(&c, cfg, sizeof c);
    return synth::write_model(path, c, seed);
}
size_t lq4_write_synth_tokenizer(const char* path, int vocab_size) { return synth::write_tokenizer(path, vocab_size); }

}  // extern "C"

// engine.cu -- host side of the C ABI declared in include/llama_q4_b200.h.
//
// Mirrors the reference's host wrappers (llama2_q4.cu:207-432) one for one, but every launch goes to
// the sm_100a kernels in kernels_sm100.cuh.  No CPU fallback exists: without a CUDA device the calls
// fail loudly.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <map>
#include <string>

#include "lq4_types.h"
#include "kernels_sm100.cuh"
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

constexpr int MAX_GRAPHS = 8;   // llama2_q4.cu:342

struct Engine {
    bool inited = false;
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int opt_pdl = 1;        // programmatic dependent launch between the fused kernels
    int opt_fused = 1;      // run_llama_network: fused 5-kernel layer (1) or op-by-op like the reference (0)
    int opt_graphs = 1;     // llama2_q4.cu:33 USE_CUDA_GRAPHS
    std::map<RopeKey, float2*> rope_tabs;
    cudaGraphExec_t graph_exec[MAX_GRAPHS] = {};
    bool graph_captured[MAX_GRAPHS] = {};
    const void* graph_owner = nullptr;     // RunState the cached graphs were captured for
    std::map<const void*, void*> arenas;   // Transformer* -> device arena (loader)
    char err[512] = {0};
};
Engine g;

void set_err(const char* what, cudaError_t e) {
    snprintf(g.err, sizeof g.err, "%s: %s", what, cudaGetErrorString(e));
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
    if (!g.stream) { LQ4_CHECK(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking)); g.own_stream = true; }
    const char* env;
    if ((env = getenv("LQ4_PDL"))) g.opt_pdl = atoi(env);
    if ((env = getenv("LQ4_FUSED"))) g.opt_fused = atoi(env);
    if ((env = getenv("LQ4_GRAPHS"))) g.opt_graphs = atoi(env);
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

size_t attn_smem(int head_size, int max_seq) {
    return (size_t)(head_size + 32 + 4 + ((max_seq + 3) & ~3) + 32 * head_size) * sizeof(float);
}

void launch_attention(half* out, const half* q, const half* kc, const half* vc, half* att, int num_heads,
                      int head_size, int kv_mul, int max_seq_len, const int* pPos, bool pdl) {
    if (head_size % 32 != 0 || head_size > 256) unsupported();
    if (max_seq_len > MAX_SEQ_LEN_SMEM_KERNEL) {
        // the reference switches to softmax_kernel_no_smem here (llama2_q4.cu:276-279), whose fp16
        // rounding of exp() differs; that variant is scope row f2 and not built yet.
        fprintf(stderr, "lq4: sequence bins above %d are not supported yet\n", MAX_SEQ_LEN_SMEM_KERNEL);
        exit(EXIT_FAILURE);
    }
    AttnParams p;
    p.out = out; p.q = q; p.kcache = kc; p.vcache = vc; p.att_out = att;
    p.head_size = head_size; p.kv_mul = kv_mul; p.kv_stride = (num_heads * head_size) / kv_mul;
    p.pPos = pPos;
    p.alpha = (float)(1.0 / sqrt((double)head_size));     // llama2_q4.cu:273
    p.max_seq = max_seq_len;
    const size_t smem = attn_smem(head_size, max_seq_len);
    allow_smem(attention_kernel, smem);
    launch(attention_kernel, dim3(num_heads), dim3(kAttnThreads), smem, pdl, p);
}

void launch_classifier(half* out, const half* x, const half* norm_w, const half* w, int n, int d, int w_row_stride,
                       float alpha, bool pdl) {
    ClsParams p;
    p.x = x; p.norm_w = norm_w; p.w = w; p.out = out; p.x_norm_out = nullptr;
    p.n = n; p.d = d; p.w_row_stride = w_row_stride; p.alpha = alpha;
    const size_t smem = (size_t)(((n / 2 + 3) & ~3) + 32) * sizeof(float);
    allow_smem(gemv_f16_kernel, smem);
    const int groups = (d + kClsRows - 1) / kClsRows;
    int blocks = (groups + kClsThreads / 32 - 1) / (kClsThreads / 32);
    blocks = std::min(blocks, 2 * g.sm_count);
    launch(gemv_f16_kernel, dim3(blocks), dim3(kClsThreads), smem, pdl, p);
}

long time_in_ms() {   // llama2_q4.cu:400-405
    struct timespec time;
    timespec_get(&time, TIME_UTC);
    return time.tv_sec * 1000 + time.tv_nsec / 1000000;
}

void destroy_graphs() {
    for (int i = 0; i < MAX_GRAPHS; i++)
        if (g.graph_captured[i]) { cudaGraphExecDestroy(g.graph_exec[i]); g.graph_captured[i] = false; }
    g.graph_owner = nullptr;
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
    destroy_graphs();
}
int lq4_stream_synchronize(void) {
    ensure_init();
    cudaError_t e = cudaStreamSynchronize(g.stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { set_err("stream synchronize", e); return 1; }
    return 0;
}
const char* lq4_last_error(void) { return g.err; }
int lq4_sm_count(void) { ensure_init(); return g.sm_count; }
void lq4_set_option(const char* name, int value) {
    ensure_init();
    if (!strcmp(name, "pdl")) g.opt_pdl = value;
    else if (!strcmp(name, "fused")) g.opt_fused = value;
    else if (!strcmp(name, "graphs")) g.opt_graphs = value;
    destroy_graphs();
}

// ---------------------------------------------------------------------------------- operator API
void lq4_rmsnorm(half* o, half* x, half* weight, int size) {
    ensure_init();
    rmsnorm_kernel<<<1, 1024, 0, g.stream>>>(o, x, weight, size);
}

void lq4_matmul_fp16(half* xout, half* x, half* w, int n, int d, int batch, int x_stride, int w_stride,
                     int op_stride, int w_row_stride, float alpha) {
    ensure_init();
    if ((n & 7) || (d & 7)) unsupported();
    if (w_row_stride == -1) w_row_stride = n;
    for (int b = 0; b < batch; b++)   // the reference batches through blockIdx.y; the only caller uses batch 1
        launch_classifier(xout + (size_t)b * op_stride, x + (size_t)b * x_stride, nullptr, w + (size_t)b * w_stride, n, d,
                          w_row_stride, alpha, false);
}

void lq4_matmul_q4(half* xout, half* x, const QWeight* w, int inpSize, int opSize, int accum, int loff, int* pPos) {
    ensure_init();
    check_q4_shape(inpSize, opSize);
    GemvParams p = {};
    p.x = x; p.K = inpSize;
    p.m[0] = qw_view(w); p.n[0] = opSize; p.out[0] = xout;
    p.accum = accum; p.loff = loff; p.pPos = (loff != -1) ? pPos : nullptr;
    launch_gemv<GEMV_PLAIN>(p, opSize / 4, false);
}

void lq4_qkv_matvec(half* q, half* key_cache, half* value_cache, half* x, const QWeight* qw, const QWeight* kw,
                    const QWeight* vw, int inpSize, int opSize, int loff, int* pPos) {
    ensure_init();
    check_q4_shape(inpSize, opSize);
    GemvParams p = {};
    p.x = x; p.K = inpSize;
    p.m[0] = qw_view(qw); p.m[1] = qw_view(kw); p.m[2] = qw_view(vw);
    p.n[0] = p.n[1] = p.n[2] = opSize;
    p.out[0] = q; p.out[1] = key_cache; p.out[2] = value_cache;
    p.loff = loff; p.pPos = pPos;
    launch_gemv<GEMV_QKV>(p, 3 * opSize / 4, false);
}

void lq4_ffn_matvec_silu(half* xout, half* x, const QWeight* gate_w, const QWeight* up_w, int inpSize, int opSize) {
    ensure_init();
    check_q4_shape(inpSize, opSize);
    GemvParams p = {};
    p.x = x; p.K = inpSize;
    p.m[0] = qw_view(gate_w); p.m[1] = qw_view(up_w);
    p.n[0] = opSize; p.out[0] = xout;
    p.loff = -1;
    launch_gemv<GEMV_FFN>(p, opSize / 2, false);
}

void lq4_rope_rotation(half* q, half* k, int num_heads, int num_kv_heads, int head_size, int* pPos, int loff,
                       float rope_theta) {
    ensure_init();
    rope_kernel<<<num_heads, head_size / 2, 0, g.stream>>>(q, k, num_kv_heads, head_size, pPos, loff, rope_theta);
}

void lq4_multi_head_attention(half* output, half* q, half* key_cache, half* value_cache, half* att, int num_heads,
                              int head_size, int kv_mul, int max_seq_len, int* pPos) {
    ensure_init();
    launch_attention(output, q, key_cache, value_cache, att, num_heads, head_size, kv_mul, max_seq_len, pPos, false);
}

// ---------------------------------------------------------------------------------- forward pass
static void run_network_unfused(int* pPos, Config* p, RunState* s, TransformerWeights* w, int seq_len_bin) {
    // op-by-op, exactly the sequence of llama2_q4.cu:286-340
    half* x = s->x;
    const int dim = p->dim, hidden_dim = p->hidden_dim;
    const int head_size = dim / p->n_heads;
    const int kv_dim = (p->dim * p->n_kv_heads) / p->n_heads;
    const int kv_mul = p->n_heads / p->n_kv_heads;
    copy_embedding_kernel<<<divUp(dim, 256), 256, 0, g.stream>>>(x, w->token_embedding_table, dim,
                                                                 s->shared_data->tokens, pPos);
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

static void run_network_fused(int* pPos, Config* p, RunState* s, TransformerWeights* w, int seq_len_bin,
                              const float2* rope_tab) {
    // Same dataflow in 5 kernels per layer: [embed+]RMSNorm+QKV+RoPE | attention | O+residual |
    // RMSNorm+gate/up+SiLU | down+residual, then RMSNorm+classifier.
    const int dim = p->dim, hidden_dim = p->hidden_dim;
    const int head_size = dim / p->n_heads;
    const int kv_dim = (p->dim * p->n_kv_heads) / p->n_heads;
    const int kv_mul = p->n_heads / p->n_kv_heads;
    const bool pdl = g.opt_pdl != 0;
    check_q4_shape(dim, dim);
    check_q4_shape(dim, kv_dim);
    check_q4_shape(dim, hidden_dim);
    check_q4_shape(hidden_dim, dim);
    if (head_size % 4 != 0) unsupported();
    for (int l = 0; l < p->n_layers; l++) {
        PerLayerWeight& L = w->layers[l];
        const int loff = l * p->seq_len * kv_dim;
        {
            GemvParams a = {};
            a.x = s->x; a.norm_w = L.rms_att_weight; a.K = dim;
            if (l == 0) { a.emb_table = w->token_embedding_table; a.tokens = s->shared_data->tokens; a.x_copy = s->x; }
            a.m[0] = qw_view(&L.wq_q); a.m[1] = qw_view(&L.wq_k); a.m[2] = qw_view(&L.wq_v);
            a.n[0] = dim; a.n[1] = kv_dim; a.n[2] = kv_dim;
            a.out[0] = s->q; a.out[1] = s->key_cache; a.out[2] = s->value_cache;
            a.loff = loff; a.pPos = pPos;
            a.rope_tab = rope_tab; a.head_size = head_size;
            launch_gemv<GEMV_QKV>(a, (dim + 2 * kv_dim) / 4, pdl && l > 0);
        }
        launch_attention(s->xb, s->q, s->key_cache + loff, s->value_cache + loff, nullptr, p->n_heads, head_size,
                         kv_mul, seq_len_bin, pPos, pdl);
        {
            GemvParams a = {};
            a.x = s->xb; a.K = dim;
            a.m[0] = qw_view(&L.wq_o); a.n[0] = dim; a.out[0] = s->x;
            a.accum = 1; a.loff = -1;
            launch_gemv<GEMV_PLAIN>(a, dim / 4, pdl);
        }
        {
            GemvParams a = {};
            a.x = s->x; a.norm_w = L.rms_ffn_weight; a.K = dim;
            a.m[0] = qw_view(&L.wq_gate); a.m[1] = qw_view(&L.wq_up);
            a.n[0] = hidden_dim; a.out[0] = s->hb; a.loff = -1;
            launch_gemv<GEMV_FFN>(a, hidden_dim / 2, pdl);
        }
        {
            GemvParams a = {};
            a.x = s->hb; a.K = hidden_dim;
            a.m[0] = qw_view(&L.wq_down); a.n[0] = dim; a.out[0] = s->x;
            a.accum = 1; a.loff = -1;
            launch_gemv<GEMV_PLAIN>(a, dim / 4, pdl);
        }
    }
    launch_classifier(s->logits, s->x, w->rms_final_weight, w->wcls, dim, p->vocab_size, dim, 1.0f, pdl);
}

void lq4_run_llama_network(int* pPos, Config* p, RunState* s, TransformerWeights* w, int seq_len_bin) {
    ensure_init();
    if (g.opt_fused) {
        // the table must exist before a capture starts; run_transformer guarantees that, a direct caller
        // gets it built here (a synchronising call, so not legal inside capture on first use)
        const float2* tab = rope_table(p->rope_theta, p->dim / p->n_heads, p->seq_len);
        run_network_fused(pPos, p, s, w, seq_len_bin, tab);
    } else {
        run_network_unfused(pPos, p, s, w, seq_len_bin);
    }
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

void lq4_sample(Sampler* sampler, RunState* s, int gen_token, void* cuda_stream) {
    ensure_init();
    (void)random_u32(&sampler->rng_state);   // the reference burns one draw per step (sampler.h:45)
    if (sampler->temperature != 0.0f && gen_token) {
        // temperature / top-p sampling (sampler.h:51-81) is scope row f3: not built yet, fail loudly
        fprintf(stderr, "lq4: only greedy sampling (-t 0) is implemented\n");
        exit(EXIT_FAILURE);
    }
    argmax_kernel<<<1, 1024, 0, (cudaStream_t)cuda_stream>>>(s->logits, sampler->vocab_size,
                                                             &(s->shared_data->tokens[0]), &(s->shared_data->pos),
                                                             s->pos, nullptr, gen_token != 0);
}

static void run_forward_graphed(Config* p, RunState* s, TransformerWeights* w, int seq_len) {
    if (!g.opt_graphs) { lq4_run_llama_network(s->pos, p, s, w, seq_len); return; }
    // length bins 128,256,...,8192, last bin = max seq len (llama2_q4.cu:356-360)
    int graphIndex, seq_len_bin = 128;
    for (graphIndex = 0; graphIndex < MAX_GRAPHS - 1; seq_len_bin *= 2, graphIndex++)
        if (seq_len <= seq_len_bin) break;
    if ((seq_len > seq_len_bin) || (graphIndex == MAX_GRAPHS - 1)) seq_len_bin = p->seq_len;
    if (g.graph_owner != (const void*)s) { destroy_graphs(); g.graph_owner = s; }
    if (!g.graph_captured[graphIndex]) {
        if (g.opt_fused) (void)rope_table(p->rope_theta, p->dim / p->n_heads, p->seq_len);   // before capture
        cudaGraph_t graph = {};
        LQ4_CHECK(cudaStreamBeginCapture(g.stream, cudaStreamCaptureModeThreadLocal));
        lq4_run_llama_network(s->pos, p, s, w, seq_len_bin);
        LQ4_CHECK(cudaStreamEndCapture(g.stream, &graph));
        LQ4_CHECK(cudaGraphInstantiate(&g.graph_exec[graphIndex], graph, 0));
        cudaGraphDestroy(graph);
        g.graph_captured[graphIndex] = true;
    }
    LQ4_CHECK(cudaGraphLaunch(g.graph_exec[graphIndex], g.stream));
}

void lq4_run_transformer(int gen_token, Config* p, RunState* s, TransformerWeights* w, int copyLogits,
                         Sampler* pSampler) {
    ensure_init();
    const int seq_len = s->shared_data->pos + 1;       // llama2_q4.cu:354
    run_forward_graphed(p, s, w, seq_len);
    if (copyLogits) {                                  // llama2_q4.cu:377-382 (perplexity mode)
        float* pOutput = s->logits_array + (size_t)p->vocab_size * s->shared_data->pos;
        convert_fp16_to_fp32_kernel<<<divUp(p->vocab_size, 128), 128, 0, g.stream>>>(pOutput, s->logits, p->vocab_size);
    }
    lq4_sample(pSampler, s, gen_token, g.stream);
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
    return 0;
}

void lq4_free_transformer(Transformer* t) {
    destroy_graphs();
    RunState* s = &t->state;
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
    LQ4_CHECK(cudaStreamSynchronize(g.stream));
    lq4_run_transformer(gen_token, &t->config, &t->state, &t->weights, 0, sampler);
    LQ4_CHECK(cudaStreamSynchronize(g.stream));
    const int pos = t->state.shared_data->pos;
    if (logits_out)
        LQ4_CHECK(cudaMemcpy(logits_out, t->state.logits, sizeof(half) * t->config.vocab_size, cudaMemcpyDeviceToHost));
    if (next_token_out) *next_token_out = t->state.shared_data->tokens[pos];
    return pos;
}

// Enqueue one forward + sample without touching the host copy of the position: the caller states the
// sequence length (pos+1) that selects the graph bin.  Used by the pipelined loop and by bench.py.
void lq4_enqueue_step(Transformer* t, Sampler* sampler, int seq_len, int gen_token) {
    ensure_init();
    run_forward_graphed(&t->config, &t->state, &t->weights, seq_len);
    lq4_sample(sampler, &t->state, gen_token, g.stream);
}

int lq4_generate_tokens(Transformer* t, Sampler* sampler, const int* prompt_tokens, int n_prompt, int steps,
                        int* out_tokens, double* seconds, int pipelined) {
    ensure_init();
    if (n_prompt < 1) { fprintf(stderr, "something is wrong, expected at least 1 prompt token\n"); exit(EXIT_FAILURE); }
    if (steps <= 0 || steps > t->config.seq_len) steps = t->config.seq_len;   // llama2_q4.cu:690
    Config* p = &t->config;
    RunState* s = &t->state;
    const long start = time_in_ms();
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
        // The position and the sampled token live on the device (argmax_kernel writes both), so step
        // pos+1 can be enqueued before step pos has finished; the host only trails behind to read
        // tokens.  Events mark the end of each step.
        cudaEvent_t ev[2];
        LQ4_CHECK(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
        LQ4_CHECK(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
        int launched = 0;
        bool stop = false;
        while (pos < steps && !stop) {
            while (launched < steps && launched <= pos + 1) {
                // bin selection needs the sequence length: known on the host without reading it back
                run_forward_graphed(p, s, &t->weights, launched + 1);
                lq4_sample(sampler, s, launched >= n_prompt - 1, g.stream);
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
    const long end = time_in_ms();
    if (seconds) *seconds = (end - start) / 1000.0;
    return pos;
}

// ---------------------------------------------------------------------------------- synthetic files
size_t lq4_write_synth_model(const char* path, const Config* cfg, unsigned long long seed) {
    synth::Cfg c;
    memcpy(&c, cfg, sizeof c);
    return synth::write_model(path, c, seed);
}
size_t lq4_write_synth_tokenizer(const char* path, int vocab_size) { return synth::write_tokenizer(path, vocab_size); }

}  // extern "C"

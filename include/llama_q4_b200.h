/*
 * llama_q4_b200.h -- C ABI of the B200-native Llama-2 AWQ-INT4 decode engine.
 *
 * This is the drop-in boundary for the decode hot path of ankan-ban/llama_cu_awq.  The reference has
 * no FFI layer: its "operator API" is the set of host wrappers in llama2_q4.cu:207-395 that enqueue
 * kernels on one global stream.  Each entry point below replaces one of them (file:line cited), with
 * the same argument meaning, ownership (caller owns every device buffer; calls are asynchronous on the
 * engine stream and legal inside stream capture unless noted) and error behaviour (unsupported shapes
 * print "\nUnsupported matmul size. Exiting\n" and exit(EXIT_FAILURE), llama2_q4.cu:215,225,236,251).
 * Differences: `QWeight&` becomes `const QWeight*`, `bool` becomes `int`, default arguments are
 * explicit.  There is NO CPU fallback: every call needs a CUDA device of compute capability 10.x.
 *
 * Struct layouts (Config, QWeight, PerLayerWeight, TransformerWeights, SharedData, RunState,
 * Transformer) are byte-identical to the reference's common.h:9-78.
 */
#ifndef LLAMA_Q4_B200_H
#define LLAMA_Q4_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef LQ4_TYPES_H
typedef half lq4_half;
#else
typedef uint16_t lq4_half; /* IEEE binary16 bit pattern */

#define LQ4_MAX_SEQ_LEN (128 * 1024) /* reference common.h:7 */

typedef struct {
    int dim, hidden_dim, n_layers, n_heads, n_kv_heads, vocab_size, seq_len;
    float rope_theta;
} Config; /* common.h:9-18; also the first 32 bytes of a packed .bin */

typedef struct QWeight {
    uint32_t* weight; /* [N][ceil(K/32)*4] */
    uint32_t* zeros;  /* [N][ceil(ceil(K/128)/8)] */
    lq4_half* scales; /* [N][ceil(K/128)] */
} QWeight; /* common.h:20-24 */

typedef struct PerLayerWeight {
    lq4_half* rms_att_weight;
    lq4_half* rms_ffn_weight;
    QWeight wq_q, wq_k, wq_v, wq_o, wq_gate, wq_up, wq_down;
} PerLayerWeight; /* common.h:26-36 */

typedef struct {
    lq4_half* token_embedding_table;
    lq4_half* wcls;
    lq4_half* rms_final_weight;
    PerLayerWeight* layers;
    int num_layers;
} TransformerWeights; /* common.h:38-48 */

typedef struct SharedData {
    volatile int pos;
    int tokens[LQ4_MAX_SEQ_LEN];
} SharedData; /* common.h:51-54 (pinned host memory) */

typedef struct {
    lq4_half *x, *xb, *hb, *q, *att, *logits, *key_cache, *value_cache;
    int* pos;
    SharedData* shared_data;
    float* logits_array;
} RunState; /* common.h:56-72 */

typedef struct {
    Config config;
    TransformerWeights weights;
    RunState state;
} Transformer; /* common.h:74-78 */
#endif /* LQ4_TYPES_H */

/* reference sampler.h:3-13 (same fields, same order); define LQ4_HAVE_SAMPLER when sampler.h is already included */
#ifndef LQ4_HAVE_SAMPLER
typedef struct {
    int vocab_size;
    int* indices;
    void* tempStorage_scan;
    void* tempStorage_sort;
    size_t temp_storage_bytes_scan;
    size_t temp_storage_bytes_sort;
    float temperature;
    float topp;
    unsigned long long rng_state;
} Sampler;
#endif /* LQ4_HAVE_SAMPLER */

#ifdef __cplusplus
extern "C" {
#endif

/* ---- engine stream: replaces the global `cudaStream_t stream` (llama2_q4.cu:207, created :700) ---- */
int lq4_init(int device);                 /* select device, create the stream; 0 on success */
void* lq4_get_stream(void);               /* the cudaStream_t everything is enqueued on */
void lq4_set_stream(void* cuda_stream);   /* adopt a caller-owned stream (e.g. torch's) */
int lq4_stream_synchronize(void);         /* cudaStreamSynchronize + cudaGetLastError; 0 on success */
int lq4_stream_query(void);               /* non-blocking: 0 = everything enqueued has finished, 1 = still running, 2 = a CUDA error (see lq4_last_error) */
/* last error text.  A device-side wait that gives up (5 s: a protocol bug, or a tensor-parallel peer that died) leaves a record
 * in pinned host memory before it traps; the text then names what was waited for, the CTA, the SM and the rank. */
const char* lq4_last_error(void);
int lq4_sm_count(void);
void lq4_set_option(const char* name, int value); /* "pdl", "fused", "graphs", "opk" (0/1) ...: the run-time switches of INTEGRATION.md */

/* development aid: per-op timestamps (ns) of the last fused step after lq4_set_option("trace", 1) */
int lq4_debug_trace(unsigned long long* out, int* kinds, int max);

/* ---- operator API ---- */
/* rmsnorm, llama2_q4.cu:209-212 -> rmsnorm_kernel gpu_kernels.h:72-105 */
void lq4_rmsnorm(lq4_half* o, lq4_half* x, lq4_half* weight, int size);
/* matmul(half*), llama2_q4.cu:214-222 -> mat_vec_kernel gpu_kernels.h:109-139 */
void lq4_matmul_fp16(lq4_half* xout, lq4_half* x, lq4_half* w, int n, int d, int batch, int x_stride,
                     int w_stride, int op_stride, int w_row_stride, float alpha);
/* matmul(QWeight&), llama2_q4.cu:224-233 -> mat_vec_kernel_int4 gpu_kernels.h:171-240 */
void lq4_matmul_q4(lq4_half* xout, lq4_half* x, const QWeight* w, int inpSize, int opSize, int accum,
                   int loff, int* pPos);
/* qkv_matvec, llama2_q4.cu:235-248 -> qkv_matvec_kernel gpu_kernels.h:242-254 */
void lq4_qkv_matvec(lq4_half* q, lq4_half* key_cache, lq4_half* value_cache, lq4_half* x, const QWeight* qw,
                    const QWeight* kw, const QWeight* vw, int inpSize, int opSize, int loff, int* pPos);
/* ffn_matvec_silu, llama2_q4.cu:250-261 -> ffn_matvec_silu_kernel gpu_kernels.h:256-275 */
void lq4_ffn_matvec_silu(lq4_half* xout, lq4_half* x, const QWeight* gate_w, const QWeight* up_w,
                         int inpSize, int opSize);
/* RoPERotation, llama2_q4.cu:263-265 -> RoPERotation_kernel gpu_kernels.h:332-355 */
void lq4_rope_rotation(lq4_half* q, lq4_half* k, int num_heads, int num_kv_heads, int head_size, int* pPos,
                       int loff, float rope_theta);
/* MultiHeadAttention, llama2_q4.cu:267-284 -> mat_vec_kernel_simple :142-168, softmax_kernel :357-401,
 * vec_mat_kernel :279-329 (one fused kernel here; `att` still receives the probabilities) */
void lq4_multi_head_attention(lq4_half* output, lq4_half* q, lq4_half* key_cache, lq4_half* value_cache,
                              lq4_half* att, int num_heads, int head_size, int kv_mul, int max_seq_len,
                              int* pPos);
/* run_llama_network, llama2_q4.cu:286-340: one decode step at device position *pPos.
 * With the default fused step (option "fused" = 1) the activations travel in a private buffer: after the call ONLY
 * RunState::logits, the KV-cache rows of this position, RunState::pos / SharedData::pos and SharedData::tokens hold what the
 * reference leaves there; x, xb, hb, q and att are NOT written (a host that reads them must set option "fused" to 0, which
 * runs the reference's op sequence through the per-op wrappers above and fills every buffer).  Under tensor parallelism
 * (lq4_tp_config) this call, lq4_sample, copyLogits and non-greedy sampling exit with an error: only the fused greedy step
 * (lq4_run_transformer with temperature 0, lq4_enqueue_step, lq4_generate_tokens) is complete on every rank. */
void lq4_run_llama_network(int* pPos, Config* p, RunState* s, TransformerWeights* w, int seq_len_bin);
/* run_transformer, llama2_q4.cu:346-395: graph-cached forward + sample.  NOT capture-safe (it captures). */
void lq4_run_transformer(int gen_token, Config* p, RunState* s, TransformerWeights* w, int copyLogits,
                         Sampler* pSampler);
/* build_sampler / sample, sampler.h:15-23,43-81: greedy argmax, or temperature + top-p with the reference's cub pipeline */
void lq4_build_sampler(Sampler* sampler, int vocab_size, float temperature, float topp,
                       unsigned long long rng_seed);
void lq4_destroy_sampler(Sampler* sampler);
void lq4_sample(Sampler* sampler, RunState* s, int gen_token, void* cuda_stream);

/* ---- loader: build_transformer / free_transformer, llama2_q4.cu:408-432 (same .bin, B1 layout) ---- */
int lq4_build_transformer(Transformer* t, const char* checkpoint_path, int perplexity);
void lq4_free_transformer(Transformer* t);

/* ---- step driver on HOST buffers: the loop body of generate(), llama2_q4.cu:454-489, minus the
 * tokenizer.  prompt_tokens/out_tokens are host arrays; out_tokens[i] receives the token at sequence
 * position i (prompt echoed), returns the number of positions filled; *seconds = wall time of the loop.
 * pipelined != 0 launches step t+1 before waiting for step t (position and token live on the device). */
int lq4_generate_tokens(Transformer* t, Sampler* sampler, const int* prompt_tokens, int n_prompt, int steps,
                        int* out_tokens, double* seconds, int pipelined);
/* enqueue one forward+sample for sequence length seq_len (= pos+1) without any host<->device sync */
void lq4_enqueue_step(Transformer* t, Sampler* sampler, int seq_len, int gen_token);
/* teacher-forced single step for parity tests: resets nothing; returns new pos; logits_out host or NULL */
int lq4_step(Transformer* t, Sampler* sampler, int gen_token, lq4_half* logits_out, int* next_token_out);
void lq4_reset(Transformer* t, const int* tokens, int n);   /* generate() init, llama2_q4.cu:461-463 */

/* synchronises the engine stream, then copies device memory (e.g. RunState::logits_array, perplexity mode) to the host */
int lq4_memcpy_to_host(void* dst, const void* src_device, size_t bytes);

/* ---- batched prefill on the tensor cores (new: the reference feeds prompt tokens one by one through decode,
 * llama2_q4.cu:465-470).  Every projection of batch x seq token rows is one dense INT4 -> fp16 GEMM (tcgen05.mma, accumulators
 * in tensor memory, X tiles by TMA, W dequantised into shared memory on the fly).  Results agree with the sequential decode
 * path to fp16 tolerance, not bit for bit (tensor-core summation order). ---- */
/* y[M][N] = x[M][K] . dequant(w)[N][K]^T (+ y when accum); device pointers, row-major fp16; 1 = unsupported shape (K % 64, N % 128) */
int lq4_gemm_q4(lq4_half* y, const lq4_half* x, const QWeight* w, int M, int K, int N, int accum);
/* tokens: host [batch][seq]; logits_last: host [batch][vocab] or NULL; kv_seq >= 0 also fills the KV cache with that sequence and
 * leaves the positions at seq - 1 so that decode continues; ms_total / ms_gemm: device milliseconds (may be NULL); 1 = unsupported */
int lq4_prefill(Transformer* t, const int* tokens, int batch, int seq, int kv_seq, lq4_half* logits_last, float* ms_total,
                float* ms_gemm);

/* ---- tensor parallel, one process per GPU (new: the reference is single-GPU).  Column slices of every matrix per rank;
 * activations are broadcast by peer stores over NVLink (no collective call, no cross-GPU barrier); ids are bit-identical
 * to one GPU.  Order: lq4_tp_config -> lq4_build_transformer -> exchange lq4_tp_export handles -> lq4_tp_import each. ---- */
int lq4_tp_config(int rank, int world);
int lq4_tp_export(Transformer* t, void* handle64);
int lq4_tp_import(Transformer* t, int peer, const void* handle64);

/* ---- seeded random-init files in the reference formats (no network for real checkpoints) ---- */
size_t lq4_write_synth_model(const char* path, const Config* cfg, unsigned long long seed);
size_t lq4_write_synth_tokenizer(const char* path, int vocab_size);

#ifdef __cplusplus
}
#endif
#endif /* LLAMA_Q4_B200_H */

/*
 * oracle/cpu_ref.h -- CPU restatement of the reference decode hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Every function restates one kernel of ankan-ban/llama_cu_awq (gpu_kernels.h) in plain C,
 * keeping the exact fp32 summation DAG (per-lane FMA chains, cub shfl-down warp tree,
 * thread-0 sequential sum of warp aggregates) so that results are bit-identical to the CUDA
 * reference wherever no transcendental is involved.  expf / sinf / cosf / powf come from the
 * host libm here and may differ from CUDA's libdevice by an fp32 ulp: ops that use them
 * (SiLU, softmax, RoPE) are pinned bit-exactly against the reference CUDA build
 * (oracle/_ref) on the GPU box instead; against this file they are checked to 1 fp16 ulp.
 *
 * All fp16 data is carried as uint16_t bit patterns.
 */
#ifndef ORACLE_CPU_REF_H
#define ORACLE_CPU_REF_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int dim, hidden_dim, n_layers, n_heads, n_kv_heads, vocab_size, seq_len;
    float rope_theta;
} OracleConfig;                                 /* == reference Config, common.h:9-18 */

typedef struct {
    const uint32_t* weight;                     /* [N][pwh]  */
    const uint32_t* zeros;                      /* [N][zh]   */
    const uint16_t* scales;                     /* [N][G] fp16 */
} OracleQWeight;                                /* == reference QWeight, common.h:20-24 */

/* fp16 <-> fp32 (IEEE, round-to-nearest-even) */
float    oracle_h2f(uint16_t h);
uint16_t oracle_f2h(float f);

/* sizes: llama2_q4.cu:82-98 */
int oracle_pwh(int K);                          /* ceil(K/32)*4 words per column */
int oracle_groups(int K);                       /* ceil(K/128) */
int oracle_zh(int K);                           /* ceil(groups/8) */

/* gpu_kernels.h:171-210 + 213-233: one column; returns fp32 warp sum */
float oracle_dot_int4(int n, const uint16_t* x, const OracleQWeight* w, int K);
/* mat_vec_kernel_int4 (gpu_kernels.h:235-240): out[n] = half(sum (+ float(out[n]) if accum)) */
void oracle_matvec_int4(uint16_t* out, const uint16_t* x, const OracleQWeight* w, int K, int N, int accum);
/* ffn_matvec_silu_kernel (gpu_kernels.h:256-275) */
void oracle_ffn_matvec_silu(uint16_t* out, const uint16_t* x, const OracleQWeight* gate,
                            const OracleQWeight* up, int K, int N);
/* mat_vec_kernel (gpu_kernels.h:109-139): fp16 classifier GEMV, w row-major [d][n] */
void oracle_matvec_fp16(uint16_t* out, const uint16_t* x, const uint16_t* w, int n, int d, float alpha);
/* rmsnorm_kernel (gpu_kernels.h:72-105) */
void oracle_rmsnorm(uint16_t* o, const uint16_t* x, const uint16_t* weight, int size);
/* RoPERotation_kernel (gpu_kernels.h:332-355): q [n_heads*hs] and k row [n_kv_heads*hs] in place */
void oracle_rope(uint16_t* q, uint16_t* k, int n_heads, int n_kv_heads, int head_size, int pos, float theta);
/* mat_vec_kernel_simple + softmax_kernel + vec_mat_kernel (gpu_kernels.h:142-168,357-401,279-329)
 * K/V point at the layer's cache [seq][kv_dim]; att is scratch of n_heads*(pos+1) halfs. */
void oracle_attention(uint16_t* out, const uint16_t* q, const uint16_t* kcache, const uint16_t* vcache,
                      uint16_t* att, int n_heads, int head_size, int kv_mul, int pos);
/* pieces of the above, exposed so tests can pin each stage */
void oracle_qk_scores(uint16_t* att, const uint16_t* q, const uint16_t* kcache, int n_heads,
                      int head_size, int kv_mul, int pos);
void oracle_softmax(uint16_t* att, int n_heads, int pos);
void oracle_att_v(uint16_t* out, const uint16_t* att, const uint16_t* vcache, int n_heads,
                  int head_size, int kv_mul, int pos);
/* argmax_kernel (gpu_kernels.h:448-493); ties: lowest index (the reference's choice is a race) */
int oracle_argmax(const uint16_t* logits, int size);
/* number of indices holding the maximum (to classify reference-undefined ties) */
int oracle_argmax_ties(const uint16_t* logits, int size);

/* ---- whole model (llama2_q4.cu:286-340) on a .bin image held in host memory ---- */
typedef struct OracleModel OracleModel;
OracleModel* oracle_model_open(const char* bin_path);        /* mmap + parse, B1 layout */
void oracle_model_close(OracleModel* m);
const OracleConfig* oracle_model_config(const OracleModel* m);
/* one decode step at position pos for `token`; fills logits (vocab halfs); KV cache is internal.
 * max_layers < 0 => all layers (bench sampling may bound it). */
void oracle_model_forward(OracleModel* m, int token, int pos, uint16_t* logits, int max_layers);
/* intermediate taps after the last forward (x after all layers, pre final norm) */
const uint16_t* oracle_model_x(const OracleModel* m);
void oracle_set_threads(int n);                              /* OpenMP threads over output columns */
int oracle_get_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif

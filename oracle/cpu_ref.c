/*
 * oracle/cpu_ref.c -- CPU restatement of the reference decode hot path (see cpu_ref.h).
 *
 * TEST INFRASTRUCTURE ONLY -- never linked into, or called by, the product library.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -mfma -mf16c -shared -fPIC cpu_ref.c -lm
 *   -ffp-contract=off is REQUIRED: every fused multiply-add below is an explicit fmaf()
 *   placed where the reference SASS has an FFMA; the compiler must
 *   not invent others.
 *
 * Reduction helpers follow CUB 2.8.2 as shipped with CUDA 12.9 (the version the reference
 * compiles against here):
 *   warp:  cub/warp/specializations/warp_reduce_shfl.cuh:225-243,550-555  (shfl.down 1,2,4,8,16)
 *   block: cub/block/specializations/block_reduce_warp_reductions.cuh:140-198 (thread 0 adds
 *          warp aggregates 1..31 in order)
 */
#define _GNU_SOURCE
#include "cpu_ref.h"

#include <fcntl.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ fp16 */
float oracle_h2f(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1fu;
    uint32_t man = h & 0x3ffu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else { /* subnormal: normalise */
            int e = -1;
            do { man <<= 1; e++; } while (!(man & 0x400u));
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ffu) << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | (man << 13);
    } else {
        bits = sign | ((exp + 112u) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

uint16_t oracle_f2h(float f) { /* cvt.rn.f16.f32 */
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t ax = x & 0x7fffffffu;
    if (ax >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | (ax > 0x7f800000u ? 0x200u | ((ax >> 13) & 0x3ffu) : 0));
    if (ax >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u); /* >= 65520 rounds to inf */
    if (ax < 0x33000001u) return (uint16_t)sign;              /* <= 2^-25 rounds to 0 */
    int e = (int)(ax >> 23) - 127;
    uint32_t man = (ax & 0x7fffffu) | 0x800000u;
    int shift;
    uint32_t hexp;
    if (e < -14) { shift = 13 + (-14 - e); hexp = 0; }
    else { shift = 13; hexp = (uint32_t)(e + 15); }
    uint32_t q = man >> shift;
    uint32_t rem = man & ((1u << shift) - 1u);
    uint32_t half = 1u << (shift - 1);
    if (rem > half || (rem == half && (q & 1u))) q++;
    uint32_t out = (hexp == 0) ? q : ((hexp << 10) + (q - 0x400u)); /* carry propagates into exp */
    return (uint16_t)(sign | out);
}

/* ------------------------------------------------------------------ sizes */
static int div_up(int a, int b) { return (a - 1) / b + 1; }            /* common.h:80-82 */
int oracle_pwh(int K) { return div_up(K, 32) * 4; }                    /* llama2_q4.cu:82-88 */
int oracle_groups(int K) { return div_up(K, 128); }                    /* llama2_q4.cu:92 */
int oracle_zh(int K) { return div_up(oracle_groups(K), 8); }           /* llama2_q4.cu:93 */

/* ------------------------------------------------------------------ reductions */
static float warp_sum(const float* in) { /* lane 0 of cub::WarpReduce<float>::Sum */
    float v[32], t[32];
    memcpy(v, in, sizeof v);
    for (int o = 1; o < 32; o <<= 1) {
        for (int l = 0; l < 32; l++) t[l] = (l + o < 32) ? v[l] + v[l + o] : v[l];
        memcpy(v, t, sizeof v);
    }
    return v[0];
}

static float block_sum_1024(const float* in) { /* thread 0 of cub::BlockReduce<float,1024>::Sum */
    float agg = warp_sum(in);
    for (int w = 1; w < 32; w++) agg = agg + warp_sum(in + 32 * w);
    return agg;
}

/* ------------------------------------------------------------------ INT4 GEMV */
float oracle_dot_int4(int n, const uint16_t* x, const OracleQWeight* w, int K) {
    /* get_mat_vec_int4, gpu_kernels.h:171-210 */
    const int pwh = oracle_pwh(K), G = oracle_groups(K), zh = oracle_zh(K);
    float acc[32];
    for (int L = 0; L < 32; L++) {
        float sum = 0.0f;
        for (int ygq = 0; ygq * 128 + L * 4 < pwh; ygq++) {                  /* :176 */
            uint32_t packed_q_z = w->zeros[(size_t)n * zh + ygq];            /* :177 */
            const uint32_t* wp = &w->weight[(size_t)n * pwh + ygq * 128 + L * 4]; /* :181 */
            int group_y = ygq * 8 + (L / 4);                                 /* :183 */
            float q_z = (float)((packed_q_z >> (4 * (L / 4))) & 0xF);        /* :184 */
            float scale = oracle_h2f(w->scales[(size_t)n * G + group_y]);    /* :185 */
            int y_base = ygq * 1024 + L * 32;                                /* :186 */
            for (int qi = 0; qi < 4; qi++) {                                 /* :188 */
                int ys = y_base + qi * 8;
                if (ys < K) {                                                /* :190 */
                    uint32_t packed = wp[qi];
                    for (int i = 0; i < 8; i++) {                            /* :195-200 */
                        float q_wt = (float)(packed & 0xF);
                        float wt = (q_wt - q_z) * scale;
                        sum = fmaf(wt, oracle_h2f(x[ys + i]), sum);          /* FFMA */
                        packed >>= 4;
                    }
                }
            }
        }
        acc[L] = sum;
    }
    return warp_sum(acc);                                                    /* :205-207 */
}

void oracle_matvec_int4(uint16_t* out, const uint16_t* x, const OracleQWeight* w, int K, int N, int accum) {
    /* mat_vec_int4 epilogue, gpu_kernels.h:224-232 */
#pragma omp parallel for schedule(static)
    for (int n = 0; n < N; n++) {
        float sum = oracle_dot_int4(n, x, w, K);
        if (accum) sum = sum + oracle_h2f(out[n]);
        out[n] = oracle_f2h(sum);
    }
}

void oracle_ffn_matvec_silu(uint16_t* out, const uint16_t* x, const OracleQWeight* gate,
                            const OracleQWeight* up, int K, int N) {
    /* gpu_kernels.h:265-273 */
#pragma omp parallel for schedule(static)
    for (int n = 0; n < N; n++) {
        float g = oracle_dot_int4(n, x, gate, K);
        float u = oracle_dot_int4(n, x, up, K);
        float val = g;
        val = val * (1.0f / (1.0f + expf(-val)));   /* libm expf: see header note */
        val = val * u;
        out[n] = oracle_f2h(val);
    }
}

/* ------------------------------------------------------------------ fp16 GEMV (classifier) */
void oracle_matvec_fp16(uint16_t* out, const uint16_t* x, const uint16_t* w, int n, int d, float alpha) {
    /* mat_vec_kernel, gpu_kernels.h:109-139; host wrapper llama2_q4.cu:214-222 */
    int serialElements = div_up(n, 32);
    int numSerialLoads = div_up(serialElements, 8);
#pragma omp parallel for schedule(static)
    for (int index = 0; index < d; index++) {
        float acc[32];
        for (int L = 0; L < 32; L++) {
            float sum = 0.0f;
            for (int i = 0; i < numSerialLoads; i++) {
                int j = (i * 32 + L) * 8;
                if (j < n)
                    for (int el = 0; el < 8; el++)
                        sum = fmaf(oracle_h2f(w[(size_t)index * n + j + el]), oracle_h2f(x[j + el]), sum);
            }
            acc[L] = sum;
        }
        float sum = warp_sum(acc);
        sum = sum * alpha;
        out[index] = oracle_f2h(sum);
    }
}

/* ------------------------------------------------------------------ RMSNorm */
void oracle_rmsnorm(uint16_t* o, const uint16_t* x, const uint16_t* weight, int size) {
    /* rmsnorm_kernel, gpu_kernels.h:72-105 */
    int ept = div_up(size, 1024);
    float part[1024];
    for (int t = 0; t < 1024; t++) {
        float ss = 0.0f;
        for (int i = 0; i < ept; i++) {
            int index = t + i * 1024;
            if (index < size) {
                float val = oracle_h2f(x[index]);
                ss = fmaf(val, val, ss);                 /* ss += val*val contracts to FFMA */
            }
        }
        part[t] = ss;
    }
    float ss = block_sum_1024(part);
    ss = ss / (float)size;
    ss = ss + 1e-5f;
    ss = 1.0f / sqrtf(ss);
    /* output may alias input (final norm, llama2_q4.cu:336): element-wise, so safe */
    for (int index = 0; index < size; index++) {
        float val = oracle_h2f(x[index]);
        val = val * (ss * oracle_h2f(weight[index]));
        o[index] = oracle_f2h(val);
    }
}

/* ------------------------------------------------------------------ RoPE */
static void rope_pair(uint16_t* v, int i, int half_hs, float fcr, float fci, int is_k) {
    float v0 = oracle_h2f(v[i]);
    float v1 = oracle_h2f(v[i + half_hs]);
    /* FMA contraction as nvcc 12.9 -O3 emits it for sm_100a (cuobjdump -sass of the reference build,
     * RoPERotation_kernel):  q: out0 = fma(q0,c,-(q1*s)), out1 = fma(q1,c,q0*s)
     *                        k: out0 = fma(k0,c,-(k1*s)), out1 = fma(k0,s,k1*c)   <- different product fused */
    float o0 = fmaf(v0, fcr, -(v1 * fci));
    float o1 = is_k ? fmaf(v0, fci, v1 * fcr) : fmaf(v1, fcr, v0 * fci);
    v[i] = oracle_f2h(o0);
    v[i + half_hs] = oracle_f2h(o1);
}

void oracle_rope(uint16_t* q, uint16_t* k, int n_heads, int n_kv_heads, int head_size, int pos, float theta) {
    /* RoPERotation_kernel, gpu_kernels.h:332-355 */
    for (int h = 0; h < n_heads; h++) {
        for (int i = 0; i < head_size / 2; i++) {
            int head_dim = (i * 2) % head_size;
            float freq = 1.0f / powf(theta, (float)head_dim / (float)head_size);
            float val = (float)pos * freq;
            float fcr = cosf(val), fci = sinf(val);
            rope_pair(q + h * head_size, i, head_size / 2, fcr, fci, 0);
            if (h < n_kv_heads) rope_pair(k + h * head_size, i, head_size / 2, fcr, fci, 1);
        }
    }
}

/* ------------------------------------------------------------------ attention */
void oracle_qk_scores(uint16_t* att, const uint16_t* q, const uint16_t* kcache, int n_heads,
                      int head_size, int kv_mul, int pos) {
    /* mat_vec_kernel_simple, gpu_kernels.h:142-168; launch llama2_q4.cu:270-273 */
    int dim = n_heads * head_size;
    int row_stride = dim / kv_mul;
    int nser = div_up(head_size, 32);
    float alpha = (float)(1.0 / sqrt((double)head_size));
    int size = pos + 1;
    for (int h = 0; h < n_heads; h++) {
        const uint16_t* input = q + h * head_size;
        const uint16_t* weight = kcache + (h / kv_mul) * head_size;
        for (int t = 0; t < size; t++) {
            float acc[32];
            for (int L = 0; L < 32; L++) {
                float sum = 0.0f;
                for (int i = 0; i < nser; i++) {
                    int j = i * 32 + L;
                    if (j < head_size)
                        sum = fmaf(oracle_h2f(weight[(size_t)t * row_stride + j]), oracle_h2f(input[j]), sum);
                }
                acc[L] = sum;
            }
            float sum = warp_sum(acc) * alpha;
            att[(size_t)h * size + t] = oracle_f2h(sum);
        }
    }
}

void oracle_softmax(uint16_t* att_h, int n_heads, int pos) {
    /* softmax_kernel, gpu_kernels.h:357-401 (pos+1 <= 8192) */
    int size = pos + 1;
    float* att = (float*)malloc(sizeof(float) * (size_t)size);
    for (int h = 0; h < n_heads; h++) {
        uint16_t* arr = att_h + (size_t)h * size;
        for (int t = 0; t < size; t++) att[t] = oracle_h2f(arr[t]);
        /* idle threads contribute 0 to the max (:374) => max(true max, 0) when size < 1024 */
        float max_val = (size < 1024) ? 0.0f : att[0];
        for (int t = 0; t < size; t++)
            if (att[t] > max_val) max_val = att[t];
        float part[1024];
        for (int t = 0; t < 1024; t++) {
            float sum = 0.0f;
            for (int i = t; i < size; i += 1024) {
                att[i] = expf(att[i] - max_val);       /* libm expf: see header note */
                sum = sum + att[i];
            }
            part[t] = sum;
        }
        float sum = block_sum_1024(part);
        for (int t = 0; t < size; t++) arr[t] = oracle_f2h(att[t] / sum);
    }
    free(att);
}

void oracle_att_v(uint16_t* out, const uint16_t* att, const uint16_t* vcache, int n_heads,
                  int head_size, int kv_mul, int pos) {
    /* vec_mat_kernel, gpu_kernels.h:279-329; launch llama2_q4.cu:282-283 */
    int dim = n_heads * head_size;
    int row_stride = dim / kv_mul;
    int K = pos + 1;
    for (int h = 0; h < n_heads; h++) {
        const uint16_t* input = att + (size_t)h * K;
        const uint16_t* weight = vcache + (h / kv_mul) * head_size;
        for (int i = 0; i < head_size; i++) {
            float acc[32];
            for (int tx = 0; tx < 32; tx++) {
                float sum = 0.0f;
                for (int e = 0; e * 32 < K; e++) {
                    int k = e * 32 + tx;
                    /* out-of-range rows are zero-filled and multiplied by 0.0f: fma(0,0,sum)=sum */
                    float wv = (k < K) ? oracle_h2f(weight[(size_t)k * row_stride + i]) : 0.0f;
                    float pv = (k < K) ? oracle_h2f(input[k]) : 0.0f;
                    sum = fmaf(wv, pv, sum);
                }
                acc[tx] = sum;
            }
            out[h * head_size + i] = oracle_f2h(warp_sum(acc));
        }
    }
}

void oracle_attention(uint16_t* out, const uint16_t* q, const uint16_t* kcache, const uint16_t* vcache,
                      uint16_t* att, int n_heads, int head_size, int kv_mul, int pos) {
    /* MultiHeadAttention, llama2_q4.cu:267-284 */
    oracle_qk_scores(att, q, kcache, n_heads, head_size, kv_mul, pos);
    oracle_softmax(att, n_heads, pos);
    oracle_att_v(out, att, vcache, n_heads, head_size, kv_mul, pos);
}

/* ------------------------------------------------------------------ argmax */
int oracle_argmax(const uint16_t* logits, int size) {
    /* argmax_kernel, gpu_kernels.h:448-493.  Value compare is in fp32; the winning index among
     * equal maxima is unspecified in the reference (racy write, :474-479): lowest index here. */
    float best = -INFINITY;
    int pos = 0;
    for (int i = 0; i < size; i++) {
        float v = oracle_h2f(logits[i]);
        if (v > best) { best = v; pos = i; }
    }
    return pos;
}

int oracle_argmax_ties(const uint16_t* logits, int size) {
    float best = oracle_h2f(logits[oracle_argmax(logits, size)]);
    int n = 0;
    for (int i = 0; i < size; i++) n += (oracle_h2f(logits[i]) == best);
    return n;
}

/* ------------------------------------------------------------------ whole model */
typedef struct {
    const uint16_t *rms_att, *rms_ffn;
    OracleQWeight q, k, v, o, gate, up, down;
} OracleLayer;

struct OracleModel {
    OracleConfig cfg;
    void* map;
    size_t map_size;
    const uint16_t *emb, *wcls, *rms_final;
    OracleLayer* layers;
    uint16_t *x, *xb, *hb, *q, *att, *kcache, *vcache;
};

static const uint8_t* take_q(const uint8_t* p, OracleQWeight* w, int K, int N) {
    /* uploadQWeight, llama2_q4.cu:162-170 */
    w->weight = (const uint32_t*)p; p += (size_t)oracle_pwh(K) * N * 4;
    w->zeros = (const uint32_t*)p;  p += (size_t)oracle_zh(K) * N * 4;
    w->scales = (const uint16_t*)p; p += (size_t)oracle_groups(K) * N * 2;
    return p;
}

OracleModel* oracle_model_open(const char* bin_path) {
    int fd = open(bin_path, O_RDONLY);
    if (fd < 0) return NULL;
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); return NULL; }
    void* map = mmap(NULL, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (map == MAP_FAILED) return NULL;
    OracleModel* m = (OracleModel*)calloc(1, sizeof *m);
    m->map = map;
    m->map_size = (size_t)st.st_size;
    memcpy(&m->cfg, map, sizeof(OracleConfig));                      /* llama2_q4.cu:414 */
    const OracleConfig* c = &m->cfg;
    int kv_dim = (c->dim * c->n_kv_heads) / c->n_heads;
    const uint8_t* p = (const uint8_t*)map + sizeof(OracleConfig);
    /* checkpoint_init_weights, llama2_q4.cu:180-197 */
    m->emb = (const uint16_t*)p;  p += (size_t)c->vocab_size * c->dim * 2;
    m->wcls = (const uint16_t*)p; p += (size_t)c->vocab_size * c->dim * 2;
    m->rms_final = (const uint16_t*)p; p += (size_t)c->dim * 2;
    m->layers = (OracleLayer*)calloc((size_t)c->n_layers, sizeof(OracleLayer));
    for (int l = 0; l < c->n_layers; l++) {
        OracleLayer* L = &m->layers[l];
        p = take_q(p, &L->q, c->dim, c->dim);
        p = take_q(p, &L->k, c->dim, kv_dim);
        p = take_q(p, &L->v, c->dim, kv_dim);
        p = take_q(p, &L->o, c->dim, c->dim);
        p = take_q(p, &L->up, c->dim, c->hidden_dim);                /* up BEFORE gate (:191-192) */
        p = take_q(p, &L->gate, c->dim, c->hidden_dim);
        p = take_q(p, &L->down, c->hidden_dim, c->dim);
        L->rms_att = (const uint16_t*)p; p += (size_t)c->dim * 2;
        L->rms_ffn = (const uint16_t*)p; p += (size_t)c->dim * 2;
    }
    if ((size_t)(p - (const uint8_t*)map) != m->map_size) {
        fprintf(stderr, "oracle: .bin size mismatch: parsed %zu, file %zu\n",
                (size_t)(p - (const uint8_t*)map), m->map_size);
        oracle_model_close(m);
        return NULL;
    }
    m->x = (uint16_t*)calloc((size_t)c->dim, 2);
    m->xb = (uint16_t*)calloc((size_t)c->dim, 2);
    m->q = (uint16_t*)calloc((size_t)c->dim, 2);
    m->hb = (uint16_t*)calloc((size_t)c->hidden_dim, 2);
    m->att = (uint16_t*)calloc((size_t)c->n_heads * c->seq_len, 2);
    m->kcache = (uint16_t*)calloc((size_t)c->n_layers * c->seq_len * kv_dim, 2);
    m->vcache = (uint16_t*)calloc((size_t)c->n_layers * c->seq_len * kv_dim, 2);
    return m;
}

void oracle_model_close(OracleModel* m) {
    if (!m) return;
    if (m->map) munmap(m->map, m->map_size);
    free(m->layers); free(m->x); free(m->xb); free(m->q); free(m->hb); free(m->att);
    free(m->kcache); free(m->vcache);
    free(m);
}

const OracleConfig* oracle_model_config(const OracleModel* m) { return &m->cfg; }
const uint16_t* oracle_model_x(const OracleModel* m) { return m->x; }

void oracle_model_forward(OracleModel* m, int token, int pos, uint16_t* logits, int max_layers) {
    /* run_llama_network, llama2_q4.cu:286-340 */
    const OracleConfig* c = &m->cfg;
    int dim = c->dim, hidden = c->hidden_dim;
    int head_size = dim / c->n_heads;
    int kv_dim = (dim * c->n_kv_heads) / c->n_heads;
    int kv_mul = c->n_heads / c->n_kv_heads;
    int nl = (max_layers < 0 || max_layers > c->n_layers) ? c->n_layers : max_layers;
    memcpy(m->x, m->emb + (size_t)token * dim, (size_t)dim * 2);      /* copy_embedding_kernel */
    for (int l = 0; l < nl; l++) {
        const OracleLayer* L = &m->layers[l];
        oracle_rmsnorm(m->xb, m->x, L->rms_att, dim);
        size_t loff = (size_t)l * c->seq_len * kv_dim;
        uint16_t* krow = m->kcache + loff + (size_t)pos * kv_dim;
        uint16_t* vrow = m->vcache + loff + (size_t)pos * kv_dim;
        oracle_matvec_int4(m->q, m->xb, &L->q, dim, dim, 0);
        oracle_matvec_int4(krow, m->xb, &L->k, dim, kv_dim, 0);
        oracle_matvec_int4(vrow, m->xb, &L->v, dim, kv_dim, 0);
        oracle_rope(m->q, krow, c->n_heads, c->n_kv_heads, head_size, pos, c->rope_theta);
        oracle_attention(m->xb, m->q, m->kcache + loff, m->vcache + loff, m->att, c->n_heads,
                         head_size, kv_mul, pos);
        oracle_matvec_int4(m->x, m->xb, &L->o, dim, dim, 1);
        oracle_rmsnorm(m->xb, m->x, L->rms_ffn, dim);
        oracle_ffn_matvec_silu(m->hb, m->xb, &L->gate, &L->up, dim, hidden);
        oracle_matvec_int4(m->x, m->hb, &L->down, hidden, dim, 1);
    }
    oracle_rmsnorm(m->x, m->x, m->rms_final, dim);
    if (logits) oracle_matvec_fp16(logits, m->x, m->wcls, dim, c->vocab_size, 1.0f);
}

void oracle_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n > 0 ? n : 1);
#else
    (void)n;
#endif
}

int oracle_get_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// lq4_types.h -- the struct contract of the drop-in boundary.
//
// Byte-compatible with the reference's common.h (ankan-ban/llama_cu_awq common.h:6-82): same field
// order and types, so host code written against the reference's Config / QWeight / RunState /
// Transformer links against this engine unchanged, and the first 32 bytes of a packed `.bin` are a
// Config.  Sizes are asserted below (values probed from the reference header).
#pragma once
#define LQ4_TYPES_H

#include <cuda_fp16.h>
#include <stddef.h>
#include <stdint.h>

constexpr int MAX_SEQ_LEN_SMEM_KERNEL = 8192;  // reference common.h:6
constexpr int MAX_SEQ_LEN = 128 * 1024;        // reference common.h:7

typedef struct {
    int dim;         // transformer dimension
    int hidden_dim;  // ffn hidden dimension
    int n_layers;
    int n_heads;     // query heads
    int n_kv_heads;  // key/value heads (<= n_heads)
    int vocab_size;
    int seq_len;     // max sequence length
    float rope_theta;
} Config;  // reference common.h:9-18

struct QWeight {          // reference common.h:20-24
    uint32_t* weight;     // [N][ceil(K/32)*4]  8 little-endian nibbles per word along K
    uint32_t* zeros;      // [N][ceil(G/8)]     nibble j of word y = zero point of group 8y+j
    half* scales;         // [N][G]             G = ceil(K/128)
};

struct PerLayerWeight {   // reference common.h:26-36
    half* rms_att_weight;
    half* rms_ffn_weight;
    QWeight wq_q;
    QWeight wq_k;
    QWeight wq_v;
    QWeight wq_o;
    QWeight wq_gate;
    QWeight wq_up;
    QWeight wq_down;
};

typedef struct {          // reference common.h:38-48
    half* token_embedding_table;  // (vocab_size, dim)
    half* wcls;                   // (vocab_size, dim)
    half* rms_final_weight;       // (dim,)
    PerLayerWeight* layers;
    int num_layers;
} TransformerWeights;

struct SharedData {       // reference common.h:51-54 (pinned host memory, GPU-written)
    volatile int pos;
    int tokens[MAX_SEQ_LEN];
};

typedef struct {          // reference common.h:56-72
    half* x;
    half* xb;
    half* hb;
    half* q;
    half* att;
    half* logits;
    half* key_cache;    // (layer, seq_len, kv_dim)
    half* value_cache;  // (layer, seq_len, kv_dim)
    int* pos;           // device copy of the current position
    SharedData* shared_data;
    float* logits_array;
} RunState;

typedef struct {          // reference common.h:74-78
    Config config;
    TransformerWeights weights;
    RunState state;
} Transformer;

static_assert(sizeof(Config) == 32 && offsetof(Config, rope_theta) == 28, "Config layout");
static_assert(sizeof(QWeight) == 24, "QWeight layout");
static_assert(sizeof(PerLayerWeight) == 184 && offsetof(PerLayerWeight, wq_q) == 16, "PerLayerWeight layout");
static_assert(sizeof(TransformerWeights) == 40, "TransformerWeights layout");
static_assert(sizeof(SharedData) == 524292 && offsetof(SharedData, tokens) == 4, "SharedData layout");
static_assert(sizeof(RunState) == 88 && offsetof(RunState, pos) == 64 && offsetof(RunState, shared_data) == 72,
              "RunState layout");
static_assert(sizeof(Transformer) == 160, "Transformer layout");

static inline int divUp(int a, int b) { return (a - 1) / b + 1; }  // reference common.h:80-82

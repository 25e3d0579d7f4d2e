// synth.h -- seeded random-init model / tokenizer writers in the reference's on-disk formats.
//
// There is no network for real AWQ checkpoints, so benchmarks and parity runs use a random-init
// `.bin` with exactly the layout the reference loader reads (checkpoint_init_weights,
// llama2_q4.cu:172-202; uploadQWeight :162-170) and that its packer writes (weight_packer.cpp:
// 256-291): Config(32 B) | token_embedding fp16 [vocab][dim] | wcls fp16 [vocab][dim] |
// rms_final fp16 [dim] | per layer { q k v o up gate down : qweight u32 [N][pwh], qzeros u32
// [N][zh], scales fp16 [N][G] } rms_att fp16 [dim] rms_ffn fp16 [dim].
//
// Distributions: qweight/qzeros uniform nibbles; scales = fp16 of
// s0*U(0.5,1.5) with s0 = 1/(6.52*sqrt(K)) so dequantised weights have std ~ 1/sqrt(K);
// norms fp16 U(0.9,1.1); embeddings ~N(0,1); classifier ~N(0,0.02) with row 2 (EOS) zeroed so
// greedy decoding never stops early (llama2_q4.cu:477).
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

namespace synth {

struct Cfg {  // == reference Config (common.h:9-18)
    int dim, hidden_dim, n_layers, n_heads, n_kv_heads, vocab_size, seq_len;
    float rope_theta;
};

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    inline uint64_t next() {  // splitmix64
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    inline float uniform() { return (float)(next() >> 40) * (1.0f / 16777216.0f); }  // [0,1)
    inline float normal() {  // Irwin-Hall(4) scaled to unit variance: plenty for synthetic init
        uint64_t r = next();
        float a = (float)(r & 0xffff) + (float)((r >> 16) & 0xffff) + (float)((r >> 32) & 0xffff) +
                  (float)((r >> 48) & 0xffff);
        return (a * (1.0f / 65536.0f) - 2.0f) * 1.7320508f;
    }
};

inline uint16_t f2h(float f) {  // round-to-nearest-even fp32 -> fp16
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u, ax = x & 0x7fffffffu;
    if (ax >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);
    if (ax < 0x33000001u) return (uint16_t)sign;
    int e = (int)(ax >> 23) - 127;
    uint32_t man = (ax & 0x7fffffu) | 0x800000u;
    int shift = (e < -14) ? 13 + (-14 - e) : 13;
    uint32_t hexp = (e < -14) ? 0u : (uint32_t)(e + 15);
    uint32_t q = man >> shift, rem = man & ((1u << shift) - 1u), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (q & 1u))) q++;
    return (uint16_t)(sign | (hexp == 0 ? q : ((hexp << 10) + (q - 0x400u))));
}

inline int div_up(int a, int b) { return (a - 1) / b + 1; }

struct Writer {
    FILE* f;
    uint64_t seed;
    uint64_t tensor_idx = 0;
    size_t bytes = 0;
    std::vector<uint8_t> buf;
    Writer(FILE* f_, uint64_t seed_) : f(f_), seed(seed_), buf(1 << 22) {}
    Rng rng_for_next() { return Rng(seed * 0x100000001B3ull + 0x5EED + 0x9E37ull * (tensor_idx++)); }
    bool put(const void* p, size_t n) {
        bytes += n;
        return fwrite(p, 1, n, f) == n;
    }
    template <class Gen>
    bool fill_u16(size_t count, Gen gen) {
        uint16_t* b = (uint16_t*)buf.data();
        size_t cap = buf.size() / 2;
        while (count) {
            size_t n = count < cap ? count : cap;
            for (size_t i = 0; i < n; i++) b[i] = gen();
            if (!put(b, n * 2)) return false;
            count -= n;
        }
        return true;
    }
    bool fill_u32_random(size_t count, Rng& r) {
        uint64_t* b = (uint64_t*)buf.data();
        size_t cap = buf.size() / 8;
        size_t pairs = count / 2;
        while (pairs) {
            size_t n = pairs < cap ? pairs : cap;
            for (size_t i = 0; i < n; i++) b[i] = r.next();
            if (!put(b, n * 8)) return false;
            pairs -= n;
        }
        if (count & 1) {
            uint32_t w = (uint32_t)r.next();
            return put(&w, 4);
        }
        return true;
    }
    bool fp16_normal(size_t rows, size_t cols, float std, long zero_row) {
        Rng r = rng_for_next();
        for (size_t row = 0; row < rows; row++) {
            bool z = ((long)row == zero_row);
            if (!fill_u16(cols, [&]() { float v = r.normal() * std; return z ? (uint16_t)0 : f2h(v); })) return false;
        }
        return true;
    }
    bool fp16_uniform(size_t count, float lo, float hi) {
        Rng r = rng_for_next();
        return fill_u16(count, [&]() { return f2h(lo + (hi - lo) * r.uniform()); });
    }
    bool qweight(int K, int N) {  // uploadQWeight order: weight, zeros, scales
        int pwh = div_up(K, 32) * 4, G = div_up(K, 128), zh = div_up(G, 8);
        Rng rw = rng_for_next();
        if (!fill_u32_random((size_t)pwh * N, rw)) return false;
        Rng rz = rng_for_next();
        if (!fill_u32_random((size_t)zh * N, rz)) return false;
        Rng rs = rng_for_next();
        float s0 = 1.0f / (6.52f * sqrtf((float)K));
        return fill_u16((size_t)G * N, [&]() { return f2h(s0 * (0.5f + rs.uniform())); });
    }
};

// returns bytes written, 0 on failure
inline size_t write_model(const char* path, const Cfg& c, uint64_t seed) {
    FILE* f = fopen(path, "wb");
    if (!f) return 0;
    Writer w(f, seed);
    int kv_dim = (c.dim * c.n_kv_heads) / c.n_heads;
    bool ok = w.put(&c, sizeof(Cfg));
    ok = ok && w.fp16_normal((size_t)c.vocab_size, (size_t)c.dim, 1.0f, -1);
    ok = ok && w.fp16_normal((size_t)c.vocab_size, (size_t)c.dim, 0.02f, 2);
    ok = ok && w.fp16_uniform((size_t)c.dim, 0.9f, 1.1f);
    for (int l = 0; ok && l < c.n_layers; l++) {
        ok = ok && w.qweight(c.dim, c.dim);          // q
        ok = ok && w.qweight(c.dim, kv_dim);         // k
        ok = ok && w.qweight(c.dim, kv_dim);         // v
        ok = ok && w.qweight(c.dim, c.dim);          // o
        ok = ok && w.qweight(c.dim, c.hidden_dim);   // up   (before gate: llama2_q4.cu:191-192)
        ok = ok && w.qweight(c.dim, c.hidden_dim);   // gate
        ok = ok && w.qweight(c.hidden_dim, c.dim);   // down
        ok = ok && w.fp16_uniform((size_t)c.dim, 0.9f, 1.1f);
        ok = ok && w.fp16_uniform((size_t)c.dim, 0.9f, 1.1f);
    }
    size_t bytes = w.bytes;
    if (fclose(f) != 0) ok = false;
    return ok ? bytes : 0;
}

inline size_t model_bytes(const Cfg& c) {
    auto q = [](size_t K, size_t N) {
        size_t pwh = (size_t)div_up((int)K, 32) * 4, G = (size_t)div_up((int)K, 128), zh = (size_t)div_up((int)G, 8);
        return N * (pwh * 4 + zh * 4 + G * 2);
    };
    size_t kv_dim = (size_t)(c.dim * c.n_kv_heads) / c.n_heads, d = c.dim, h = c.hidden_dim;
    size_t per_layer = q(d, d) * 2 + q(d, kv_dim) * 2 + q(d, h) * 2 + q(h, d) + d * 4;
    return sizeof(Cfg) + (size_t)c.vocab_size * d * 4 + d * 2 + per_layer * c.n_layers;
}

// Synthetic tokenizer in the reference format (tokenizer.h:49-57): int max_token_length, then
// vocab x { float score; int len; char[len] }.  ids 0..2 = <unk> <s> </s>; ids 3..258 = the raw
// byte (id-3), as in the shipped tokenizer.bin, so encode()'s byte fallback (+3, tokenizer.h:
// 178-183) and its " " lookup (:132-136) work; ids >= 259 print as "[id]" so stdout is an exact
// id transcript.  All scores 0 and no piece is a concatenation of two others => no BPE merges.
inline size_t write_tokenizer(const char* path, int vocab) {
    FILE* f = fopen(path, "wb");
    if (!f) return 0;
    size_t bytes = 0;
    int maxlen = 16;
    bytes += fwrite(&maxlen, 1, 4, f);
    for (int i = 0; i < vocab; i++) {
        char piece[32];
        int len;
        if (i == 0) len = snprintf(piece, sizeof piece, "<unk>");
        else if (i == 1) len = snprintf(piece, sizeof piece, "<s>");
        else if (i == 2) len = snprintf(piece, sizeof piece, "</s>");
        else if (i < 259) { piece[0] = (char)(i - 3); piece[1] = 0; len = 1; }
        else len = snprintf(piece, sizeof piece, "[%d]", i);
        float score = 0.0f;
        bytes += fwrite(&score, 1, 4, f);
        bytes += fwrite(&len, 1, 4, f);
        bytes += fwrite(piece, 1, (size_t)len, f);
    }
    fclose(f);
    return bytes;
}

}  // namespace synth

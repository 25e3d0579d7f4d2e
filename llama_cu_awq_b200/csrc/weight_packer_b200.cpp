// weight_packer_b200.cpp -- offline packer: AWQ per-tensor dumps -> the `.bin` this engine (and the reference) loads.
//
// Drop-in for the reference tool (ankan-ban/llama_cu_awq weight_packer.cpp:224-296): same command line
//   weight_packer_b200 <config.json from huggingface> <path_to_awq_bin_weights> <output_bin_filename> <OldAwqFormat: 0 or 1>
// same input file names (`model.layers.N.self_attn.q_proj.qweight.bin` ...), same output bytes (surface
// B1): Config | embed_tokens | lm_head | norm | per layer { q k v o up gate down : qweight [N][K/8], qzeros [N][ceil(G/8)],
// scales [N][G] } input_layernorm, post_attention_layernorm.
//
// Two input flavours (weight_packer.cpp:146-205):
//   old AWQ (1): qweight int32 [K][N/8], qzeros int32 [G][N/8] -- eight consecutive OUTPUT columns per word, nibble i of a
//                word holding column {0,2,4,6,1,3,5,7}[i] (:92) -- and scales fp16 [G][N]; all three are transposed so that
//                the reduction index runs fastest and nibble i of an output word is element 8*word + i (:86-128);
//   new AWQ (0): already [N][K/8] / [N][ceil(G/8)]; only the scales are padded to [N][8*ceil(G/8)] and the padding is dropped.
//
// Written as direct index arithmetic, one output word at a time (the reference goes through a K*N array of unpacked
// nibbles).  Where the reference reads past a column -- the high nibbles of the last zeros word when G % 8 != 0 (:117-121)
// -- it picks up the first groups of the NEXT column; that is reproduced, and the last column, where the reference reads
// past its allocation, gets zeros (the loader never looks at those nibbles).  Pure host code: there is nothing here for a GPU.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/llama_q4_b200.h"   // Config (== common.h:9-18)

namespace {

constexpr int kGroup = 128;   // weight_packer.cpp:7

[[noreturn]] void die(const std::string& msg) {
    printf("%s", msg.c_str());
    exit(1);
}

int json_int(const std::string& json, const char* key, bool required, int fallback) {
    const std::string pat = std::string("\"") + key + "\":";
    const size_t p = json.find(pat);
    if (p == std::string::npos) {
        if (required) die(std::string("error parsing config.json ") + key + " not found");
        return fallback;
    }
    return atoi(json.c_str() + p + pat.size());
}

Config read_config(const char* path) {   // weight_packer.cpp:22-73
    FILE* f = fopen(path, "rb");
    if (!f) die("unable to open config file\n");
    std::string json(1 << 20, '\0');
    const size_t n = fread(&json[0], 1, json.size() - 1, f);
    fclose(f);
    if (n == 0) die("unable to read config file\n");
    json.resize(n);
    Config c;
    c.dim = json_int(json, "hidden_size", true, 0);
    c.hidden_dim = json_int(json, "intermediate_size", true, 0);
    c.n_layers = json_int(json, "num_hidden_layers", true, 0);
    c.n_heads = json_int(json, "num_attention_heads", true, 0);
    c.n_kv_heads = json_int(json, "num_key_value_heads", false, c.n_heads);
    c.vocab_size = json_int(json, "vocab_size", true, 0);
    c.seq_len = json_int(json, "max_position_embeddings", true, 0);
    const size_t p = json.find("\"rope_theta\":");
    c.rope_theta = (p == std::string::npos) ? 10000.0f : (float)atof(json.c_str() + p + strlen("\"rope_theta\":"));
    printf("\nModel params:- \ndim: %d \nhidden_dim: %d\nn_heads: %d\nn_kv_heads: %d\nn_layers: %d\nseq_len: %d\nvocab_size: %d\nrope_theta: %g\n",
           c.dim, c.hidden_dim, c.n_heads, c.n_kv_heads, c.n_layers, c.seq_len, c.vocab_size, c.rope_theta);
    return c;
}

template <typename T>
std::vector<T> read_file(const std::string& name, size_t count) {
    std::vector<T> v(count);
    FILE* f = fopen(name.c_str(), "rb");
    if (!f) die("\nUnable to open " + name + "\n");
    if (fread(v.data(), sizeof(T), count, f) != count) die("error reading weights from " + name);
    fclose(f);
    return v;
}
template <typename T>
void write_all(FILE* f, const std::vector<T>& v, const std::string& what) {
    if (fwrite(v.data(), sizeof(T), v.size(), f) != v.size()) die("error writing output file from input " + what);
}
void copy_fp16(FILE* out, const std::string& name, size_t count) { write_all(out, read_file<uint16_t>(name, count), name); }

inline int div_up(int a, int b) { return (a - 1) / b + 1; }

// Element (row y, column x) of an old-AWQ packed matrix [height][width/8]: column x sits in nibble kNibbleOf[x % 8].
constexpr int kNibbleOf[8] = {0, 4, 1, 5, 2, 6, 3, 7};   // inverse of the reference's order_map {0,2,4,6,1,3,5,7}
inline uint32_t old_awq_at(const std::vector<uint32_t>& in, int width, int y, int x) {
    return (in[((size_t)y * width + x) / 8] >> (4 * kNibbleOf[x & 7])) & 0xFu;
}

// old AWQ [height][width/8] -> [width][ceil(height/8)], nibble i of word w of column x = element (8w + i, x).
// Elements past `height` continue into the next column (flat index x*height + y), zeros past the last one.
std::vector<uint32_t> transpose_pack(const std::vector<uint32_t>& in, int height, int width) {
    const int ph = div_up(height, 8);
    std::vector<uint32_t> out((size_t)width * ph);
    const long long total = (long long)width * height;
    for (int x = 0; x < width; x++)
        for (int w = 0; w < ph; w++) {
            uint32_t v = 0;
            for (int i = 0; i < 8; i++) {
                const long long flat = (long long)x * height + w * 8 + i;
                if (flat < total) v |= old_awq_at(in, width, (int)(flat % height), (int)(flat / height)) << (4 * i);
            }
            out[(size_t)x * ph + w] = v;
        }
    return out;
}

void pack_matrix(FILE* out, const std::string& base, const char* name, int height, int width, bool old_format) {   // :146-222
    const int G = div_up(height, kGroup), ph = div_up(height, 8), zh = div_up(G, 8);
    const std::string stem = base + "." + name;
    std::vector<uint32_t> qweight, qzeros;
    std::vector<uint16_t> scales((size_t)G * width);
    if (old_format) {
        const size_t wq = (size_t)div_up(width, 8);
        qweight = transpose_pack(read_file<uint32_t>(stem + ".qweight.bin", wq * height), height, width);
        qzeros = transpose_pack(read_file<uint32_t>(stem + ".qzeros.bin", wq * G), G, width);
        const std::vector<uint16_t> s = read_file<uint16_t>(stem + ".scales.bin", (size_t)G * width);
        for (int x = 0; x < width; x++)
            for (int g = 0; g < G; g++) scales[(size_t)x * G + g] = s[(size_t)g * width + x];
    } else {
        qweight = read_file<uint32_t>(stem + ".qweight.bin", (size_t)ph * width);
        qzeros = read_file<uint32_t>(stem + ".qzeros.bin", (size_t)zh * width);
        const int padded = zh * 8;            // the AWQ repo pads the scales of a column to a multiple of 8 groups
        const std::vector<uint16_t> s = read_file<uint16_t>(stem + ".scales.bin", (size_t)padded * width);
        for (int x = 0; x < width; x++)
            for (int g = 0; g < G; g++) scales[(size_t)x * G + g] = s[(size_t)x * padded + g];
    }
    write_all(out, qweight, stem);
    write_all(out, qzeros, stem);
    write_all(out, scales, stem);
}

}  // namespace

int main(int argc, char* argv[]) {
    if (argc != 5) {
        printf("usage: weight_packer <config.json from huggingface> <path_to_awq_bin_weights> <output_bin_filename> [OldAwqFormat: 0 or 1]\n");
        return 0;
    }
    const std::string dir = argv[2];
    const bool old_format = atoi(argv[4]) != 0;
    const Config c = read_config(argv[1]);
    FILE* out = fopen(argv[3], "wb+");
    if (!out) { printf("unable to open output file\n"); return 0; }
    if (fwrite(&c, sizeof c, 1, out) != 1) { printf("unable to write model metadata\n"); return 0; }

    copy_fp16(out, dir + "/model.embed_tokens.weight.bin", (size_t)c.vocab_size * c.dim);
    copy_fp16(out, dir + "/lm_head.weight.bin", (size_t)c.vocab_size * c.dim);
    copy_fp16(out, dir + "/model.norm.weight.bin", (size_t)c.dim);
    const int kv_dim = (c.dim * c.n_kv_heads) / c.n_heads;
    for (int l = 0; l < c.n_layers; l++) {
        printf("\nProcessing weights for layer: %d\n", l);
        const std::string base = dir + "/model.layers." + std::to_string(l);
        pack_matrix(out, base, "self_attn.q_proj", c.dim, c.dim, old_format);
        pack_matrix(out, base, "self_attn.k_proj", c.dim, kv_dim, old_format);
        pack_matrix(out, base, "self_attn.v_proj", c.dim, kv_dim, old_format);
        pack_matrix(out, base, "self_attn.o_proj", c.dim, c.dim, old_format);
        pack_matrix(out, base, "mlp.up_proj", c.dim, c.hidden_dim, old_format);      // up before gate: llama2_q4.cu:186-193
        pack_matrix(out, base, "mlp.gate_proj", c.dim, c.hidden_dim, old_format);
        pack_matrix(out, base, "mlp.down_proj", c.hidden_dim, c.dim, old_format);
        copy_fp16(out, base + ".input_layernorm.weight.bin", (size_t)c.dim);
        copy_fp16(out, base + ".post_attention_layernorm.weight.bin", (size_t)c.dim);
    }
    printf("\nDone!\n");
    fclose(out);
    return 0;
}

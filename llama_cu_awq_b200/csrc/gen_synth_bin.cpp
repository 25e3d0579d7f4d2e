// gen_synth_bin.cpp -- stand-alone writer of the seeded random-init `.bin` / `tokenizer.bin` files (synth.h): pure host C++,
// no CUDA, no engine.  bench.py's reference arm uses it so that timing the reference never maps libllama_q4_b200.so; the
// bytes are the ones lq4_write_synth_model / lq4_write_synth_tokenizer write (same header, same seed).
//   gen_synth_bin model <out.bin> <dim> <hidden_dim> <n_layers> <n_heads> <n_kv_heads> <vocab_size> <seq_len> <rope_theta> <seed>
//   gen_synth_bin tokenizer <out.bin> <vocab_size>
#include "synth.h"

int main(int argc, char** argv) {
    if (argc == 12 && !strcmp(argv[1], "model")) {
        synth::Cfg c;
        c.dim = atoi(argv[3]); c.hidden_dim = atoi(argv[4]); c.n_layers = atoi(argv[5]); c.n_heads = atoi(argv[6]);
        c.n_kv_heads = atoi(argv[7]); c.vocab_size = atoi(argv[8]); c.seq_len = atoi(argv[9]); c.rope_theta = (float)atof(argv[10]);
        const size_t n = synth::write_model(argv[2], c, strtoull(argv[11], nullptr, 0));
        printf("%zu\n", n);
        return n > 0 ? 0 : 1;
    }
    if (argc == 4 && !strcmp(argv[1], "tokenizer")) {
        const size_t n = synth::write_tokenizer(argv[2], atoi(argv[3]));
        printf("%zu\n", n);
        return n > 0 ? 0 : 1;
    }
    fprintf(stderr, "usage: %s model <out> dim hidden layers heads kv_heads vocab seq_len rope_theta seed | tokenizer <out> vocab\n", argv[0]);
    return 2;
}

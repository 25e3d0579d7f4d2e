// llama2_q4_b200.cpp -- drop-in command line for the sm_100a engine (boundary surface B2).
//
// Same invocation and the same stdout as the reference program (ankan-ban/llama_cu_awq llama2_q4.cu:604-720,
// generate() :436-492): `llama2_q4_b200 <checkpoint> [-n int] [-i str] [-f file] [-t float] [-p float]
// [-s int] [-z tokenizer] [-m generate|chat|perplexity] [-y sys] [-q dataset]` (all three modes), flags strictly `-x value`
// pairs, defaults temperature 0.5 / topp 0.6 / tokenizer.bin, and the closing
// `achieved tok/s: %f. Tokens: %d, seconds: %g` line.  Pure host C++ over the C ABI in
// include/llama_q4_b200.h; nothing here touches CUDA directly.
//
// The tokenizer is the llama2.c BPE the reference uses (tokenizer.h:35-223): byte-level pieces merged
// greedily by vocabulary score, `" "` dummy prefix, byte fallback at id+3, `<0xNN>` pieces decoded to raw
// bytes.  It is plumbing, written here with std containers.
//
// One difference from the reference's loop, invisible in the output: the position and the sampled token live
// on the device, so step t+1 is enqueued before the host has seen token t (LQ4_PIPELINE=0 restores the
// reference's launch-wait-launch order).
#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/llama_q4_b200.h"

namespace {

constexpr int kBos = 1, kEos = 2;   // tokenizer.h:8-9

struct Tokenizer {
    std::vector<std::string> vocab;
    std::vector<float> scores;
    std::unordered_map<std::string, int> index;   // first id wins, like bsearch over a stable set
    unsigned max_token_length = 0;

    void load(const char* path, int vocab_size) {
        FILE* f = fopen(path, "rb");
        if (!f) { fprintf(stderr, "couldn't load %s\n", path); exit(EXIT_FAILURE); }
        auto must = [&](size_t got) { if (got != 1) { fprintf(stderr, "failed read\n"); exit(EXIT_FAILURE); } };
        must(fread(&max_token_length, sizeof(int), 1, f));
        vocab.resize(vocab_size);
        scores.resize(vocab_size);
        for (int i = 0; i < vocab_size; i++) {
            int len = 0;
            must(fread(&scores[i], sizeof(float), 1, f));
            must(fread(&len, sizeof(int), 1, f));
            std::string s((size_t)len, '\0');
            if (len > 0) must(fread(&s[0], (size_t)len, 1, f));
            // the reference keeps C strings: an embedded NUL ends the piece
            s.resize(strlen(s.c_str()));
            vocab[i] = s;
        }
        fclose(f);
    }
    int lookup(const std::string& s) {
        if (index.empty())
            for (int i = (int)vocab.size() - 1; i >= 0; i--) index[vocab[i]] = i;
        auto it = index.find(s);
        return it == index.end() ? -1 : it->second;
    }
    // tokenizer.h:103-223
    std::vector<int> encode(const char* text, bool bos, bool eos) {
        std::vector<int> t;
        if (bos) t.push_back(kBos);
        if (text[0] != '\0') t.push_back(lookup(" "));
        std::string cp;
        for (const char* c = text; *c != '\0'; c++) {
            if ((*c & 0xC0) != 0x80) cp.clear();
            cp.push_back(*c);
            if ((*(c + 1) & 0xC0) == 0x80 && cp.size() < 4) continue;
            const int id = lookup(cp);
            if (id != -1) t.push_back(id);
            else for (unsigned char b : cp) t.push_back((int)b + 3);
            cp.clear();
        }
        for (;;) {
            float best_score = -1e10f;
            int best_id = -1, best_idx = -1;
            for (size_t i = 0; i + 1 < t.size(); i++) {
                const int id = lookup(vocab[t[i]] + vocab[t[i + 1]]);
                if (id != -1 && scores[id] > best_score) { best_score = scores[id]; best_id = id; best_idx = (int)i; }
            }
            if (best_idx < 0) break;
            t[best_idx] = best_id;
            t.erase(t.begin() + best_idx + 1);
        }
        if (eos) t.push_back(kEos);
        return t;
    }
    // tokenizer.h:67-79 + safe_printf :81-93
    void print_piece(int prev, int token) const {
        const char* piece = vocab[token].c_str();
        if (prev == kBos && piece[0] == ' ') piece++;
        unsigned char byte_val;
        char raw[2] = {0, 0};
        if (sscanf(piece, "<0x%02hhX>", &byte_val) == 1) { raw[0] = (char)byte_val; piece = raw; }
        if (piece[0] == '\0') return;
        if (piece[1] == '\0') {
            const unsigned char b = (unsigned char)piece[0];
            if (!(isprint(b) || isspace(b))) return;
        }
        printf("%s", piece);
    }
};

long time_in_ms() {   // llama2_q4.cu:400-405
    struct timespec t;
    timespec_get(&t, TIME_UTC);
    return t.tv_sec * 1000 + t.tv_nsec / 1000000;
}

[[noreturn]] void error_usage(char* argv[]) {   // llama2_q4.cu:604-619
    fprintf(stderr, "Usage:   %s <checkpoint> [options]\n", argv[0]);
    fprintf(stderr, "Example: %s model.bin -n 256 -i \"Write a poem on GPUs\"\n", argv[0]);
    fprintf(stderr, "Options:\n");
    fprintf(stderr, "  -n <int>    max number of steps to run for, default = max_seq_len\n");
    fprintf(stderr, "  -i <string> input prompt\n");
    fprintf(stderr, "  -f <string> path to file containing input prompt. Can be used with for multi-line prompts.\n");
    fprintf(stderr, "  -t <float>  temperature in [0,inf], default 0.5\n");
    fprintf(stderr, "  -p <float>  p value in top-p (nucleus) sampling in [0,1] default 0.9\n");
    fprintf(stderr, "  -s <int>    random seed, default time(NULL)\n");
    fprintf(stderr, "  -z <string> optional path to custom tokenizer\n");
    fprintf(stderr, "  -m <string> mode: generate|chat|perplexity, default: generate\n");
    fprintf(stderr, "  -y <string> (optional) system prompt in chat mode\n");
    fprintf(stderr, "  -q <string> dataset file for computing perplexity\n");
    exit(EXIT_FAILURE);
}

// Spin on the pinned doorbell the device-side sampler writes (SharedData::pos).  The spin is bounded: every so often the
// stream is queried, so that a launch that died (a CUDA error, or the kernel's own protocol time-out trap) ends the program
// with the error instead of hanging it; the reference's cudaStreamSynchronize loop (llama2_q4.cu:468) would return likewise.
void wait_doorbell(SharedData* sd, int pos) {
    for (unsigned spins = 0; sd->pos < pos; spins++) {
        if ((spins & 0xFFFFu) != 0xFFFFu) continue;
        const int q = lq4_stream_query();
        if (q == 2 || (q == 0 && sd->pos < pos)) {
            fprintf(stderr, "\nlq4: the decode step for position %d never published its token: %s\n", pos,
                    q == 2 ? lq4_last_error() : "the stream is idle");
            exit(EXIT_FAILURE);
        }
    }
}

void generate(Transformer* t, Tokenizer* tok, Sampler* sampler, const char* prompt, int steps) {
    if (prompt == nullptr) prompt = "";
    printf("\nEncoding Prompt... ");
    std::vector<int> prompt_tokens = tok->encode(prompt, true, false);
    printf("Done!\n");
    const int n_prompt = (int)prompt_tokens.size();
    if (n_prompt < 1) { fprintf(stderr, "something is wrong, expected at least 1 prompt token\n"); exit(EXIT_FAILURE); }

    const char* env = getenv("LQ4_PIPELINE");
    const bool pipelined = !(env && atoi(env) == 0);
    const int vocab = t->config.vocab_size;
    SharedData* sd = t->state.shared_data;

    // LQ4_DUMP_IDS=<file>: also write the id at every sequence position (parity tooling: ids without parsing the text)
    const char* dump_path = getenv("LQ4_DUMP_IDS");
    std::vector<int> ids(1, prompt_tokens[0]);

    const long start = time_in_ms();
    int token = prompt_tokens[0], pos = 0, launched = 0;
    lq4_reset(t, prompt_tokens.data(), n_prompt);
    // LQ4_PREFILL=1 (opt-in; the default keeps the reference's behaviour and its bit-identical ids): the prompt goes through the
    // batched tensor-core prefill (lq4_prefill) in one pass instead of one decode step per prompt token (llama2_q4.cu:465-470);
    // decoding then continues from the last prompt position over the prefilled KV cache.  Logits agree with the sequential
    // path to fp16 tolerance, so a near-tie can pick a different token.
    const char* pf_env = getenv("LQ4_PREFILL");
    if (pf_env && atoi(pf_env) != 0 && n_prompt >= 2 && n_prompt <= steps &&
        lq4_prefill(t, prompt_tokens.data(), 1, n_prompt, 0, nullptr, nullptr, nullptr) == 0) {
        for (int p = 1; p < n_prompt; p++) {          // the echo of the prompt the stepped loop would have printed
            tok->print_piece(token, prompt_tokens[p]);
            ids.push_back(prompt_tokens[p]);
            token = prompt_tokens[p];
        }
        pos = launched = n_prompt - 1;
    }
    while (pos < steps) {
        if (pipelined) {
            // keep up to two steps in flight; step s is complete once the device has published pos == s+1
            while (launched < steps && launched <= pos + 1) {
                lq4_enqueue_step(t, sampler, launched + 1, launched >= n_prompt - 1);
                launched++;
            }
            if (pos > 0) wait_doorbell(sd, pos);
        } else {
            lq4_stream_synchronize();
            lq4_run_transformer(pos >= n_prompt - 1, &t->config, &t->state, &t->weights, 0, sampler);
        }
        if (pos > 0) {
            int next = sd->tokens[pos];
            if (next >= vocab) next = 0;
            tok->print_piece(token, next);
            ids.push_back(next);
            if (next == kEos) break;
            token = next;
        }
        pos++;
    }
    printf("\n");
    lq4_stream_synchronize();
    const long end = time_in_ms();
    const double secs = (end - start) / 1000.0;
    const int timed_tokens = pos - 1;
    printf("\nachieved tok/s: %f. Tokens: %d, seconds: %g\n", timed_tokens / secs, timed_tokens, secs);
    if (dump_path != nullptr) {
        if (FILE* f = fopen(dump_path, "w")) {
            fprintf(f, "%d\n", n_prompt);
            for (int id : ids) fprintf(f, "%d\n", id);
            fclose(f);
        }
    }
}

// ---------------------------------------------------------------------------- chat mode (llama2_q4.cu:494-601)
void read_stdin(const char* guide, char* buffer, size_t bufsize) {
    printf("%s", guide);
    if (fgets(buffer, (int)bufsize, stdin) != NULL) {
        size_t len = strlen(buffer);
        if (len > 0 && buffer[len - 1] == '\n') buffer[len - 1] = '\0';
    } else {
        buffer[0] = '\0';
    }
}

void chat(Transformer* t, Tokenizer* tok, Sampler* sampler, const char* cli_user_prompt, const char* cli_system_prompt, int steps) {
    char system_prompt[512] = {0}, user_prompt[512] = {0};
    std::string rendered;
    std::vector<int> prompt_tokens;
    int num_prompt_tokens = 0, user_idx = 0;
    bool user_turn = true;
    int next = 0, token = 0, pos = 0;
    SharedData* sd = t->state.shared_data;
    lq4_reset(t, &pos, 0);                       // device and host position to 0, no tokens yet
    while (pos < steps) {
        if (user_turn) {
            if (pos == 0) {
                if (cli_system_prompt == nullptr) read_stdin("Enter system prompt (optional): ", system_prompt, sizeof system_prompt);
                else snprintf(system_prompt, sizeof system_prompt, "%s", cli_system_prompt);
            }
            if (pos == 0 && cli_user_prompt != nullptr) snprintf(user_prompt, sizeof user_prompt, "%s", cli_user_prompt);
            else {
                if (feof(stdin)) break;          // the reference would spin on an empty prompt; end the dialog instead
                read_stdin("User: ", user_prompt, sizeof user_prompt);
            }
            if (pos == 0 && system_prompt[0] != '\0')
                rendered = std::string("[INST] <<SYS>>\n") + system_prompt + "\n<</SYS>>\n\n" + user_prompt + " [/INST]";
            else
                rendered = std::string("[INST] ") + user_prompt + " [/INST]";
            printf("\nRendered prompt: %s\n", rendered.c_str());
            prompt_tokens = tok->encode(rendered.c_str(), true, false);
            num_prompt_tokens = (int)prompt_tokens.size();
            if (pos + num_prompt_tokens >= LQ4_MAX_SEQ_LEN) break;
            user_idx = 0;
            user_turn = false;
            printf("Assistant: ");
            lq4_stream_synchronize();
            memcpy((void*)&sd->tokens[pos], prompt_tokens.data(), sizeof(int) * num_prompt_tokens);
        }
        lq4_stream_synchronize();
        lq4_run_transformer(user_idx >= num_prompt_tokens - 1, &t->config, &t->state, &t->weights, 0, sampler);
        user_idx++;
        if (user_idx > 0) {
            next = sd->tokens[pos];              // output of the previous iteration (or the prompt token at this position)
            if (next == kEos) {
                user_turn = true;
                printf("\n");
            } else if (user_idx > num_prompt_tokens) {
                if (next >= 0 && next < t->config.vocab_size) tok->print_piece(token, next);
                fflush(stdout);
            }
            token = next;
        }
        pos++;
    }
    printf("\n");
    lq4_stream_synchronize();
}

// ---------------------------------------------------------------------------- perplexity mode (perplexity.h)
float compute_perplexity(const int* tokens, float* logits, int num_tokens, int vocab_size) {   // perplexity.h:3-50
    double sum = 0.0;
    for (int i = 0; i < num_tokens; i++) {
        float* x = logits + (size_t)i * vocab_size;
        float max_val = x[0];
        for (int v = 1; v < vocab_size; v++) if (x[v] > max_val) max_val = x[v];
        float s = 0.0f;
        for (int v = 0; v < vocab_size; v++) { x[v] = expf(x[v] - max_val); s += x[v]; }
        for (int v = 0; v < vocab_size; v++) x[v] /= s;
        sum += log((double)x[tokens[i]]);
    }
    return (float)exp(-sum / num_tokens);
}

float dataset_perplexity(const char* dataset, Transformer* t, Tokenizer* tok, Sampler* sampler) {   // perplexity.h:57-96
    SharedData* sd = t->state.shared_data;
    printf("\nTokenizing Dataset...");
    std::vector<int> toks = tok->encode(dataset, false, false);
    printf("done!\n");
    printf("Found %d characters, %d tokens", (int)strlen(dataset), (int)toks.size());
    int num = (int)toks.size();
    if (num >= t->config.seq_len) { num = t->config.seq_len - 1; printf("\nTruncated to %d tokens", num); }
    printf("\nRunning the network to get logits...");
    int zero = 0;
    lq4_reset(t, &zero, 0);
    sd->tokens[0] = kBos;
    memcpy((void*)&sd->tokens[1], toks.data(), sizeof(int) * (size_t)std::min<size_t>(toks.size(), LQ4_MAX_SEQ_LEN - 2));
    for (int pos = 0; pos < num; pos++) {
        lq4_run_transformer(0, &t->config, &t->state, &t->weights, 1, sampler);
        lq4_stream_synchronize();
    }
    printf("done!\n");
    printf("Computing perplexity...");
    std::vector<float> logits((size_t)num * t->config.vocab_size);
    if (num > 0 && lq4_memcpy_to_host(logits.data(), t->state.logits_array, logits.size() * sizeof(float)) != 0) exit(EXIT_FAILURE);
    const float pplx = num > 0 ? compute_perplexity(&toks[0], logits.data(), num, t->config.vocab_size) : 1.0f;
    printf("\nPerplexity computed on %d tokens: %f\n\n", num, pplx);
    return pplx;
}

void parse_dataset_and_compute_perplexity(const char* path, Transformer* t, Tokenizer* tok, Sampler* sampler) {   // perplexity.h:99-139
    FILE* fp = path ? fopen(path, "rb") : nullptr;
    if (!fp) { printf("Couldn't open file %s\n", path ? path : "(null)"); exit(1); }
    printf("\nLoading Dataset...");
    fseek(fp, 0, SEEK_END);
    const long bytes = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    std::string data((size_t)bytes, '\0');
    if (bytes > 0 && fread(&data[0], 1, (size_t)bytes, fp) != (size_t)bytes) { printf("error reading dataset\n"); exit(1); }
    fclose(fp);
    printf("done!\n");
    int count = 0;
    double product = 1;
    size_t cur = 0;
    const std::string sep = "<|endoftext|>";
    for (;;) {
        const size_t nxt = data.find(sep, cur);
        const std::string seq = data.substr(cur, nxt == std::string::npos ? std::string::npos : nxt - cur);
        product *= dataset_perplexity(seq.c_str(), t, tok, sampler);
        count++;
        if (nxt == std::string::npos) break;
        cur = nxt + sep.size();
    }
    printf("\nGeomean perplexity on %d sequences: %f\n\n", count, pow(product, 1.0 / count));
}

}  // namespace

int main(int argc, char* argv[]) {
    const char* checkpoint_path = nullptr;
    const char* tokenizer_path = "tokenizer.bin";
    const char* dataset_path = nullptr;
    int steps = 0;
    char* prompt = nullptr;
    const char* system_prompt = nullptr;
    float temperature = 0.5f, topp = 0.6f;   // llama2_q4.cu:632-633
    unsigned long long rng_seed = 0;
    const char* mode = "generate";

    if (argc >= 2) checkpoint_path = argv[1]; else error_usage(argv);
    for (int i = 2; i < argc; i += 2) {
        if (i + 1 >= argc) error_usage(argv);
        if (argv[i][0] != '-') error_usage(argv);
        if (strlen(argv[i]) != 2) error_usage(argv);
        switch (argv[i][1]) {
            case 'n': steps = atoi(argv[i + 1]); break;
            case 'i': prompt = argv[i + 1]; break;
            case 'z': tokenizer_path = argv[i + 1]; break;
            case 't': temperature = (float)atof(argv[i + 1]); break;
            case 'p': topp = (float)atof(argv[i + 1]); break;
            case 's': rng_seed = (unsigned long long)atoi(argv[i + 1]); break;
            case 'm': mode = argv[i + 1]; break;
            case 'y': system_prompt = argv[i + 1]; break;
            case 'q': dataset_path = argv[i + 1]; break;
            case 'f': {
                FILE* file = fopen(argv[i + 1], "r");
                if (!file) { printf("Couldn't open file %s\n", argv[i + 1]); exit(1); }
                fseek(file, 0, SEEK_END);
                const long fsize = ftell(file);
                fseek(file, 0, SEEK_SET);
                if (prompt) printf("Warning: -f overrides -i\n");
                prompt = (char*)malloc((size_t)fsize + 1);
                if (fread(prompt, (size_t)fsize, 1, file) != 1 && fsize > 0) { printf("Couldn't read file %s\n", argv[i + 1]); exit(1); }
                fclose(file);
                prompt[fsize] = 0;
                break;
            }
            default: error_usage(argv);
        }
    }
    const bool perplexity = strcmp(mode, "perplexity") == 0;
    if (rng_seed <= 0) rng_seed = (unsigned int)time(NULL);
    if (temperature < 0.0f) temperature = 0.0f;
    if (topp < 0.0f || 1.0f < topp) topp = 0.9f;
    if (!perplexity && dataset_path) printf("Warning: dataset path is ignored in non-perplexity mode\n");
    if (strcmp(mode, "generate") != 0 && strcmp(mode, "chat") != 0 && !perplexity) error_usage(argv);

    if (lq4_init(0) != 0) { fprintf(stderr, "%s\n", lq4_last_error()); return EXIT_FAILURE; }
    Transformer transformer;
    lq4_build_transformer(&transformer, checkpoint_path, perplexity);
    if (steps <= 0 || steps > transformer.config.seq_len) steps = transformer.config.seq_len;
    Tokenizer tokenizer;
    tokenizer.load(tokenizer_path, transformer.config.vocab_size);
    Sampler sampler;
    lq4_build_sampler(&sampler, transformer.config.vocab_size, temperature, topp, rng_seed);

    if (perplexity) parse_dataset_and_compute_perplexity(dataset_path, &transformer, &tokenizer, &sampler);
    else if (strcmp(mode, "generate") == 0) generate(&transformer, &tokenizer, &sampler, prompt, steps);
    else chat(&transformer, &tokenizer, &sampler, prompt, system_prompt, steps);

    lq4_destroy_sampler(&sampler);
    lq4_free_transformer(&transformer);
    return 0;
}

// oracle/ref_harness.cu -- C-ABI shim around the UNMODIFIED reference translation unit.
//
// TEST INFRASTRUCTURE ONLY.  Built by oracle/build_ref.sh with -I/root/reference into
// oracle/_ref/libq4ref.so (git-ignored).  The reference source is #included from where it lies;
// nothing is copied.  The shim only renames the reference `main` and forwards to the reference's
// own host wrappers (llama2_q4.cu:207-395) so tests can run the stock kernels on device buffers
// they own, and step the stock run_transformer while reading back its logits.
#define main ref_main
#include "llama2_q4.cu"
#undef main

static Transformer g_ref_t;
static Sampler g_ref_sampler;
static bool g_ref_open = false;

static void ref_ensure_stream() {
    if (!stream) cudaStreamCreate(&stream);           // reference main does this at llama2_q4.cu:700
}

extern "C" {

int ref_last_cuda_error() { return (int)cudaGetLastError(); }

// ---- op-level: stock host wrappers on caller-owned device memory (synchronous) ----
void ref_rmsnorm(half* o, half* x, half* w, int size) {
    ref_ensure_stream();
    rmsnorm(o, x, w, size);
    cudaStreamSynchronize(stream);
}
void ref_matmul_fp16(half* xout, half* x, half* w, int n, int d) {
    ref_ensure_stream();
    matmul(xout, x, w, n, d);
    cudaStreamSynchronize(stream);
}
void ref_matmul_q4(half* xout, half* x, uint32_t* qw, uint32_t* qz, half* sc, int inpSize, int opSize,
                   int accum, int loff, int* pPos) {
    ref_ensure_stream();
    QWeight w{qw, qz, sc};
    matmul(xout, x, w, inpSize, opSize, accum != 0, loff, pPos);
    cudaStreamSynchronize(stream);
}
void ref_qkv_matvec(half* q, half* kc, half* vc, half* x, uint32_t* qw, uint32_t* qz, half* qs,
                    uint32_t* kw, uint32_t* kz, half* ks, uint32_t* vw, uint32_t* vz, half* vs,
                    int inpSize, int opSize, int loff, int* pPos) {
    ref_ensure_stream();
    QWeight a{qw, qz, qs}, b{kw, kz, ks}, c{vw, vz, vs};
    qkv_matvec(q, kc, vc, x, a, b, c, inpSize, opSize, loff, pPos);
    cudaStreamSynchronize(stream);
}
void ref_ffn_matvec_silu(half* out, half* x, uint32_t* gw, uint32_t* gz, half* gs, uint32_t* uw,
                         uint32_t* uz, half* us, int inpSize, int opSize) {
    ref_ensure_stream();
    QWeight g{gw, gz, gs}, u{uw, uz, us};
    ffn_matvec_silu(out, x, g, u, inpSize, opSize);
    cudaStreamSynchronize(stream);
}
void ref_rope(half* q, half* k, int num_heads, int num_kv_heads, int head_size, int* pPos, int loff,
              float theta) {
    ref_ensure_stream();
    RoPERotation(q, k, num_heads, num_kv_heads, head_size, pPos, loff, theta);
    cudaStreamSynchronize(stream);
}
void ref_mha(half* out, half* q, half* kc, half* vc, half* att, int num_heads, int head_size, int kv_mul,
             int max_seq_len, int* pPos) {
    ref_ensure_stream();
    MultiHeadAttention(out, q, kc, vc, att, num_heads, head_size, kv_mul, max_seq_len, pPos);
    cudaStreamSynchronize(stream);
}

// ---- model-level: stock build_transformer / run_transformer ----
int ref_open(const char* bin_path) {
    if (g_ref_open) return -1;
    build_transformer(&g_ref_t, (char*)bin_path, false);
    build_sampler(&g_ref_sampler, g_ref_t.config.vocab_size, 0.0f, 0.9f, 1234ull);   // greedy
    ref_ensure_stream();
    cudaDeviceSynchronize();
    g_ref_open = true;
    return 0;
}
void ref_config(int* out8) { memcpy(out8, &g_ref_t.config, sizeof(Config)); }
// switch the stock sampler to temperature / top-p sampling with a fixed seed (main() does this from -t -p -s)
void ref_set_sampler(float temperature, float topp, unsigned long long seed) {
    g_ref_sampler.temperature = temperature;
    g_ref_sampler.topp = topp;
    g_ref_sampler.rng_state = seed;
}
// start a sequence: tokens become the prompt (generate(), llama2_q4.cu:461-463)
void ref_reset(const int* tokens, int n) {
    cudaMemset(g_ref_t.state.pos, 0, sizeof(int));
    g_ref_t.state.shared_data->pos = 0;
    memcpy((void*)g_ref_t.state.shared_data->tokens, tokens, sizeof(int) * n);
}
// one forward + sample, exactly the loop body at llama2_q4.cu:468-470; returns the new pos.
// logits_out (host, vocab halfs) may be null; next_token_out receives tokens[new_pos].
int ref_step(int gen_token, half* logits_out, int* next_token_out) {
    cudaStreamSynchronize(stream);
    run_transformer(gen_token != 0, &g_ref_t.config, &g_ref_t.state, &g_ref_t.weights, false, &g_ref_sampler);
    cudaStreamSynchronize(stream);
    int pos = g_ref_t.state.shared_data->pos;
    if (logits_out)
        cudaMemcpy(logits_out, g_ref_t.state.logits, sizeof(half) * g_ref_t.config.vocab_size, cudaMemcpyDeviceToHost);
    if (next_token_out) *next_token_out = g_ref_t.state.shared_data->tokens[pos];
    return pos;
}
// Steady-state timing of the stock loop body (llama2_q4.cu:465-470: cudaStreamSynchronize, then run_transformer) for `steps`
// positions from the current one, with CUDA events on the reference's own stream.  Call it after a pass over the same
// positions, so that every length bin's graph is already captured (llama2_q4.cu:362-371).  Returns milliseconds.
float ref_time_steps(int steps, int gen_from) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaStreamSynchronize(stream);
    cudaEventRecord(e0, stream);
    for (int i = 0; i < steps; i++) {
        cudaStreamSynchronize(stream);
        run_transformer(i >= gen_from, &g_ref_t.config, &g_ref_t.state, &g_ref_t.weights, false, &g_ref_sampler);
    }
    cudaEventRecord(e1, stream);
    cudaEventSynchronize(e1);
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return ms;
}
// token at sequence position i (pinned SharedData::tokens), after the stream has been synchronised
int ref_token_at(int i) { cudaStreamSynchronize(stream); return g_ref_t.state.shared_data->tokens[i]; }
// device pointers of the reference run state, for bit-compares of intermediates
void* ref_state_ptr(int which) {
    RunState* s = &g_ref_t.state;
    switch (which) {
        case 0: return s->x;
        case 1: return s->xb;
        case 2: return s->hb;
        case 3: return s->q;
        case 4: return s->att;
        case 5: return s->logits;
        case 6: return s->key_cache;
        case 7: return s->value_cache;
        default: return nullptr;
    }
}
void ref_close() {
    if (!g_ref_open) return;
    cudaStreamSynchronize(stream);
    free_transformer(&g_ref_t);
    // the reference caches captured graphs in globals (llama2_q4.cu:342-344) and only drops them at the end
    // of main (:713-716); do the same here, or the next model would replay graphs holding freed pointers
    for (int i = 0; i < MAX_GRAPHS; i++)
        if (graphCaptured[i]) { cudaGraphExecDestroy(cudaGraphInstance[i]); graphCaptured[i] = false; }
    g_ref_open = false;
}

}  // extern "C"

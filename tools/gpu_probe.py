#!/usr/bin/env python
"""GPU probe (development aid, run under gpurun): where does a decode step's time go?
Prints per-kernel cold (HBM) and hot (L2-resident) timings of the C-ABI operators on the 7B shapes, and
whole-step timings under the engine options.  Output: JSON lines into gpurun_out/probe.jsonl."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import llama_cu_awq_b200 as E
import bench as B

lib = E.lib()
assert lib.lq4_init(0) == 0
model = sys.argv[1] if len(sys.argv) > 1 else "7b"
cfg = B.model_cfg(model)
path, tok = B.ensure_files(lib, E, model, cfg)
t = E.Transformer()
lib.lq4_build_transformer(C.byref(t), path.encode(), 0)
s = E.Sampler()
lib.lq4_build_sampler(C.byref(s), cfg["vocab_size"], 0.0, 0.9, 1)
stream = torch.cuda.ExternalStream(lib.lq4_get_stream())
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = open(os.path.join(ROOT, "gpurun_out", "probe.jsonl"), "a")
d, h, L, V = cfg["dim"], cfg["hidden_dim"], cfg["n_layers"], cfg["vocab_size"]
layers = t.weights.layers
st = t.state
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def qbytes(K, N):
    return N * (E.packed_weight_height(K) * 4 + E.packed_zeros_height(K) * 4 + E.num_groups(K) * 2)


def timeit(fn, n):
    fn(); fn()
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        ev0.record(stream)
    for _ in range(n):
        fn()
    with torch.cuda.stream(stream):
        ev1.record(stream)
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) * 1000.0 / n   # us


def emit(**kw):
    print(json.dumps(kw)); out.write(json.dumps(kw) + "\n"); out.flush()


pos = torch.zeros(1, dtype=torch.int32, device="cuda")
ops = {
    "qkv": (lambda l: lib.lq4_qkv_matvec(st.q, st.key_cache, st.value_cache, st.xb, C.byref(layers[l].wq_q), C.byref(layers[l].wq_k),
                                         C.byref(layers[l].wq_v), d, d, 0, pos.data_ptr()), 3 * qbytes(d, d)),
    "o": (lambda l: lib.lq4_matmul_q4(st.x, st.xb, C.byref(layers[l].wq_o), d, d, 1, -1, None), qbytes(d, d)),
    "ffn": (lambda l: lib.lq4_ffn_matvec_silu(st.hb, st.xb, C.byref(layers[l].wq_gate), C.byref(layers[l].wq_up), d, h), 2 * qbytes(d, h)),
    "down": (lambda l: lib.lq4_matmul_q4(st.x, st.hb, C.byref(layers[l].wq_down), h, d, 1, -1, None), qbytes(h, d)),
}
for name, (fn, nbytes) in ops.items():
    cold = timeit(lambda: [fn(l) for l in range(L)], 3) / L
    hot = timeit(lambda: fn(0), 50)
    emit(kind="op", op=name, bytes=nbytes, cold_us=cold, cold_gbs=nbytes / cold / 1e3, hot_us=hot, hot_gbs=nbytes / hot / 1e3)
cls_bytes = V * d * 2
cold = timeit(lambda: lib.lq4_matmul_fp16(st.logits, st.x, t.weights.wcls, d, V, 1, 0, 0, 0, -1, 1.0), 5)
emit(kind="op", op="classifier", bytes=cls_bytes, cold_us=cold, cold_gbs=cls_bytes / cold / 1e3)
for p in (0, 127, 255, 1023):
    pos.fill_(p); torch.cuda.synchronize()
    us = timeit(lambda: lib.lq4_multi_head_attention(st.xb, st.q, st.key_cache, st.value_cache, None, cfg["n_heads"], d // cfg["n_heads"], 1,
                                                     2048, pos.data_ptr()), 20)
    emit(kind="op", op="attention", pos=p, us=us)
us = timeit(lambda: lib.lq4_rmsnorm(st.xb, st.x, layers[0].rms_att_weight, d), 50)
emit(kind="op", op="rmsnorm", us=us)

bos = (C.c_int * 1)(1)
K = 256
wbytes = E.weight_bytes_per_token(cfg)
kv = sum(E.kv_bytes_at(cfg, p) for p in range(K)) / K
for (fused, pdl, graphs) in ((1, 1, 1), (1, 0, 1), (1, 1, 0), (0, 0, 1), (0, 0, 0)):
    lib.lq4_set_option(b"fused", fused); lib.lq4_set_option(b"pdl", pdl); lib.lq4_set_option(b"graphs", graphs)

    def run():
        lib.lq4_reset(C.byref(t), bos, 1)
        for i in range(K):
            lib.lq4_enqueue_step(C.byref(t), C.byref(s), i + 1, 1)
    us = timeit(run, 2) / K
    toks = [int(t.state.shared_data.contents.tokens[i]) for i in range(1, 9)]
    emit(kind="step", fused=fused, pdl=pdl, graphs=graphs, us_per_token=us, tok_s=1e6 / us, gbs=(wbytes + kv) / us / 1e3, first_tokens=toks)
lib.lq4_free_transformer(C.byref(t))

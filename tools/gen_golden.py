#!/usr/bin/env python
"""Generates tests/golden/*.npz ON THE GPU BOX from the UNMODIFIED reference CUDA build
(oracle/_ref/libq4ref.so = the reference translation unit behind a C shim, oracle/ref_harness.cu):

    gpurun -- 'python tools/gen_golden.py'      # writes gpurun_out/golden/*.npz; copy them to tests/golden/

The reference ships no golden vectors of its own, so these are outputs of the
reference itself on seeded inputs: op-level vectors (inputs stored beside the outputs) and, for the
tiny synthetic model (seed in the file), the fp16 logits and greedy token ids of a 12-step run.
tests/test_golden.py pins the CPU oracle (no GPU needed) and the engine (GPU) against them."""
import ctypes as C
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import helpers as H
import llama_cu_awq_b200 as E

OUT = os.path.join(ROOT, "gpurun_out", "golden")
os.makedirs(OUT, exist_ok=True)
r = H.ref()
assert r is not None, "oracle/_ref/libq4ref.so missing: run oracle/build_ref.sh where /root/reference exists"
lib = E.lib()
rng = np.random.default_rng(20261017)


def dev(a):
    return H.to_dev(a)


def u16(a):
    return np.ascontiguousarray(a).view(np.uint16)


def dpos(p):
    t = torch.tensor([p], dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    return t


# ---- INT4 GEMV (mat_vec_kernel_int4), plain and residual-accumulate, incl. a ragged last trip ----
for K, N in [(256, 64), (1024, 32), (1408, 16)]:
    w, z, s = H.random_qweight(rng, K, N)
    x = rng.standard_normal(K).astype(np.float16)
    out0 = rng.standard_normal(N).astype(np.float16)
    res = {}
    for accum in (0, 1):
        dw, dz, ds, dx, do = dev(w), dev(z), dev(u16(s)), dev(u16(x)), dev(u16(out0))
        r.ref_matmul_q4(do.data_ptr(), dx.data_ptr(), dw.data_ptr(), dz.data_ptr(), ds.data_ptr(), K, N, accum, -1, None)
        torch.cuda.synchronize()
        res[accum] = H.dev_u16(do)
    np.savez_compressed(os.path.join(OUT, f"gemv_q4_K{K}_N{N}.npz"), w=w, z=z, s=u16(s), x=u16(x), out0=u16(out0), y=res[0], y_accum=res[1])

# ---- gate/up + SiLU (ffn_matvec_silu_kernel) ----
K, N = 256, 48
g, u = H.random_qweight(rng, K, N, 4.0), H.random_qweight(rng, K, N, 4.0)
x = rng.standard_normal(K).astype(np.float16)
dg = [dev(g[0]), dev(g[1]), dev(u16(g[2]))]
du = [dev(u[0]), dev(u[1]), dev(u16(u[2]))]
dx, do = dev(u16(x)), dev(np.zeros(N, np.uint16))
r.ref_ffn_matvec_silu(do.data_ptr(), dx.data_ptr(), *[t.data_ptr() for t in dg], *[t.data_ptr() for t in du], K, N)
torch.cuda.synchronize()
np.savez_compressed(os.path.join(OUT, "ffn_silu_K256_N48.npz"), gw=g[0], gz=g[1], gs=u16(g[2]), uw=u[0], uz=u[1], us=u16(u[2]), x=u16(x), y=H.dev_u16(do))

# ---- fp16 GEMV (mat_vec_kernel) ----
n, d = 512, 40
w = (rng.standard_normal((d, n)) * 0.05).astype(np.float16)
x = rng.standard_normal(n).astype(np.float16)
dw, dx, do = dev(u16(w)), dev(u16(x)), dev(np.zeros(d, np.uint16))
r.ref_matmul_fp16(do.data_ptr(), dx.data_ptr(), dw.data_ptr(), n, d)
torch.cuda.synchronize()
np.savez_compressed(os.path.join(OUT, "gemv_f16_n512_d40.npz"), w=u16(w), x=u16(x), y=H.dev_u16(do))

# ---- RMSNorm ----
size = 1000
x = (rng.standard_normal(size) * 3).astype(np.float16)
w = (0.9 + 0.2 * rng.random(size)).astype(np.float16)
dx, dw, do = dev(u16(x)), dev(u16(w)), dev(np.zeros(size, np.uint16))
r.ref_rmsnorm(do.data_ptr(), dx.data_ptr(), dw.data_ptr(), size)
torch.cuda.synchronize()
np.savez_compressed(os.path.join(OUT, "rmsnorm_1000.npz"), x=u16(x), w=u16(w), y=H.dev_u16(do))

# ---- RoPE (q of 4 heads and the cache row of 2 kv heads, head size 64) ----
nh, nkv, hs, pos, theta = 4, 2, 64, 37, 10000.0
q = rng.standard_normal(nh * hs).astype(np.float16)
krow = rng.standard_normal(nkv * hs).astype(np.float16)
cache = np.zeros((pos + 1) * nkv * hs, np.float16)
cache[pos * nkv * hs:] = krow
dq, dk, dp = dev(u16(q)), dev(u16(cache)), dpos(pos)
r.ref_rope(dq.data_ptr(), dk.data_ptr(), nh, nkv, hs, dp.data_ptr(), 0, theta)
torch.cuda.synchronize()
np.savez_compressed(os.path.join(OUT, "rope_h4_kv2_hs64_pos37.npz"), q=u16(q), k=u16(krow), q_out=H.dev_u16(dq), k_out=H.dev_u16(dk)[pos * nkv * hs:])

# ---- attention (QK^T, softmax, PV) ----
nh, hs, kv_mul, pos, seq = 4, 64, 2, 40, 64
kv_dim = nh * hs // kv_mul
q = rng.standard_normal(nh * hs).astype(np.float16)
kc = rng.standard_normal((seq, kv_dim)).astype(np.float16)
vc = rng.standard_normal((seq, kv_dim)).astype(np.float16)
dq, dk, dv = dev(u16(q)), dev(u16(kc)), dev(u16(vc))
datt, do, dp = dev(np.zeros(nh * seq, np.uint16)), dev(np.zeros(nh * hs, np.uint16)), dpos(pos)
r.ref_mha(do.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), datt.data_ptr(), nh, hs, kv_mul, seq, dp.data_ptr())
torch.cuda.synchronize()
np.savez_compressed(os.path.join(OUT, "attention_h4_hs64_pos40.npz"), q=u16(q), k=u16(kc), v=u16(vc), y=H.dev_u16(do), att=H.dev_u16(datt)[:nh * (pos + 1)])

# ---- whole model: tiny synthetic .bin (lq4_write_synth_model, seed 2024), greedy run ----
cfg = dict(H.TINY)
seed, steps, prompt = 2024, 12, [1, 35, 72]
with tempfile.TemporaryDirectory() as dd:
    path = os.path.join(dd, "tiny.bin")
    c = E.Config(**cfg)
    assert lib.lq4_write_synth_model(path.encode(), C.byref(c), seed) == os.path.getsize(path)
    assert r.ref_open(path.encode()) == 0
    toks = np.array(prompt, dtype=np.int32)
    r.ref_reset(toks.ctypes.data_as(C.c_void_p), len(toks))
    logits = np.zeros((steps, cfg["vocab_size"]), np.uint16)
    out = [prompt[0]]
    nxt = C.c_int(0)
    for st in range(steps):
        r.ref_step(1 if st >= len(prompt) - 1 else 0, logits[st].ctypes.data_as(C.c_void_p), C.byref(nxt))
        out.append(int(nxt.value))
    r.ref_close()
np.savez_compressed(os.path.join(OUT, "tiny_model_seed2024.npz"), cfg=np.array([cfg[k] for k in ("dim", "hidden_dim", "n_layers", "n_heads", "n_kv_heads", "vocab_size", "seq_len")], np.int32),
                    rope_theta=np.float32(cfg["rope_theta"]), seed=np.int64(seed), prompt=np.array(prompt, np.int32), logits=logits, tokens=np.array(out, np.int32))
print("golden vectors written to", OUT, sorted(os.listdir(OUT)))

"""Golden vectors produced by the UNMODIFIED reference CUDA build on a B200 (tools/gen_golden.py, which
drives oracle/_ref/libq4ref.so) pin both checkers:

  * CPU oracle vs golden (no GPU): bit-exact where no transcendental is involved (INT4 GEMV, fp16 GEMV,
    RMSNorm); <= 2 fp16 ulps where host libm stands in for CUDA's expf/sinf/cosf/powf (SiLU, RoPE,
    softmax) -- the tolerance is the oracle's, the engine below is held to zero;
  * engine vs golden (GPU, through the C ABI): bit-exact, every vector.

The reference itself ships no test vectors."""
import ctypes as C
import glob
import os
import tempfile

import numpy as np
import pytest

import helpers as H

G = H.GOLDEN


def load(name):
    return np.load(os.path.join(G, name))


def test_golden_files_present():
    names = sorted(os.path.basename(p) for p in glob.glob(os.path.join(G, "*.npz")))
    assert len(names) >= 9, names


# ------------------------------------------------------------------ oracle vs golden (CPU)
@pytest.mark.parametrize("name,K,N", [("gemv_q4_K256_N64.npz", 256, 64), ("gemv_q4_K1024_N32.npz", 1024, 32), ("gemv_q4_K1408_N16.npz", 1408, 16)])
def test_oracle_gemv_q4(name, K, N):
    g, o = load(name), H.oracle()
    w, z, s, x = (np.ascontiguousarray(g[k]) for k in ("w", "z", "s", "x"))
    qw = H.OracleQWeight(H.ptr(w), H.ptr(z), H.ptr(s))
    for accum, key in ((0, "y"), (1, "y_accum")):
        out = g["out0"].copy()
        o.oracle_matvec_int4(H.ptr(out), H.ptr(x), qw, K, N, accum)
        assert (out == g[key]).all(), f"{name} accum={accum}: oracle differs from the reference kernel"


def test_oracle_gemv_f16():
    g, o = load("gemv_f16_n512_d40.npz"), H.oracle()
    out = np.zeros(40, np.uint16)
    o.oracle_matvec_fp16(H.ptr(out), H.ptr(np.ascontiguousarray(g["x"])), H.ptr(np.ascontiguousarray(g["w"])), 512, 40, 1.0)
    assert (out == g["y"]).all()


def test_oracle_rmsnorm():
    g, o = load("rmsnorm_1000.npz"), H.oracle()
    out = np.zeros(1000, np.uint16)
    o.oracle_rmsnorm(H.ptr(out), H.ptr(np.ascontiguousarray(g["x"])), H.ptr(np.ascontiguousarray(g["w"])), 1000)
    assert (out == g["y"]).all()


def test_oracle_ffn_silu():
    g, o = load("ffn_silu_K256_N48.npz"), H.oracle()
    gw, gz, gs, uw, uz, us, x = (np.ascontiguousarray(g[k]) for k in ("gw", "gz", "gs", "uw", "uz", "us", "x"))
    out = np.zeros(48, np.uint16)
    o.oracle_ffn_matvec_silu(H.ptr(out), H.ptr(x), H.OracleQWeight(H.ptr(gw), H.ptr(gz), H.ptr(gs)),
                             H.OracleQWeight(H.ptr(uw), H.ptr(uz), H.ptr(us)), 256, 48)
    assert H.ulp_diff_f16(out, g["y"]).max() <= 1       # expf: host libm vs CUDA libdevice


def test_oracle_rope():
    g, o = load("rope_h4_kv2_hs64_pos37.npz"), H.oracle()
    q, k = g["q"].copy(), g["k"].copy()
    o.oracle_rope(H.ptr(q), H.ptr(k), 4, 2, 64, 37, 10000.0)
    for got, want, src in ((q, g["q_out"], g["q"]), (k, g["k_out"], g["k"])):
        fa, fb = got.view(np.float16).astype(np.float32), want.view(np.float16).astype(np.float32)
        scale = float(np.abs(src.view(np.float16).astype(np.float32)).max())
        assert np.abs(fa - fb).max() <= 2.0 ** -9 * scale   # sinf/cosf/powf: libm vs libdevice, q0*c - q1*s can cancel


def test_oracle_attention():
    g, o = load("attention_h4_hs64_pos40.npz"), H.oracle()
    nh, hs, kv_mul, pos = 4, 64, 2, 40
    out, att = np.zeros(nh * hs, np.uint16), np.zeros(nh * (pos + 1), np.uint16)
    o.oracle_attention(H.ptr(out), H.ptr(np.ascontiguousarray(g["q"])), H.ptr(np.ascontiguousarray(g["k"])),
                       H.ptr(np.ascontiguousarray(g["v"])), H.ptr(att), nh, hs, kv_mul, pos)
    assert H.ulp_diff_f16(out, g["y"]).max() <= 2
    assert H.ulp_diff_f16(att, g["att"]).max() <= 2


def _tiny_model(lib, E, g, path):
    keys = ("dim", "hidden_dim", "n_layers", "n_heads", "n_kv_heads", "vocab_size", "seq_len")
    cfg = {k: int(v) for k, v in zip(keys, g["cfg"])}
    cfg["rope_theta"] = float(g["rope_theta"])
    c = E.Config(**cfg)
    assert lib.lq4_write_synth_model(path.encode(), C.byref(c), int(g["seed"])) == os.path.getsize(path)
    return cfg


def test_oracle_tiny_model_logits():
    """Whole-model restatement vs the reference's logits, teacher-forced with the reference's own tokens."""
    import llama_cu_awq_b200 as E
    lib, o = E.lib(), H.oracle()          # writing the synthetic .bin is host code: no GPU needed
    g = load("tiny_model_seed2024.npz")
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "tiny.bin")
        cfg = _tiny_model(lib, E, g, path)
        m = o.oracle_model_open(path.encode())
        assert m
        try:
            toks = [int(t) for t in g["tokens"]]
            for pos in range(g["logits"].shape[0]):
                lg = np.zeros(cfg["vocab_size"], np.uint16)
                o.oracle_model_forward(m, toks[pos], pos, H.ptr(lg), -1)
                want = g["logits"][pos]
                a, b = lg.view(np.float16).astype(np.float32), want.view(np.float16).astype(np.float32)
                ok = (H.ulp_diff_f16(lg, want) <= 4) | (np.abs(a - b) < 2e-3 * np.maximum(np.abs(b), 1.0))
                assert ok.all(), f"pos {pos}: {np.count_nonzero(~ok)} logits off"
                if pos + 1 >= len(g["prompt"]) and o.oracle_argmax_ties(H.ptr(np.ascontiguousarray(want)), cfg["vocab_size"]) == 1:
                    assert o.oracle_argmax(H.ptr(np.ascontiguousarray(want)), cfg["vocab_size"]) == toks[pos + 1]
        finally:
            o.oracle_model_close(m)


# ------------------------------------------------------------------ engine vs golden (GPU, bit-exact)
@pytest.fixture(scope="module")
def eng():
    import llama_cu_awq_b200 as E
    lib = E.lib()
    assert lib.lq4_init(0) == 0
    return E, lib


def _sync(lib):
    assert lib.lq4_stream_synchronize() == 0, lib.lq4_last_error()


@pytest.mark.gpu
@pytest.mark.parametrize("name,K,N", [("gemv_q4_K256_N64.npz", 256, 64), ("gemv_q4_K1024_N32.npz", 1024, 32), ("gemv_q4_K1408_N16.npz", 1408, 16)])
def test_engine_gemv_q4(eng, name, K, N):
    E, lib = eng
    g = load(name)
    tw, tz, ts, tx = H.to_dev(g["w"]), H.to_dev(g["z"]), H.to_dev(g["s"]), H.to_dev(g["x"])
    q = E.QWeight(tw.data_ptr(), tz.data_ptr(), ts.data_ptr())
    for accum, key in ((0, "y"), (1, "y_accum")):
        out = H.to_dev(g["out0"])
        lib.lq4_matmul_q4(out.data_ptr(), tx.data_ptr(), C.byref(q), K, N, accum, -1, None)
        _sync(lib)
        assert (H.dev_u16(out) == g[key]).all()


@pytest.mark.gpu
def test_engine_ffn_silu(eng):
    E, lib = eng
    g = load("ffn_silu_K256_N48.npz")
    keep = [H.to_dev(g[k]) for k in ("gw", "gz", "gs", "uw", "uz", "us", "x")]
    gq = E.QWeight(*(t.data_ptr() for t in keep[0:3]))
    uq = E.QWeight(*(t.data_ptr() for t in keep[3:6]))
    out = H.to_dev(np.zeros(48, np.uint16))
    lib.lq4_ffn_matvec_silu(out.data_ptr(), keep[6].data_ptr(), C.byref(gq), C.byref(uq), 256, 48)
    _sync(lib)
    assert (H.dev_u16(out) == g["y"]).all()


@pytest.mark.gpu
def test_engine_gemv_f16_rmsnorm_rope_attention(eng):
    import torch
    E, lib = eng
    g = load("gemv_f16_n512_d40.npz")
    w, x, out = H.to_dev(g["w"]), H.to_dev(g["x"]), H.to_dev(np.zeros(40, np.uint16))
    lib.lq4_matmul_fp16(out.data_ptr(), x.data_ptr(), w.data_ptr(), 512, 40, 1, 0, 0, 0, -1, 1.0)
    _sync(lib)
    assert (H.dev_u16(out) == g["y"]).all()

    g = load("rmsnorm_1000.npz")
    x, w, out = H.to_dev(g["x"]), H.to_dev(g["w"]), H.to_dev(np.zeros(1000, np.uint16))
    lib.lq4_rmsnorm(out.data_ptr(), x.data_ptr(), w.data_ptr(), 1000)
    _sync(lib)
    assert (H.dev_u16(out) == g["y"]).all()

    g = load("rope_h4_kv2_hs64_pos37.npz")
    nh, nkv, hs, pos = 4, 2, 64, 37
    cache = np.zeros((pos + 1) * nkv * hs, np.uint16)
    cache[pos * nkv * hs:] = g["k"]
    q, k = H.to_dev(g["q"]), H.to_dev(cache)
    dp = torch.tensor([pos], dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    lib.lq4_rope_rotation(q.data_ptr(), k.data_ptr(), nh, nkv, hs, dp.data_ptr(), 0, 10000.0)
    _sync(lib)
    assert (H.dev_u16(q) == g["q_out"]).all() and (H.dev_u16(k)[pos * nkv * hs:] == g["k_out"]).all()

    g = load("attention_h4_hs64_pos40.npz")
    nh, hs, kv_mul, pos, seq = 4, 64, 2, 40, 64
    q, k, v = H.to_dev(g["q"]), H.to_dev(g["k"]), H.to_dev(g["v"])
    att, out = H.to_dev(np.zeros(nh * seq, np.uint16)), H.to_dev(np.zeros(nh * hs, np.uint16))
    dp = torch.tensor([pos], dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    lib.lq4_multi_head_attention(out.data_ptr(), q.data_ptr(), k.data_ptr(), v.data_ptr(), att.data_ptr(), nh, hs, kv_mul, seq, dp.data_ptr())
    _sync(lib)
    assert (H.dev_u16(out) == g["y"]).all()
    assert (H.dev_u16(att)[:nh * (pos + 1)] == g["att"]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [1, 0])
def test_engine_tiny_model(eng, fused):
    """Bit-identical logits at every step (teacher-forced with the reference's tokens) and identical greedy ids."""
    E, lib = eng
    g = load("tiny_model_seed2024.npz")
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "tiny.bin")
        cfg = _tiny_model(lib, E, g, path)
        t = E.Transformer()
        assert lib.lq4_build_transformer(C.byref(t), path.encode(), 0) == 0
        s = E.Sampler()
        lib.lq4_build_sampler(C.byref(s), cfg["vocab_size"], 0.0, 0.9, 1)
        lib.lq4_set_option(b"fused", fused)
        try:
            steps = g["logits"].shape[0]
            toks = np.ascontiguousarray(g["tokens"][: steps + 1].astype(np.int32))
            lib.lq4_reset(C.byref(t), toks.ctypes.data_as(C.POINTER(C.c_int)), len(toks))
            nxt = C.c_int(0)
            for pos in range(steps):
                lg = np.zeros(cfg["vocab_size"], np.uint16)
                lib.lq4_step(C.byref(t), C.byref(s), 0, lg.ctypes.data_as(C.c_void_p), C.byref(nxt))
                assert (lg == g["logits"][pos]).all(), f"fused={fused}: logits differ at step {pos}"
            prompt = np.ascontiguousarray(g["prompt"].astype(np.int32))
            out = (C.c_int * (steps + 1))()
            secs = C.c_double(0)
            n = lib.lq4_generate_tokens(C.byref(t), C.byref(s), prompt.ctypes.data_as(C.POINTER(C.c_int)), len(prompt), steps + 1, out, C.byref(secs), 1)
            o = H.oracle()
            for i in range(1, n):
                if out[i] != int(g["tokens"][i]):
                    assert o.oracle_argmax_ties(H.ptr(np.ascontiguousarray(g["logits"][i - 1])), cfg["vocab_size"]) > 1, f"token {i} differs without a tie"
                    break
        finally:
            lib.lq4_set_option(b"fused", 1)
            lib.lq4_free_transformer(C.byref(t))

"""Scope row f1 (BASELINE.json configs[4]): batched prefill on the tcgen05 tensor cores.

The reference has no prefill (prompt tokens go one by one through decode, llama2_q4.cu:465-470), so the checker is the
sequential decode path itself -- which IS bit-identical to the reference (test_gpu_e2e.py) -- on the same tokens.  The bar is
fp16 tolerance, written below, not bit-exactness: the tensor cores sum in their own order and the weights are rounded to fp16
once ((q - z) * s has up to 15 significant bits)."""
import ctypes as C
import os
import sys
import tempfile

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import llama_cu_awq_b200 as E
    lib = E.lib()
    assert lib.lq4_init(0) == 0
    return E, lib


def sync(lib):
    assert lib.lq4_stream_synchronize() == 0, lib.lq4_last_error()


@pytest.mark.parametrize("M,K,N,accum", [(96, 256, 256, 0), (300, 1024, 384, 1), (256, 4096, 4096, 0), (1000, 4096, 11008, 0),
                                         (512, 11008, 4096, 1), (2048, 5120, 5120, 0), (257, 13824, 5120, 0)])
def test_gemm_q4_tensor_core(eng, M, K, N, accum):
    """y = x . dequant(W)^T on the tensor cores against a float64 product of the same fp16 inputs.  Tolerance: each weight
    is rounded to fp16 (relative 2^-11), products accumulate in fp32, the result is rounded to fp16 (relative 2^-11): the
    error of an output is bounded by a few 2^-11 of sum |x_k w_k|; 2^-9 of that sum (plus one fp16 ulp of the output) is
    the bar."""
    import torch
    E, lib = eng
    rng = np.random.default_rng(M + K + N)
    w, z, s = H.random_qweight(rng, K, N)
    x = rng.standard_normal((M, K)).astype(np.float16)
    y0 = rng.standard_normal((M, N)).astype(np.float16)
    wd = H.dequant_f64(w, z, s, K)                     # [N][K]
    xf = x.astype(np.float64)
    want = xf @ wd.T + (y0.astype(np.float64) if accum else 0.0)
    bound = np.abs(xf) @ np.abs(wd).T + (np.abs(y0.astype(np.float64)) if accum else 0.0)
    dx = H.to_dev(x.view(np.uint16))
    dy = H.to_dev(y0.view(np.uint16).copy())
    dw, dz, ds = H.to_dev(w), H.to_dev(z), H.to_dev(s.view(np.uint16))
    q = E.QWeight(dw.data_ptr(), dz.data_ptr(), ds.data_ptr())
    assert lib.lq4_gemm_q4(dy.data_ptr(), dx.data_ptr(), C.byref(q), M, K, N, accum) == 0
    sync(lib)
    got = H.dev_u16(dy).view(np.float16).astype(np.float64).reshape(M, N)
    assert np.isfinite(got).all()
    err = np.abs(got - want)
    tol = 2.0 ** -9 * bound + 2.0 ** -10 * np.abs(want) + 1e-4
    bad = err > tol
    assert not bad.any(), f"{bad.sum()} of {M * N} outputs off; worst {err.max():.4g} (tolerance there {tol.flat[err.argmax()]:.4g})"


def test_gemm_q4_rejects_unsupported_shapes(eng):
    import torch
    E, lib = eng
    t = torch.zeros(1 << 16, dtype=torch.int32, device="cuda")
    q = E.QWeight(t.data_ptr(), t.data_ptr(), t.data_ptr())
    assert lib.lq4_gemm_q4(t.data_ptr(), t.data_ptr(), C.byref(q), 64, 96, 128, 0) == 1      # K % 64
    assert lib.lq4_gemm_q4(t.data_ptr(), t.data_ptr(), C.byref(q), 64, 128, 64, 0) == 1      # N % 128


def _decode_logits(E, lib, t, s, tokens, vocab):
    """Sequential decode of one sequence, teacher-forced; returns the fp16 logits bits after the last token."""
    toks = (C.c_int * len(tokens))(*tokens)
    lib.lq4_reset(C.byref(t), toks, len(tokens))
    lg = np.zeros(vocab, np.uint16)
    nxt = C.c_int(0)
    for i in range(len(tokens)):
        lib.lq4_step(C.byref(t), C.byref(s), 0, lg.ctypes.data_as(C.c_void_p) if i == len(tokens) - 1 else None, C.byref(nxt))
    return lg


def _check_prefill(E, lib, cfg, path, batch, seq, seed):
    t = E.Transformer()
    assert lib.lq4_build_transformer(C.byref(t), path.encode(), 0) == 0
    s = E.Sampler()
    lib.lq4_build_sampler(C.byref(s), cfg["vocab_size"], 0.0, 0.9, 1)
    try:
        rng = np.random.default_rng(seed)
        vocab = cfg["vocab_size"]
        toks = rng.integers(3, vocab, size=(batch, seq)).astype(np.int32)
        toks[:, 0] = 1
        out = np.zeros((batch, vocab), np.uint16)
        ms, msg = C.c_float(0), C.c_float(0)
        rc = lib.lq4_prefill(C.byref(t), toks.ctypes.data_as(C.POINTER(C.c_int)), batch, seq, 0, out.ctypes.data_as(C.c_void_p), C.byref(ms), C.byref(msg))
        assert rc == 0
        # (a) hand-over to decode: the next decode step recomputes the last prompt position of sequence 0 over the prefilled cache
        lg = np.zeros(vocab, np.uint16)
        nxt = C.c_int(0)
        lib.lq4_step(C.byref(t), C.byref(s), 1, lg.ctypes.data_as(C.c_void_p), C.byref(nxt))
        cont = lg.view(np.float16).astype(np.float32)
        # (b) every sequence's last-position logits against the sequential decode path
        for b in range(batch):
            want = _decode_logits(E, lib, t, s, [int(v) for v in toks[b]], vocab).view(np.float16).astype(np.float32)
            got = out[b].view(np.float16).astype(np.float32)
            assert np.isfinite(got).all()
            scale = max(1.0, float(np.abs(want).max()))
            # fp16 tolerance: activations are fp16 between every op (relative 2^-11 each), a handful of ops per layer
            tol = 0.02 * scale
            assert np.abs(got - want).max() <= tol, f"sequence {b}: max |prefill - decode| = {np.abs(got - want).max():.4g} > {tol:.4g}"
            cos = float(np.dot(got, want) / (np.linalg.norm(got) * np.linalg.norm(want) + 1e-30))
            assert cos > 0.9995, f"sequence {b}: cosine {cos}"
            if b == 0:
                assert np.abs(cont - want).max() <= tol, "decode over the prefilled KV cache disagrees with pure decode"
        return ms.value, msg.value
    finally:
        lib.lq4_free_transformer(C.byref(t))


@pytest.mark.parametrize("cfg_name,batch,seq", [("TINY", 3, 40), ("TINY_PF", 2, 33), ("SMALL", 2, 96), ("SMALL", 8, 64)])
def test_prefill_matches_sequential_decode(eng, cfg_name, batch, seq):
    E, lib = eng
    # TINY_PF: grouped-query attention with sizes the tensor-core GEMM takes (kv_dim a multiple of 128)
    cfg = dict(dim=512, hidden_dim=768, n_layers=2, n_heads=8, n_kv_heads=2, vocab_size=320, seq_len=160, rope_theta=1000000.0) if cfg_name == "TINY_PF" else getattr(H, cfg_name)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "m.bin")
        c = E.Config(**cfg)
        assert lib.lq4_write_synth_model(path.encode(), C.byref(c), 4242) == os.path.getsize(path)
        _check_prefill(E, lib, cfg, path, batch, seq, batch * 1000 + seq)


def test_prefill_7b_matches_sequential_decode(eng):
    """The 7B model itself: one sequence of 96 tokens and a batch of 2 x 64 (BASELINE's configs[4] is batch 8 x seq 2048, which
    bench.py times; the checker -- sequential decode of every sequence -- is what bounds the test size)."""
    E, lib = eng
    sys.path.insert(0, H.ROOT)
    import bench as B
    cfg = B.model_cfg("7b")
    path, _ = B.ensure_files(lib, E, "7b", cfg)
    _check_prefill(E, lib, cfg, path, 2, 64, 7)

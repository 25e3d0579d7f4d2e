"""GPU parity of every operator of the C ABI (include/llama_q4_b200.h) against
  (a) the CPU oracle (oracle/cpu_ref.c)            -- bit-exact where no transcendental is involved,
                                                      <= 1 fp16 ulp for SiLU / softmax / RoPE;
  (b) the UNMODIFIED reference kernels (oracle/_ref/libq4ref.so) on the same device buffers
                                                   -- bit-exact, always.
All calls go through the C ABI with raw device pointers; torch only owns the memory."""
import ctypes as C

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=[0, 1], ids=["interp", "opk"])
def eng(request):
    """Every operator test runs twice: through one-op launches of the persistent kernel (the default) and, with option opk=1,
    through the stand-alone kernels chained by programmatic dependent launch; both must give the reference's bits."""
    import llama_cu_awq_b200 as E
    lib = E.lib()
    assert lib.lq4_init(0) == 0
    lib.lq4_set_option(b"opk", request.param)
    yield E, lib
    lib.lq4_set_option(b"opk", 0)


def sync(lib):
    assert lib.lq4_stream_synchronize() == 0, lib.lq4_last_error()


def dev_qw(E, w, z, s):
    tw, tz, ts = H.to_dev(w), H.to_dev(z), H.to_dev(s.view(np.uint16))
    q = E.QWeight(tw.data_ptr(), tz.data_ptr(), ts.data_ptr())
    return q, (tw, tz, ts)


def dev_pos(pos):
    import torch
    t = torch.tensor([pos], dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    return t


SHAPES = [(128, 8), (256, 64), (384, 16), (1024, 24), (4096, 4096), (11008, 256), (5120, 64), (13824, 32)]


@pytest.mark.parametrize("K,N", SHAPES)
@pytest.mark.parametrize("accum", [0, 1])
def test_matmul_q4(eng, K, N, accum):
    """config 0 of BASELINE.json: single INT4 GEMV (K=4096,N=4096 among others) vs host-CPU dequant."""
    import torch
    E, lib = eng
    o = H.oracle()
    rng = np.random.default_rng(K * 31 + N + accum)
    w, z, s = H.random_qweight(rng, K, N)
    x = rng.standard_normal(K).astype(np.float16)
    out0 = rng.standard_normal(N).astype(np.float16)
    want = out0.view(np.uint16).copy()
    o.oracle_matvec_int4(H.ptr(want), H.ptr(x.view(np.uint16)), H.oracle_qw(w, z, s), K, N, accum)
    q, keep = dev_qw(E, w, z, s)
    dx = H.to_dev(x.view(np.uint16))
    dout = H.to_dev(out0.view(np.uint16))
    lib.lq4_matmul_q4(dout.data_ptr(), dx.data_ptr(), C.byref(q), K, N, accum, -1, None)
    sync(lib)
    got = H.dev_u16(dout)
    assert (got == want).all(), f"{(got != want).sum()} of {N} outputs differ from the oracle"
    r = H.ref()
    if r is not None:
        dref = H.to_dev(out0.view(np.uint16))
        r.ref_matmul_q4(dref.data_ptr(), dx.data_ptr(), keep[0].data_ptr(), keep[1].data_ptr(), keep[2].data_ptr(), K, N,
                        accum, -1, None)
        torch.cuda.synchronize()
        assert (H.dev_u16(dref) == got).all()


def test_matmul_q4_cache_offset(eng):
    """loff/pPos addressing of the KV-cache row (gpu_kernels.h:225-227)."""
    E, lib = eng
    o = H.oracle()
    rng = np.random.default_rng(99)
    K, N, seq, pos, layer = 256, 64, 16, 5, 1
    w, z, s = H.random_qweight(rng, K, N)
    x = rng.standard_normal(K).astype(np.float16)
    want = np.zeros(N, np.uint16)
    o.oracle_matvec_int4(H.ptr(want), H.ptr(x.view(np.uint16)), H.oracle_qw(w, z, s), K, N, 0)
    q, keep = dev_qw(E, w, z, s)
    cache = H.to_dev(np.zeros(2 * seq * N, np.uint16))
    dp = dev_pos(pos)
    loff = layer * seq * N
    lib.lq4_matmul_q4(cache.data_ptr(), H.to_dev(x.view(np.uint16)).data_ptr(), C.byref(q), K, N, 0, loff, dp.data_ptr())
    sync(lib)
    got = H.dev_u16(cache)
    assert (got[loff + pos * N: loff + (pos + 1) * N] == want).all()
    got[loff + pos * N: loff + (pos + 1) * N] = 0
    assert not got.any()


@pytest.mark.parametrize("K,N", [(256, 64), (4096, 4096), (5120, 128)])
def test_qkv_matvec(eng, K, N):
    import torch
    E, lib = eng
    o = H.oracle()
    rng = np.random.default_rng(K + 3 * N)
    mats = [H.random_qweight(rng, K, N) for _ in range(3)]
    x = rng.standard_normal(K).astype(np.float16)
    seq, pos, loff = 8, 3, 8 * N
    wants = []
    for (w, z, s) in mats:
        out = np.zeros(N, np.uint16)
        o.oracle_matvec_int4(H.ptr(out), H.ptr(x.view(np.uint16)), H.oracle_qw(w, z, s), K, N, 0)
        wants.append(out)
    qs = [dev_qw(E, *m) for m in mats]
    dx = H.to_dev(x.view(np.uint16))
    dq = H.to_dev(np.zeros(N, np.uint16))
    kc = H.to_dev(np.zeros(2 * seq * N, np.uint16))
    vc = H.to_dev(np.zeros(2 * seq * N, np.uint16))
    dp = dev_pos(pos)
    lib.lq4_qkv_matvec(dq.data_ptr(), kc.data_ptr(), vc.data_ptr(), dx.data_ptr(), C.byref(qs[0][0]), C.byref(qs[1][0]),
                       C.byref(qs[2][0]), K, N, loff, dp.data_ptr())
    sync(lib)
    row = slice(loff + pos * N, loff + (pos + 1) * N)
    assert (H.dev_u16(dq) == wants[0]).all()
    assert (H.dev_u16(kc)[row] == wants[1]).all()
    assert (H.dev_u16(vc)[row] == wants[2]).all()
    r = H.ref()
    if r is not None:
        rq = H.to_dev(np.zeros(N, np.uint16)); rk = H.to_dev(np.zeros(2 * seq * N, np.uint16)); rv = H.to_dev(np.zeros(2 * seq * N, np.uint16))
        args = [rq.data_ptr(), rk.data_ptr(), rv.data_ptr(), dx.data_ptr()]
        for (_, keep) in qs:
            args += [t.data_ptr() for t in keep]
        r.ref_qkv_matvec(*args, K, N, loff, dp.data_ptr())
        torch.cuda.synchronize()
        assert (H.dev_u16(rq) == H.dev_u16(dq)).all() and (H.dev_u16(rk) == H.dev_u16(kc)).all() and (H.dev_u16(rv) == H.dev_u16(vc)).all()


@pytest.mark.parametrize("K,N", [(256, 64), (4096, 11008), (5120, 256)])
def test_ffn_matvec_silu(eng, K, N):
    import torch
    E, lib = eng
    o = H.oracle()
    rng = np.random.default_rng(K + 5 * N)
    g = H.random_qweight(rng, K, N)
    u = H.random_qweight(rng, K, N)
    x = rng.standard_normal(K).astype(np.float16)
    want = np.zeros(N, np.uint16)
    o.oracle_ffn_matvec_silu(H.ptr(want), H.ptr(x.view(np.uint16)), H.oracle_qw(*g), H.oracle_qw(*u), K, N)
    gq, gk = dev_qw(E, *g)
    uq, uk = dev_qw(E, *u)
    dx = H.to_dev(x.view(np.uint16))
    dout = H.to_dev(np.zeros(N, np.uint16))
    lib.lq4_ffn_matvec_silu(dout.data_ptr(), dx.data_ptr(), C.byref(gq), C.byref(uq), K, N)
    sync(lib)
    got = H.dev_u16(dout)
    # SiLU uses expf: host libm vs CUDA libdevice may differ by an fp32 ulp -> <= 1 fp16 ulp here
    assert H.ulp_diff_f16(got, want).max() <= 1
    r = H.ref()
    if r is not None:
        dref = H.to_dev(np.zeros(N, np.uint16))
        r.ref_ffn_matvec_silu(dref.data_ptr(), dx.data_ptr(), *[t.data_ptr() for t in gk], *[t.data_ptr() for t in uk], K, N)
        torch.cuda.synchronize()
        assert (H.dev_u16(dref) == got).all(), "SiLU path must be bit-identical to the reference CUDA build"


@pytest.mark.parametrize("n,d", [(256, 512), (4096, 32000), (5120, 1000), (1024, 20)])
def test_matmul_fp16(eng, n, d):
    import torch
    E, lib = eng
    o = H.oracle()
    if d % 8:
        d = d // 8 * 8
    rng = np.random.default_rng(n + d)
    w = (rng.standard_normal((d, n)) * 0.02).astype(np.float16)
    x = rng.standard_normal(n).astype(np.float16)
    want = np.zeros(d, np.uint16)
    o.oracle_matvec_fp16(H.ptr(want), H.ptr(x.view(np.uint16)), H.ptr(w.view(np.uint16)), n, d, 1.0)
    dw, dx, dout = H.to_dev(w.view(np.uint16)), H.to_dev(x.view(np.uint16)), H.to_dev(np.zeros(d, np.uint16))
    lib.lq4_matmul_fp16(dout.data_ptr(), dx.data_ptr(), dw.data_ptr(), n, d, 1, 0, 0, 0, -1, 1.0)
    sync(lib)
    got = H.dev_u16(dout)
    assert (got == want).all()
    r = H.ref()
    if r is not None:
        dref = H.to_dev(np.zeros(d, np.uint16))
        r.ref_matmul_fp16(dref.data_ptr(), dx.data_ptr(), dw.data_ptr(), n, d)
        torch.cuda.synchronize()
        assert (H.dev_u16(dref) == got).all()


@pytest.mark.parametrize("size", [256, 1000, 4096, 5120, 8192])
def test_rmsnorm(eng, size):
    import torch
    E, lib = eng
    o = H.oracle()
    rng = np.random.default_rng(size)
    x = (rng.standard_normal(size) * 4).astype(np.float16)
    w = (0.9 + 0.2 * rng.random(size)).astype(np.float16)
    want = np.zeros(size, np.uint16)
    o.oracle_rmsnorm(H.ptr(want), H.ptr(x.view(np.uint16)), H.ptr(w.view(np.uint16)), size)
    dx, dw, dout = H.to_dev(x.view(np.uint16)), H.to_dev(w.view(np.uint16)), H.to_dev(np.zeros(size, np.uint16))
    lib.lq4_rmsnorm(dout.data_ptr(), dx.data_ptr(), dw.data_ptr(), size)
    sync(lib)
    got = H.dev_u16(dout)
    assert (got == want).all()
    r = H.ref()
    if r is not None:
        dref = H.to_dev(np.zeros(size, np.uint16))
        r.ref_rmsnorm(dref.data_ptr(), dx.data_ptr(), dw.data_ptr(), size)
        torch.cuda.synchronize()
        assert (H.dev_u16(dref) == got).all()


@pytest.mark.parametrize("nh,nkv,hs,theta", [(4, 4, 64, 10000.0), (32, 32, 128, 10000.0), (8, 2, 32, 1000000.0)])
def test_rope_rotation(eng, nh, nkv, hs, theta):
    import torch
    E, lib = eng
    o = H.oracle()
    r = H.ref()
    rng = np.random.default_rng(nh * hs)
    seq = 2048
    for pos in (0, 1, 7, 255, 1000, 2047):
        q = rng.standard_normal(nh * hs).astype(np.float16)
        kc = rng.standard_normal(seq * nkv * hs).astype(np.float16)
        qo = q.view(np.uint16).copy()
        ko = kc.view(np.uint16)[pos * nkv * hs:(pos + 1) * nkv * hs].copy()
        o.oracle_rope(H.ptr(qo), H.ptr(ko), nh, nkv, hs, pos, theta)
        dq, dk, dp = H.to_dev(q.view(np.uint16)), H.to_dev(kc.view(np.uint16)), dev_pos(pos)
        lib.lq4_rope_rotation(dq.data_ptr(), dk.data_ptr(), nh, nkv, hs, dp.data_ptr(), 0, theta)
        sync(lib)
        gq, gk = H.dev_u16(dq), H.dev_u16(dk)
        # vs the CPU oracle: host libm vs CUDA sinf/cosf differ by an fp32 ulp or so, and q0*c - q1*s can cancel, so
        # the tolerance is absolute (2 fp16 ulps of the largest input); the bit-exact check is the reference below
        def close(a, b, scale):
            fa, fb = a.view(np.float16).astype(np.float32), b.view(np.float16).astype(np.float32)
            return np.abs(fa - fb).max() <= 2.0 ** -9 * scale
        assert close(gq, qo, float(np.abs(q.astype(np.float32)).max()))
        assert close(gk[pos * nkv * hs:(pos + 1) * nkv * hs], ko, float(np.abs(kc.astype(np.float32)).max()))
        if r is not None:
            rq, rk = H.to_dev(q.view(np.uint16)), H.to_dev(kc.view(np.uint16))
            r.ref_rope(rq.data_ptr(), rk.data_ptr(), nh, nkv, hs, dp.data_ptr(), 0, theta)
            torch.cuda.synchronize()
            assert (H.dev_u16(rq) == gq).all() and (H.dev_u16(rk) == gk).all(), f"pos {pos}"


@pytest.mark.parametrize("nh,hs,kv_mul", [(4, 64, 1), (32, 128, 1), (8, 32, 4), (40, 128, 1)])
@pytest.mark.parametrize("pos", [0, 1, 31, 32, 33, 255, 383, 384, 700, 1000])
def test_multi_head_attention(eng, nh, hs, kv_mul, pos):
    """Positions from 384 on run several CTAs per head (interp_sm100.cuh, run_attn_split): 4 parts for (32, 128), 2 parts for
    (4, 64) and -- 40 heads x 4 exceed the SM count -- for (40, 128), the 13B shape; (8, 32) stays on one CTA per head."""
    import torch
    E, lib = eng
    o = H.oracle()
    r = H.ref()
    rng = np.random.default_rng(nh + hs + pos)
    seq = 1024
    kv_dim = nh * hs // kv_mul
    q = rng.standard_normal(nh * hs).astype(np.float16)
    kc = rng.standard_normal((seq, kv_dim)).astype(np.float16)
    vc = rng.standard_normal((seq, kv_dim)).astype(np.float16)
    want_att = np.zeros(nh * (pos + 1), np.uint16)
    want = np.zeros(nh * hs, np.uint16)
    o.oracle_attention(H.ptr(want), H.ptr(q.view(np.uint16)), H.ptr(kc.view(np.uint16)), H.ptr(vc.view(np.uint16)),
                       H.ptr(want_att), nh, hs, kv_mul, pos)
    dq, dk, dv = H.to_dev(q.view(np.uint16)), H.to_dev(kc.view(np.uint16)), H.to_dev(vc.view(np.uint16))
    datt = H.to_dev(np.zeros(nh * seq, np.uint16))
    dout = H.to_dev(np.zeros(nh * hs, np.uint16))
    dp = dev_pos(pos)
    lib.lq4_multi_head_attention(dout.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), datt.data_ptr(), nh, hs,
                                 kv_mul, seq, dp.data_ptr())
    sync(lib)
    got = H.dev_u16(dout)
    # vs the CPU oracle: host libm vs CUDA expf differ in the last bit of a probability, and an output that is a sum of ~pos
    # terms of mixed sign can sit near zero, where an fp16 ulp is tiny: 2 ulps, or 2^-9 of the largest output.  The bit-exact
    # comparison is the one against the reference kernels below.
    gf, wf = got.view(np.float16).astype(np.float32), want.view(np.float16).astype(np.float32)
    assert ((H.ulp_diff_f16(got, want) <= 2) | (np.abs(gf - wf) <= 2.0 ** -9 * np.abs(wf).max())).all()
    if r is not None:
        ratt = H.to_dev(np.zeros(nh * seq, np.uint16))
        rout = H.to_dev(np.zeros(nh * hs, np.uint16))
        r.ref_mha(rout.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), ratt.data_ptr(), nh, hs, kv_mul, seq,
                  dp.data_ptr())
        torch.cuda.synchronize()
        assert (H.dev_u16(rout) == got).all(), "attention output must be bit-identical to the reference kernels"
        n = nh * (pos + 1)
        assert (H.dev_u16(ratt)[:n] == H.dev_u16(datt)[:n]).all(), "softmax probabilities must be bit-identical"


@pytest.mark.parametrize("nh,hs,kv_mul", [(4, 64, 1), (8, 128, 2), (4, 96, 1)])
@pytest.mark.parametrize("pos", [100, 1500, 8191, 8192, 9000, 12287])
def test_multi_head_attention_long_context(eng, nh, hs, kv_mul, pos):
    """Scope row f2: sequence bins above 8192.  MultiHeadAttention then runs softmax_kernel_no_smem (llama2_q4.cu:276-279,
    gpu_kernels.h:403-446), which keeps exp() in the fp16 score buffer: the sum adds the unrounded values while the quotient is
    formed from the rounded one.  Bit-exact against the reference kernels (output and probabilities); hs = 96 takes the
    stand-alone attention kernel, the others the persistent kernel's attention op."""
    import torch
    E, lib = eng
    r = H.ref()
    rng = np.random.default_rng(nh * 1000 + hs + pos)
    seq = 12288
    kv_dim = nh * hs // kv_mul
    q = rng.standard_normal(nh * hs).astype(np.float16)
    kc = (0.5 * rng.standard_normal((seq, kv_dim))).astype(np.float16)
    vc = rng.standard_normal((seq, kv_dim)).astype(np.float16)
    dq, dk, dv = H.to_dev(q.view(np.uint16)), H.to_dev(kc.view(np.uint16)), H.to_dev(vc.view(np.uint16))
    datt = H.to_dev(np.zeros(nh * seq, np.uint16))
    dout = H.to_dev(np.zeros(nh * hs, np.uint16))
    dp = dev_pos(pos)
    lib.lq4_multi_head_attention(dout.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), datt.data_ptr(), nh, hs,
                                 kv_mul, seq, dp.data_ptr())
    sync(lib)
    got = H.dev_u16(dout)
    ratt = H.to_dev(np.zeros(nh * seq, np.uint16))
    rout = H.to_dev(np.zeros(nh * hs, np.uint16))
    r.ref_mha(rout.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), ratt.data_ptr(), nh, hs, kv_mul, seq, dp.data_ptr())
    torch.cuda.synchronize()
    assert (H.dev_u16(rout) == got).all(), "attention output must be bit-identical to the reference kernels"
    n = nh * (pos + 1)
    assert (H.dev_u16(ratt)[:n] == H.dev_u16(datt)[:n]).all(), "softmax probabilities must be bit-identical"

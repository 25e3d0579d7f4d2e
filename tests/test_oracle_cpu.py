"""CPU-only checks of the oracle (oracle/cpu_ref.c): fp16 conversions against numpy, every restated
kernel against an independent float64 computation, and the summation-order properties that make it
a bit-exact restatement (the golden-vector pin against the real reference is test_golden.py)."""
import numpy as np
import pytest

import helpers as H


def test_fp16_roundtrip_all_bit_patterns():
    o = H.oracle()
    bits = np.arange(0, 65536, dtype=np.uint16)
    ref = bits.view(np.float16).astype(np.float32)
    for b in range(0, 65536, 97):   # sample; the full sweep is below in vectorised form
        f = o.oracle_h2f(int(bits[b]))
        if np.isnan(ref[b]):
            assert np.isnan(f)
        else:
            assert f == ref[b]
            assert o.oracle_f2h(float(ref[b])) == bits[b]


def test_f2h_rounding_matches_numpy():
    o = H.oracle()
    rng = np.random.default_rng(1)
    vals = np.concatenate([
        rng.standard_normal(4000).astype(np.float32) * 10.0 ** rng.integers(-9, 5, 4000),
        np.array([0.0, -0.0, 65504.0, 65519.9, 65520.0, 1e9, -1e9, 2.0**-24, 2.0**-25, 2.0**-25 * 1.0001, 2.0**-14,
                  (2.0**-14) * (1 - 2.0**-11), 1.0 + 2.0**-11, 1.0 + 2.0**-11 + 2.0**-20, 1.0 + 3 * 2.0**-11], np.float32)])
    with np.errstate(over="ignore"):
        want = vals.astype(np.float16).view(np.uint16)
    got = np.array([o.oracle_f2h(float(v)) for v in vals], dtype=np.uint16)
    assert (got == want).all()


@pytest.mark.parametrize("K,N", [(128, 8), (256, 16), (384, 8), (1024, 8), (4096, 16), (11008, 8)])
def test_matvec_int4_close_to_float64(K, N):
    o = H.oracle()
    rng = np.random.default_rng(K + N)
    w, z, s = H.random_qweight(rng, K, N)
    x = rng.standard_normal(K).astype(np.float16)
    out = np.zeros(N, np.uint16)
    qw = H.oracle_qw(w, z, s)
    o.oracle_matvec_int4(H.ptr(out), H.ptr(x.view(np.uint16)), qw, K, N, 0)
    want = H.dequant_f64(w, z, s, K) @ x.astype(np.float64)
    got = out.view(np.float16).astype(np.float64)
    assert np.allclose(got, want, rtol=2e-3, atol=2e-3)
    # accumulate variant: out += (gpu_kernels.h:229-230)
    acc0 = rng.standard_normal(N).astype(np.float16)
    out2 = acc0.view(np.uint16).copy()
    o.oracle_matvec_int4(H.ptr(out2), H.ptr(x.view(np.uint16)), qw, K, N, 1)
    for n in range(N):
        f = o.oracle_dot_int4(n, H.ptr(x.view(np.uint16)), qw, K)
        assert out2[n] == np.float32(np.float32(f) + np.float32(acc0[n])).astype(np.float16).view(np.uint16)


def test_dot_int4_lane_order_is_the_references():
    """Recompute one column with an independent numpy restatement of the lane/trip/qi/i order and the
    shfl-down tree (gpu_kernels.h:176-207): must agree to the bit, while a plain left-to-right fp32
    sum over k generally does not."""
    o = H.oracle()
    rng = np.random.default_rng(7)
    K = 4096
    w, z, s = H.random_qweight(rng, K, 8)
    x = rng.standard_normal(K).astype(np.float16)
    qw = H.oracle_qw(w, z, s)
    deq = (H.dequant_f64(w, z, s, K)).astype(np.float32)   # (q-z)*s is exact in fp32
    xf = x.astype(np.float32)
    differs = 0
    for n in range(8):
        lanes = np.zeros(32, np.float32)
        for L in range(32):
            acc = np.float32(0)
            for ygq in range(K // 1024):
                base = ygq * 1024 + L * 32
                for k in range(base, base + 32):
                    # fp32 fma: the product of two fp32 is exact in float64; one float64 add, then RN to fp32
                    acc = np.float32(np.float64(deq[n, k]) * np.float64(xf[k]) + np.float64(acc))
            lanes[L] = acc
        v = lanes.copy()
        for off in (1, 2, 4, 8, 16):
            t = v.copy()
            t[: 32 - off] = v[: 32 - off] + v[off:]
            v = t
        got = np.float32(o.oracle_dot_int4(n, H.ptr(x.view(np.uint16)), qw, K))
        assert got == v[0]
        seq = np.float32(0)
        for k in range(K):
            seq = np.float32(seq + deq[n, k] * xf[k])
        differs += int(seq != got)
    assert differs > 0


def test_rmsnorm_close_and_alias_safe():
    o = H.oracle()
    rng = np.random.default_rng(3)
    for size in (256, 4096, 5120):
        x = (rng.standard_normal(size) * 3).astype(np.float16)
        w = (0.9 + 0.2 * rng.random(size)).astype(np.float16)
        out = np.zeros(size, np.uint16)
        o.oracle_rmsnorm(H.ptr(out), H.ptr(x.view(np.uint16)), H.ptr(w.view(np.uint16)), size)
        xf = x.astype(np.float64)
        want = xf / np.sqrt((xf * xf).mean() + 1e-5) * w.astype(np.float64)
        assert np.allclose(out.view(np.float16).astype(np.float64), want, rtol=2e-3, atol=1e-3)
        inplace = x.view(np.uint16).copy()
        o.oracle_rmsnorm(H.ptr(inplace), H.ptr(inplace), H.ptr(w.view(np.uint16)), size)
        assert (inplace == out).all()


def test_matvec_fp16_close():
    o = H.oracle()
    rng = np.random.default_rng(4)
    n, d = 4096, 24
    w = (rng.standard_normal((d, n)) * 0.02).astype(np.float16)
    x = rng.standard_normal(n).astype(np.float16)
    out = np.zeros(d, np.uint16)
    o.oracle_matvec_fp16(H.ptr(out), H.ptr(x.view(np.uint16)), H.ptr(w.view(np.uint16)), n, d, 1.0)
    want = w.astype(np.float64) @ x.astype(np.float64)
    assert np.allclose(out.view(np.float16).astype(np.float64), want, rtol=2e-3, atol=2e-3)


def test_rope_matches_float64_rotation():
    o = H.oracle()
    rng = np.random.default_rng(5)
    nh, nkv, hs, pos, theta = 4, 2, 64, 37, 10000.0
    q = rng.standard_normal(nh * hs).astype(np.float16)
    k = rng.standard_normal(nkv * hs).astype(np.float16)
    qo, ko = q.view(np.uint16).copy(), k.view(np.uint16).copy()
    o.oracle_rope(H.ptr(qo), H.ptr(ko), nh, nkv, hs, pos, theta)
    i = np.arange(hs // 2)
    ang = pos / theta ** (2 * i / hs)
    for vec, outv, heads in ((q, qo, nh), (k, ko, nkv)):
        v = vec.astype(np.float64).reshape(heads, hs)
        want = np.concatenate([v[:, : hs // 2] * np.cos(ang) - v[:, hs // 2:] * np.sin(ang),
                               v[:, : hs // 2] * np.sin(ang) + v[:, hs // 2:] * np.cos(ang)], axis=1).reshape(-1)
        assert np.allclose(outv.view(np.float16).astype(np.float64), want, rtol=3e-3, atol=3e-3)


@pytest.mark.parametrize("pos", [0, 1, 31, 32, 100])
def test_attention_close_to_float64(pos):
    o = H.oracle()
    rng = np.random.default_rng(6 + pos)
    nh, hs, kv_mul, seq = 4, 64, 2, 128
    kv_dim = nh * hs // kv_mul
    q = rng.standard_normal(nh * hs).astype(np.float16)
    kc = rng.standard_normal((seq, kv_dim)).astype(np.float16)
    vc = rng.standard_normal((seq, kv_dim)).astype(np.float16)
    att = np.zeros(nh * (pos + 1), np.uint16)
    out = np.zeros(nh * hs, np.uint16)
    o.oracle_attention(H.ptr(out), H.ptr(q.view(np.uint16)), H.ptr(kc.view(np.uint16)), H.ptr(vc.view(np.uint16)),
                       H.ptr(att), nh, hs, kv_mul, pos)
    want = np.zeros(nh * hs)
    for h in range(nh):
        kh = kc[: pos + 1, (h // kv_mul) * hs:(h // kv_mul + 1) * hs].astype(np.float64)
        vh = vc[: pos + 1, (h // kv_mul) * hs:(h // kv_mul + 1) * hs].astype(np.float64)
        sc = kh @ q[h * hs:(h + 1) * hs].astype(np.float64) / np.sqrt(hs)
        m = max(sc.max(), 0.0)       # the reference's max(true max, 0) quirk does not change the value
        e = np.exp(sc - m)
        want[h * hs:(h + 1) * hs] = (e / e.sum()) @ vh
    assert np.allclose(out.view(np.float16).astype(np.float64), want, rtol=1e-2, atol=5e-3)
    p = att.view(np.float16).astype(np.float64).reshape(nh, pos + 1)
    assert np.allclose(p.sum(axis=1), 1.0, atol=5e-3)


def test_argmax_lowest_index_and_tie_count():
    o = H.oracle()
    logits = np.zeros(1000, np.float16)
    logits[[17, 400, 999]] = 3.5
    logits[5] = -2
    b = logits.view(np.uint16)
    assert o.oracle_argmax(H.ptr(b), 1000) == 17
    assert o.oracle_argmax_ties(H.ptr(b), 1000) == 3

"""Shared test plumbing: the CPU oracle (oracle/liboracle.so), the reference harness
(oracle/_ref/libq4ref.so, GPU box only), seeded inputs in the reference's packed layout."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libq4ref.so")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "llama2_q4_ref")
GOLDEN = os.path.join(ROOT, "tests", "golden")

u16p = C.POINTER(C.c_uint16)
u32p = C.POINTER(C.c_uint32)


class OracleQWeight(C.Structure):
    _fields_ = [("weight", C.c_void_p), ("zeros", C.c_void_p), ("scales", C.c_void_p)]


class OracleConfig(C.Structure):
    _fields_ = [("dim", C.c_int), ("hidden_dim", C.c_int), ("n_layers", C.c_int), ("n_heads", C.c_int),
                ("n_kv_heads", C.c_int), ("vocab_size", C.c_int), ("seq_len", C.c_int), ("rope_theta", C.c_float)]


_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)
        o = C.CDLL(ORACLE_SO)
        P = C.c_void_p
        o.oracle_h2f.restype = C.c_float
        o.oracle_h2f.argtypes = [C.c_uint16]
        o.oracle_f2h.restype = C.c_uint16
        o.oracle_f2h.argtypes = [C.c_float]
        o.oracle_dot_int4.restype = C.c_float
        o.oracle_dot_int4.argtypes = [C.c_int, P, C.POINTER(OracleQWeight), C.c_int]
        o.oracle_matvec_int4.argtypes = [P, P, C.POINTER(OracleQWeight), C.c_int, C.c_int, C.c_int]
        o.oracle_ffn_matvec_silu.argtypes = [P, P, C.POINTER(OracleQWeight), C.POINTER(OracleQWeight), C.c_int, C.c_int]
        o.oracle_matvec_fp16.argtypes = [P, P, P, C.c_int, C.c_int, C.c_float]
        o.oracle_rmsnorm.argtypes = [P, P, P, C.c_int]
        o.oracle_rope.argtypes = [P, P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float]
        o.oracle_attention.argtypes = [P, P, P, P, P, C.c_int, C.c_int, C.c_int, C.c_int]
        o.oracle_qk_scores.argtypes = [P, P, P, C.c_int, C.c_int, C.c_int, C.c_int]
        o.oracle_softmax.argtypes = [P, C.c_int, C.c_int]
        o.oracle_att_v.argtypes = [P, P, P, C.c_int, C.c_int, C.c_int, C.c_int]
        o.oracle_argmax.restype = C.c_int
        o.oracle_argmax.argtypes = [P, C.c_int]
        o.oracle_argmax_ties.restype = C.c_int
        o.oracle_argmax_ties.argtypes = [P, C.c_int]
        o.oracle_model_open.restype = C.c_void_p
        o.oracle_model_open.argtypes = [C.c_char_p]
        o.oracle_model_close.argtypes = [C.c_void_p]
        o.oracle_model_config.restype = C.POINTER(OracleConfig)
        o.oracle_model_config.argtypes = [C.c_void_p]
        o.oracle_model_forward.argtypes = [C.c_void_p, C.c_int, C.c_int, P, C.c_int]
        o.oracle_model_x.restype = C.c_void_p
        o.oracle_model_x.argtypes = [C.c_void_p]
        o.oracle_set_threads.argtypes = [C.c_int]
        o.oracle_get_max_threads.restype = C.c_int
        _oracle = o
    return _oracle


def ptr(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


# ------------------------------------------------------------------ packed-weight inputs
def pwh(K):
    return ((K - 1) // 32 + 1) * 4


def groups(K):
    return (K - 1) // 128 + 1


def zh(K):
    return (groups(K) - 1) // 8 + 1


def random_qweight(rng: np.random.Generator, K: int, N: int, scale_mul: float = 1.0):
    """Random QWeight(K, N) in the runtime layout (llama2_q4.cu:90-98): uniform nibbles,
    scales ~ fp16(s0 * U(0.5,1.5)), s0 = 1/(6.52 sqrt(K))."""
    w = rng.integers(0, 2**32, size=(N, pwh(K)), dtype=np.uint32)
    z = rng.integers(0, 2**32, size=(N, zh(K)), dtype=np.uint32)
    s0 = scale_mul / (6.52 * np.sqrt(K))
    s = (s0 * (0.5 + rng.random((N, groups(K))))).astype(np.float16)
    return w, z, s


def dequant_f64(w, z, s, K):
    """Plain float64 dequantisation of a packed QWeight -> [N][K] (a NON-bit-exact sanity reference)."""
    N = w.shape[0]
    nib = ((w[:, :, None] >> (4 * np.arange(8, dtype=np.uint32))[None, None, :]) & 0xF).reshape(N, -1)[:, :K]
    G = groups(K)
    zn = ((z[:, :, None] >> (4 * np.arange(8, dtype=np.uint32))[None, None, :]) & 0xF).reshape(N, -1)[:, :G]
    gidx = np.arange(K) // 128
    return (nib.astype(np.float64) - zn[:, gidx].astype(np.float64)) * s.astype(np.float64)[:, gidx]


def oracle_qw(w, z, s):
    return OracleQWeight(ptr(w), ptr(z), ptr(s.view(np.uint16)))


def f16_bits(a):
    return np.ascontiguousarray(a.astype(np.float16)).view(np.uint16)


def ulp_diff_f16(a_bits, b_bits):
    """Distance in fp16 ulps between two uint16 bit-pattern arrays (sign-magnitude -> ordered ints)."""
    def key(x):
        x = x.astype(np.int32)
        return np.where(x & 0x8000, -(x & 0x7FFF), x & 0x7FFF)
    return np.abs(key(a_bits) - key(b_bits))


# ------------------------------------------------------------------ reference harness (GPU only)
_ref = None


def ref():
    """oracle/_ref/libq4ref.so: the UNMODIFIED reference TU behind a C shim (oracle/ref_harness.cu)."""
    global _ref
    if _ref is None:
        # The GPU parity tests are reference comparisons: on a GPU box a missing reference build is a FAILURE, never a
        # silent pass (oracle/build_ref.sh builds it wherever /root/reference exists; the built files travel with the snapshot).
        assert os.path.exists(REF_SO), f"{REF_SO} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` where /root/reference exists"
        r = C.CDLL(REF_SO)
        P = C.c_void_p
        r.ref_rmsnorm.argtypes = [P, P, P, C.c_int]
        r.ref_matmul_fp16.argtypes = [P, P, P, C.c_int, C.c_int]
        r.ref_matmul_q4.argtypes = [P, P, P, P, P, C.c_int, C.c_int, C.c_int, C.c_int, P]
        r.ref_qkv_matvec.argtypes = [P] * 13 + [C.c_int, C.c_int, C.c_int, P]
        r.ref_ffn_matvec_silu.argtypes = [P] * 8 + [C.c_int, C.c_int]
        r.ref_rope.argtypes = [P, P, C.c_int, C.c_int, C.c_int, P, C.c_int, C.c_float]
        r.ref_mha.argtypes = [P, P, P, P, P, C.c_int, C.c_int, C.c_int, C.c_int, P]
        r.ref_open.restype = C.c_int
        r.ref_open.argtypes = [C.c_char_p]
        r.ref_config.argtypes = [P]
        r.ref_set_sampler.argtypes = [C.c_float, C.c_float, C.c_ulonglong]
        r.ref_reset.argtypes = [P, C.c_int]
        r.ref_step.restype = C.c_int
        r.ref_step.argtypes = [C.c_int, P, P]
        r.ref_state_ptr.restype = C.c_void_p
        r.ref_state_ptr.argtypes = [C.c_int]
        r.ref_time_steps.restype = C.c_float
        r.ref_time_steps.argtypes = [C.c_int, C.c_int]
        r.ref_token_at.restype = C.c_int
        r.ref_token_at.argtypes = [C.c_int]
        r.ref_close.argtypes = []
        r.ref_last_cuda_error.restype = C.c_int
        _ref = r
    return _ref


def require_ref_bin():
    """oracle/_ref/llama2_q4_ref (the stock reference program): required by the GPU CLI / transcript tests."""
    assert os.path.exists(REF_BIN), f"{REF_BIN} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` where /root/reference exists"
    return REF_BIN


def our_ids_are_reference_argmaxes(path, ids_path, vocab):
    """Teacher-force our id sequence through oracle/_ref/libq4ref.so; assert each generated id attains the maximum of the
    reference's fp16 logits of the step before it; return how many of those maxima were tied."""
    vals = [int(x) for x in open(ids_path).read().split()]
    n_prompt, ids = vals[0], vals[1:]
    r = ref()
    assert r.ref_open(path.encode()) == 0
    try:
        toks = np.array(ids, dtype=np.int32)
        r.ref_reset(toks.ctypes.data_as(C.c_void_p), len(toks))
        lg = np.zeros(vocab, np.uint16)
        nxt = C.c_int(0)
        tied = 0
        for step in range(len(ids) - 1):
            r.ref_step(0, lg.ctypes.data_as(C.c_void_p), C.byref(nxt))
            if step < n_prompt - 1:
                continue                               # prompt positions: the next id is given, not sampled
            f = lg.view(np.float16).astype(np.float32)
            assert f[ids[step + 1]] == f.max(), f"our id {ids[step + 1]} at position {step + 1} is not an argmax of the reference's logits"
            tied += int((f == f.max()).sum() > 1)
        return tied
    finally:
        r.ref_close()



# ------------------------------------------------------------------ torch device helpers (GPU only)
def to_dev(a: np.ndarray):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a).view(np.int16 if a.dtype == np.uint16 else
                                                      np.int32 if a.dtype == np.uint32 else a.dtype))
    t = t.cuda()
    torch.cuda.synchronize()   # the engine runs on its own non-blocking stream: finish the upload first
    return t


def dev_u16(t):
    """device tensor of fp16 bit patterns (stored as int16) -> numpy uint16"""
    return t.cpu().numpy().view(np.uint16)


TINY = dict(dim=256, hidden_dim=512, n_layers=2, n_heads=4, n_kv_heads=4, vocab_size=512, seq_len=256,
            rope_theta=10000.0)
TINY_GQA = dict(dim=256, hidden_dim=384, n_layers=2, n_heads=8, n_kv_heads=2, vocab_size=320, seq_len=160,
                rope_theta=1000000.0)
SMALL = dict(dim=1024, hidden_dim=2816, n_layers=3, n_heads=8, n_kv_heads=8, vocab_size=2048, seq_len=512,
             rope_theta=10000.0)

"""llama_cu_awq_b200 -- Python view of the sm_100a Llama-2 AWQ-INT4 decode engine.

The product is the C-ABI shared library ``libllama_q4_b200.so`` (``include/llama_q4_b200.h``); this
module is a thin ctypes binding over it so that tests and ``bench.py`` can drive exactly the calls a
C/C++ host would make.  Function names mirror the reference's host wrappers
(ankan-ban/llama_cu_awq ``llama2_q4.cu:207-432``).  PyTorch is used by callers only to own device
memory; nothing here computes anything, and there is no fallback when the library or a GPU is
missing: loading fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LQ4_LIB") or os.path.join(_HERE, "libllama_q4_b200.so")   # LQ4_LIB: development A/B builds only
CSRC = os.path.join(_HERE, "csrc")

MAX_SEQ_LEN = 128 * 1024


class Config(C.Structure):  # reference common.h:9-18
    _fields_ = [("dim", C.c_int), ("hidden_dim", C.c_int), ("n_layers", C.c_int), ("n_heads", C.c_int),
                ("n_kv_heads", C.c_int), ("vocab_size", C.c_int), ("seq_len", C.c_int),
                ("rope_theta", C.c_float)]


class QWeight(C.Structure):  # reference common.h:20-24
    _fields_ = [("weight", C.c_void_p), ("zeros", C.c_void_p), ("scales", C.c_void_p)]


class PerLayerWeight(C.Structure):  # reference common.h:26-36
    _fields_ = [("rms_att_weight", C.c_void_p), ("rms_ffn_weight", C.c_void_p), ("wq_q", QWeight),
                ("wq_k", QWeight), ("wq_v", QWeight), ("wq_o", QWeight), ("wq_gate", QWeight),
                ("wq_up", QWeight), ("wq_down", QWeight)]


class TransformerWeights(C.Structure):  # reference common.h:38-48
    _fields_ = [("token_embedding_table", C.c_void_p), ("wcls", C.c_void_p), ("rms_final_weight", C.c_void_p),
                ("layers", C.POINTER(PerLayerWeight)), ("num_layers", C.c_int)]


class SharedData(C.Structure):  # reference common.h:51-54
    _fields_ = [("pos", C.c_int), ("tokens", C.c_int * MAX_SEQ_LEN)]


class RunState(C.Structure):  # reference common.h:56-72
    _fields_ = [("x", C.c_void_p), ("xb", C.c_void_p), ("hb", C.c_void_p), ("q", C.c_void_p), ("att", C.c_void_p),
                ("logits", C.c_void_p), ("key_cache", C.c_void_p), ("value_cache", C.c_void_p), ("pos", C.c_void_p),
                ("shared_data", C.POINTER(SharedData)), ("logits_array", C.c_void_p)]


class Transformer(C.Structure):  # reference common.h:74-78
    _fields_ = [("config", Config), ("weights", TransformerWeights), ("state", RunState)]


class Sampler(C.Structure):  # reference sampler.h:3-13
    _fields_ = [("vocab_size", C.c_int), ("indices", C.c_void_p), ("tempStorage_scan", C.c_void_p),
                ("tempStorage_sort", C.c_void_p), ("temp_storage_bytes_scan", C.c_size_t),
                ("temp_storage_bytes_sort", C.c_size_t), ("temperature", C.c_float), ("topp", C.c_float),
                ("rng_state", C.c_ulonglong)]


assert C.sizeof(Config) == 32 and C.sizeof(QWeight) == 24 and C.sizeof(PerLayerWeight) == 184
assert C.sizeof(TransformerWeights) == 40 and C.sizeof(RunState) == 88 and C.sizeof(Transformer) == 160
assert C.sizeof(SharedData) == 524292

_P = C.c_void_p
_SIGNATURES = {
    "lq4_init": (C.c_int, [C.c_int]),
    "lq4_get_stream": (_P, []),
    "lq4_set_stream": (None, [_P]),
    "lq4_stream_synchronize": (C.c_int, []),
    "lq4_stream_query": (C.c_int, []),
    "lq4_last_error": (C.c_char_p, []),
    "lq4_sm_count": (C.c_int, []),
    "lq4_set_option": (None, [C.c_char_p, C.c_int]),
    "lq4_debug_trace": (C.c_int, [C.POINTER(C.c_ulonglong), C.POINTER(C.c_int), C.c_int]),
    "lq4_rmsnorm": (None, [_P, _P, _P, C.c_int]),
    "lq4_matmul_fp16": (None, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float]),
    "lq4_matmul_q4": (None, [_P, _P, C.POINTER(QWeight), C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "lq4_qkv_matvec": (None, [_P, _P, _P, _P, C.POINTER(QWeight), C.POINTER(QWeight), C.POINTER(QWeight), C.c_int,
                              C.c_int, C.c_int, _P]),
    "lq4_ffn_matvec_silu": (None, [_P, _P, C.POINTER(QWeight), C.POINTER(QWeight), C.c_int, C.c_int]),
    "lq4_rope_rotation": (None, [_P, _P, C.c_int, C.c_int, C.c_int, _P, C.c_int, C.c_float]),
    "lq4_multi_head_attention": (None, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "lq4_run_llama_network": (None, [_P, C.POINTER(Config), C.POINTER(RunState), C.POINTER(TransformerWeights), C.c_int]),
    "lq4_run_transformer": (None, [C.c_int, C.POINTER(Config), C.POINTER(RunState), C.POINTER(TransformerWeights),
                                   C.c_int, C.POINTER(Sampler)]),
    "lq4_build_sampler": (None, [C.POINTER(Sampler), C.c_int, C.c_float, C.c_float, C.c_ulonglong]),
    "lq4_destroy_sampler": (None, [C.POINTER(Sampler)]),
    "lq4_sample": (None, [C.POINTER(Sampler), C.POINTER(RunState), C.c_int, _P]),
    "lq4_build_transformer": (C.c_int, [C.POINTER(Transformer), C.c_char_p, C.c_int]),
    "lq4_free_transformer": (None, [C.POINTER(Transformer)]),
    "lq4_generate_tokens": (C.c_int, [C.POINTER(Transformer), C.POINTER(Sampler), C.POINTER(C.c_int), C.c_int, C.c_int,
                                      C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_int]),
    "lq4_enqueue_step": (None, [C.POINTER(Transformer), C.POINTER(Sampler), C.c_int, C.c_int]),
    "lq4_step": (C.c_int, [C.POINTER(Transformer), C.POINTER(Sampler), C.c_int, _P, C.POINTER(C.c_int)]),
    "lq4_reset": (None, [C.POINTER(Transformer), C.POINTER(C.c_int), C.c_int]),
    "lq4_memcpy_to_host": (C.c_int, [_P, _P, C.c_size_t]),
    "lq4_gemm_q4": (C.c_int, [_P, _P, C.POINTER(QWeight), C.c_int, C.c_int, C.c_int, C.c_int]),
    "lq4_prefill": (C.c_int, [C.POINTER(Transformer), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, _P, C.POINTER(C.c_float),
                              C.POINTER(C.c_float)]),
    "lq4_tp_config": (C.c_int, [C.c_int, C.c_int]),
    "lq4_tp_export": (C.c_int, [C.POINTER(Transformer), _P]),
    "lq4_tp_import": (C.c_int, [C.POINTER(Transformer), C.c_int, _P]),
    "lq4_write_synth_model": (C.c_size_t, [C.c_char_p, C.POINTER(Config), C.c_ulonglong]),
    "lq4_write_synth_tokenizer": (C.c_size_t, [C.c_char_p, C.c_int]),
}

_lib = None


def build(verbose: bool = False) -> str:
    """Compile the engine in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("building libllama_q4_b200.so failed")
    return LIB_PATH


def lib() -> C.CDLL:
    """The loaded C-ABI library with argtypes set.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        _lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.restype = res
            fn.argtypes = args
    return _lib


def exported_symbols():
    return sorted(_SIGNATURES)


# ---- sizes of a QWeight(K, N), llama2_q4.cu:82-98 ----
def packed_weight_height(K: int) -> int:
    return ((K - 1) // 32 + 1) * 4


def num_groups(K: int) -> int:
    return (K - 1) // 128 + 1


def packed_zeros_height(K: int) -> int:
    return (num_groups(K) - 1) // 8 + 1


LLAMA2_7B = dict(dim=4096, hidden_dim=11008, n_layers=32, n_heads=32, n_kv_heads=32, vocab_size=32000, seq_len=2048,
                 rope_theta=10000.0)
LLAMA2_13B = dict(dim=5120, hidden_dim=13824, n_layers=40, n_heads=40, n_kv_heads=40, vocab_size=32000, seq_len=2048,
                  rope_theta=10000.0)


def weight_bytes_per_token(cfg: dict) -> int:
    """Algorithmic HBM bytes of weights per decoded token (SURVEY.md section 8d / BASELINE.md section 2)."""
    d, h, L, V = cfg["dim"], cfg["hidden_dim"], cfg["n_layers"], cfg["vocab_size"]
    kv = d * cfg["n_kv_heads"] // cfg["n_heads"]

    def q(K, N):
        return N * (packed_weight_height(K) * 4 + packed_zeros_height(K) * 4 + num_groups(K) * 2)

    per_layer = 2 * q(d, d) + 2 * q(d, kv) + 2 * q(d, h) + q(h, d) + 2 * d * 2
    return L * per_layer + V * d * 2 + d * 2 + d * 2


def kv_bytes_at(cfg: dict, pos: int) -> int:
    """KV-cache bytes read+written at position pos: (pos+1) rows of K and V read, one row of each written."""
    kv = cfg["dim"] * cfg["n_kv_heads"] // cfg["n_heads"]
    return (pos + 1) * 2 * kv * 2 * cfg["n_layers"] + 2 * kv * 2 * cfg["n_layers"]


def tp_connect(lib, transformer, rank: int, world: int) -> None:
    """Exchange the CUDA IPC handles of the ranks' activation buffers over torch.distributed (plumbing only) and map them."""
    import torch.distributed as dist
    buf = C.create_string_buffer(64)
    assert lib.lq4_tp_export(C.byref(transformer), buf) == 0, lib.lq4_last_error()
    handles = [None] * world
    dist.all_gather_object(handles, bytes(buf.raw))
    for peer in range(world):
        assert lib.lq4_tp_import(C.byref(transformer), peer, C.create_string_buffer(handles[peer], 64)) == 0, lib.lq4_last_error()
    dist.barrier()

"""CPU-only: the C-ABI library loads and exports every symbol include/llama_q4_b200.h declares; the
synthetic `.bin` writer emits exactly the reference's B1 layout (checked by the oracle's parser and by
the reference's own size formulas); the oracle runs a tiny model end to end."""
import ctypes as C
import os
import re
import tempfile

import numpy as np

import helpers as H

ROOT = H.ROOT


def test_library_exports_every_declared_symbol():
    import llama_cu_awq_b200 as E
    hdr = open(os.path.join(ROOT, "include", "llama_q4_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(lq4_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 25
    lib = C.CDLL(E.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/llama_q4_b200.h but not exported"
    assert set(E.exported_symbols()) == set(declared), "python binding and header disagree"
    E.lib()   # binds argtypes for all of them


def test_struct_layouts_match_reference_contract():
    import llama_cu_awq_b200 as E
    # sizes probed from the reference common.h
    assert C.sizeof(E.Config) == 32 and E.Config.rope_theta.offset == 28
    assert C.sizeof(E.QWeight) == 24
    assert C.sizeof(E.PerLayerWeight) == 184 and E.PerLayerWeight.wq_q.offset == 16 and E.PerLayerWeight.wq_down.offset == 160
    assert C.sizeof(E.TransformerWeights) == 40
    assert C.sizeof(E.SharedData) == 524292 and E.SharedData.tokens.offset == 4
    assert C.sizeof(E.RunState) == 88 and E.RunState.pos.offset == 64 and E.RunState.shared_data.offset == 72
    assert C.sizeof(E.Transformer) == 160


def test_weight_byte_accounting_matches_baseline_md():
    import llama_cu_awq_b200 as E
    assert E.weight_bytes_per_token(E.LLAMA2_7B) == 3_627_302_912
    assert E.weight_bytes_per_token(E.LLAMA2_13B) == 6_920_622_080
    assert sum(E.kv_bytes_at(E.LLAMA2_7B, p) for p in range(256)) // 256 == 67_895_296


def test_synth_bin_layout_and_oracle_forward():
    import llama_cu_awq_b200 as E
    lib = E.lib()
    o = H.oracle()
    for cfg in (H.TINY, H.TINY_GQA):
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "m.bin")
            c = E.Config(**cfg)
            n = lib.lq4_write_synth_model(path.encode(), C.byref(c), 42)
            assert n == os.path.getsize(path)
            raw = np.fromfile(path, dtype=np.uint8)
            hdr = raw[:32].view(np.int32)
            assert list(hdr[:7]) == [cfg[k] for k in ("dim", "hidden_dim", "n_layers", "n_heads", "n_kv_heads", "vocab_size", "seq_len")]
            assert raw[28:32].view(np.float32)[0] == np.float32(cfg["rope_theta"])
            # classifier row 2 (EOS) is zero so greedy never stops early (llama2_q4.cu:477)
            dim, vocab = cfg["dim"], cfg["vocab_size"]
            wcls = raw[32 + vocab * dim * 2: 32 + 2 * vocab * dim * 2].view(np.uint16).reshape(vocab, dim)
            assert not wcls[2].any() and wcls[3].any()
            m = o.oracle_model_open(path.encode())    # fails unless the parsed size equals the file size
            assert m
            logits = np.zeros(vocab, np.uint16)
            toks = [1, 35, 36]
            for pos, tk in enumerate(toks):
                o.oracle_model_forward(m, tk, pos, H.ptr(logits), -1)
                f = logits.view(np.float16).astype(np.float32)
                assert np.isfinite(f).all() and f.std() > 0.05
            assert o.oracle_argmax(H.ptr(logits), vocab) != 2
            o.oracle_model_close(m)
            # same seed -> same bytes; different seed -> different bytes
            path2 = os.path.join(d, "m2.bin")
            lib.lq4_write_synth_model(path2.encode(), C.byref(c), 42)
            assert (np.fromfile(path2, dtype=np.uint8) == raw).all()
            lib.lq4_write_synth_model(path2.encode(), C.byref(c), 43)
            assert (np.fromfile(path2, dtype=np.uint8) != raw).any()


def test_synth_tokenizer_format():
    import llama_cu_awq_b200 as E
    lib = E.lib()
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "tok.bin")
        vocab = 600
        n = lib.lq4_write_synth_tokenizer(path.encode(), vocab)
        raw = open(path, "rb").read()
        assert n == len(raw)
        off = 4
        pieces = []
        for _ in range(vocab):   # tokenizer.h:49-57
            score = np.frombuffer(raw, np.float32, 1, off)[0]; ln = np.frombuffer(raw, np.int32, 1, off + 4)[0]
            pieces.append(raw[off + 8: off + 8 + ln]); off += 8 + ln
            assert score == 0.0
        assert off == len(raw)
        assert pieces[35] == b" " and pieces[3 + ord("h")] == b"h" and pieces[300] == b"[300]"
        assert len(set(pieces)) == vocab


def test_reference_packer_roundtrip_if_built():
    """Format cross-check against the reference's own weight_packer (oracle/_ref/weight_packer): a tiny model in
    the new-AWQ per-tensor layout is repacked by the reference tool and must parse under the B1 layout."""
    packer = os.path.join(ROOT, "oracle", "_ref", "weight_packer")
    if not os.path.exists(packer):
        import pytest
        pytest.skip("reference packer not built (oracle/build_ref.sh)")
    # the packer needs per-tensor input files + config.json; exercising it fully belongs to row f4.
    # Here: it runs and prints its usage (binary is the unmodified reference build).
    import subprocess
    r = subprocess.run([packer], capture_output=True, text=True)
    assert "config" in (r.stdout + r.stderr).lower() or r.returncode != 0


def test_makefile_rebuilds_the_library_when_any_of_its_headers_changes():
    """Every local header the engine's translation unit includes (directly or through another header) is a prerequisite of
    the library target: a stale .so after a header-only edit would make GPU results and sources disagree silently."""
    import re
    csrc = os.path.join(ROOT, "llama_cu_awq_b200", "csrc")
    seen, todo = set(), ["engine.cu"]
    while todo:
        f = todo.pop()
        for inc in re.findall(r'#include "([^"]+)"', open(os.path.join(csrc, f)).read()):
            path = os.path.normpath(os.path.join(os.path.dirname(f), inc))
            if path not in seen and os.path.exists(os.path.join(csrc, path)):
                seen.add(path)
                todo.append(path)
    mk = open(os.path.join(csrc, "Makefile")).read()
    hdrs = re.search(r"^HDRS := (.*)$", mk, re.M).group(1).split()
    missing = sorted(h for h in seen if os.path.normpath(h) not in {os.path.normpath(x) for x in hdrs})
    assert not missing, f"Makefile HDRS lacks {missing}"

"""Row f4: the offline packer `weight_packer_b200` writes the bytes the reference's `weight_packer`
writes (weight_packer.cpp:86-296) for both AWQ dump flavours.  Checked three ways: against the unmodified reference tool
(oracle/_ref/weight_packer, when built), against a numpy restatement of the layout, and -- on a GPU -- by loading the
packed file and decoding from it."""
import json
import os
import subprocess

import numpy as np
import pytest

import helpers as H

PACKER = os.path.join(H.ROOT, "llama_cu_awq_b200", "weight_packer_b200")
REF_PACKER = os.path.join(H.ROOT, "oracle", "_ref", "weight_packer")
ORDER = [0, 2, 4, 6, 1, 3, 5, 7]      # nibble i of an old-AWQ word holds column ORDER[i] of its group of eight


def _pack_rows(q):
    """[..., 8k] nibbles -> [..., k] uint32, element 8w + i in nibble i"""
    q = q.reshape(q.shape[:-1] + (q.shape[-1] // 8, 8)).astype(np.uint32)
    return (q << (4 * np.arange(8, dtype=np.uint32))).sum(axis=-1, dtype=np.uint32)


def _pack_old_awq(q):
    """[rows, cols] nibbles -> old AWQ [rows, cols/8]"""
    g = q.reshape(q.shape[0], q.shape[1] // 8, 8).astype(np.uint32)
    out = np.zeros(g.shape[:2], np.uint32)
    for i in range(8):
        out |= g[:, :, ORDER[i]] << np.uint32(4 * i)
    return out


def _matrix(rng, K, N):
    G = (K - 1) // 128 + 1
    scales = (rng.uniform(0.5, 1.5, (G, N)) / (6.52 * np.sqrt(K))).astype(np.float16).view(np.uint16)     # weights of std ~ 1/sqrt(K)
    return dict(q=rng.integers(0, 16, (K, N), dtype=np.uint8), z=rng.integers(0, 16, (G, N), dtype=np.uint8), s=scales)


def _expected(m, K, N):
    """the B1 layout of one QWeight: qweight [N][K/8] | qzeros [N][ceil(G/8)] | scales [N][G]; the zero nibbles past G continue
    into the next column (what the reference does), zeros after the last one"""
    G = (K - 1) // 128 + 1
    zh = (G - 1) // 8 + 1
    flat = np.concatenate([m["z"].T.reshape(-1), np.zeros(8, np.uint8)])
    zrows = np.stack([flat[x * G: x * G + zh * 8] for x in range(N)])
    return _pack_rows(m["q"].T).tobytes() + _pack_rows(zrows).tobytes() + np.ascontiguousarray(m["s"].T).tobytes()


def _dump(d, cfg, old, seed=0):
    rng = np.random.default_rng(seed)
    dim, hid, kv = cfg["dim"], cfg["hidden_dim"], cfg["dim"] * cfg["n_kv_heads"] // cfg["n_heads"]
    hf = dict(hidden_size=dim, intermediate_size=hid, num_hidden_layers=cfg["n_layers"], num_attention_heads=cfg["n_heads"],
              vocab_size=cfg["vocab_size"], max_position_embeddings=cfg["seq_len"])
    if cfg["n_kv_heads"] != cfg["n_heads"]:
        hf["num_key_value_heads"] = cfg["n_kv_heads"]
    if cfg["rope_theta"] != 10000.0:
        hf["rope_theta"] = cfg["rope_theta"]
    with open(os.path.join(d, "config.json"), "w") as f:
        f.write(json.dumps(hf, separators=(",", ":")))       # the reference looks for `"key":` followed by the number
    w = os.path.join(d, "w")
    os.makedirs(w)
    fp16 = lambda *shape: rng.uniform(0.9, 1.1, shape).astype(np.float16).view(np.uint16)              # norm weights
    head = [("model.embed_tokens.weight", rng.normal(0, 1, (cfg["vocab_size"], dim)).astype(np.float16).view(np.uint16)),
            ("lm_head.weight", rng.normal(0, 0.02, (cfg["vocab_size"], dim)).astype(np.float16).view(np.uint16)),
            ("model.norm.weight", fp16(dim))]
    expect = np.array([dim, hid, cfg["n_layers"], cfg["n_heads"], cfg["n_kv_heads"], cfg["vocab_size"], cfg["seq_len"]], np.int32).tobytes()
    expect += np.float32(cfg["rope_theta"]).tobytes()
    undefined = []                                          # byte ranges the reference leaves undefined (reads past its buffer)
    for name, a in head:
        a.tofile(os.path.join(w, name + ".bin"))
        expect += a.tobytes()
    for l in range(cfg["n_layers"]):
        for name, K, N in (("self_attn.q_proj", dim, dim), ("self_attn.k_proj", dim, kv), ("self_attn.v_proj", dim, kv), ("self_attn.o_proj", dim, dim),
                           ("mlp.up_proj", dim, hid), ("mlp.gate_proj", dim, hid), ("mlp.down_proj", hid, dim)):
            m = _matrix(rng, K, N)
            G = (K - 1) // 128 + 1
            zh = (G - 1) // 8 + 1
            stem = os.path.join(w, f"model.layers.{l}.{name}")
            if old:
                _pack_old_awq(m["q"]).tofile(stem + ".qweight.bin")
                _pack_old_awq(m["z"]).tofile(stem + ".qzeros.bin")
                m["s"].tofile(stem + ".scales.bin")
            else:
                _pack_rows(m["q"].T).tofile(stem + ".qweight.bin")
                zpad = np.zeros((N, zh * 8), np.uint8)
                zpad[:, :G] = m["z"].T
                _pack_rows(zpad).tofile(stem + ".qzeros.bin")
                spad = np.zeros((N, zh * 8), np.uint16)
                spad[:, :G] = m["s"].T
                spad.tofile(stem + ".scales.bin")
            e = _expected(m, K, N)
            if not old:      # new format: the zero words are copied through, so the padding nibbles stay as dumped (zero here)
                e = _pack_rows(m["q"].T).tobytes() + _pack_rows(zpad).tobytes() + np.ascontiguousarray(m["s"].T).tobytes()
            if old and G % 8:
                zoff = len(expect) + N * (K // 8) * 4
                first_bad = next(x for x in range(N) if x * G + zh * 8 > N * G)    # columns whose last zeros word runs past the matrix
                undefined.append((zoff + ((first_bad + 1) * zh - 1) * 4, zoff + N * zh * 4))
            expect += e
        for name in ("input_layernorm.weight", "post_attention_layernorm.weight"):
            a = fp16(dim)
            a.tofile(os.path.join(w, f"model.layers.{l}.{name}.bin"))
            expect += a.tobytes()
    return os.path.join(d, "config.json"), w, expect, undefined


# hidden_dim 1152 -> G = 9 for the down projection: exercises the zero nibbles that run past a column
CFGS = {"tiny": H.TINY, "gqa": H.TINY_GQA, "ragged_groups": dict(H.TINY, hidden_dim=1152)}


@pytest.mark.parametrize("name", list(CFGS))
@pytest.mark.parametrize("old", [0, 1])
def test_packer_bytes(tmp_path, name, old):
    if not os.path.exists(PACKER):
        pytest.fail("weight_packer_b200 is not built (make -C llama_cu_awq_b200/csrc)")
    cfg = CFGS[name]
    config, w, expect, undefined = _dump(str(tmp_path), cfg, old)
    out = str(tmp_path / "model.bin")
    r = subprocess.run([PACKER, config, w, out, str(old)], capture_output=True, text=True)
    assert r.returncode == 0 and "Done!" in r.stdout and f"dim: {cfg['dim']} " in r.stdout, r.stdout[-300:]
    mine = open(out, "rb").read()
    assert mine == expect, "packed file differs from the B1 layout"
    if os.path.exists(REF_PACKER):
        ref_out = str(tmp_path / "ref.bin")
        rr = subprocess.run([REF_PACKER, config, w, ref_out, str(old)], capture_output=True, text=True)
        assert rr.returncode == 0
        assert rr.stdout == r.stdout, "same console output as the reference tool"
        ref = bytearray(open(ref_out, "rb").read())
        got = bytearray(mine)
        assert len(ref) == len(got)
        for a, b in undefined:          # the reference read past its buffer there; nobody reads these nibbles back
            ref[a:b] = got[a:b]
        assert ref == got, "packed file differs from the reference packer's"


def test_packer_usage_and_missing_input(tmp_path):
    r = subprocess.run([PACKER], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("usage: weight_packer")
    config, w, _, _ = _dump(str(tmp_path), H.TINY, 0)
    os.remove(os.path.join(w, "model.layers.1.mlp.gate_proj.scales.bin"))
    r = subprocess.run([PACKER, config, w, str(tmp_path / "o.bin"), "0"], capture_output=True, text=True)
    assert r.returncode == 1 and "Unable to open" in r.stdout


@pytest.mark.gpu
def test_packed_file_loads_and_decodes(tmp_path):
    """the file the packer writes goes through the loader and the fused decode path: logits equal the CPU oracle's on the same
    file (same tolerance as test_gpu_e2e.test_cpu_oracle_forward_matches_gpu: host libm vs libdevice in RoPE / softmax)"""
    import ctypes as C
    import llama_cu_awq_b200 as E
    cfg = H.TINY
    config, w, _, _ = _dump(str(tmp_path), cfg, 1, seed=3)
    out = str(tmp_path / "model.bin")
    assert subprocess.run([PACKER, config, w, out, "1"], capture_output=True).returncode == 0
    lib, o = E.lib(), H.oracle()
    assert lib.lq4_init(0) == 0
    t = E.Transformer()
    lib.lq4_build_transformer(C.byref(t), out.encode(), 0)
    s = E.Sampler()
    lib.lq4_build_sampler(C.byref(s), cfg["vocab_size"], 0.0, 0.9, 1)
    m = o.oracle_model_open(out.encode())
    assert m
    try:
        prompt = [1, 300, 41, 77, 263, 12]
        arr = (C.c_int * len(prompt))(*prompt)
        lib.lq4_reset(C.byref(t), arr, len(prompt))
        for pos in range(len(prompt) - 1):
            mine = np.zeros(cfg["vocab_size"], np.uint16)
            nxt = C.c_int(0)
            lib.lq4_step(C.byref(t), C.byref(s), 0, H.ptr(mine), C.byref(nxt))
            lg = np.zeros(cfg["vocab_size"], np.uint16)
            o.oracle_model_forward(m, prompt[pos], pos, H.ptr(lg), -1)
            a, b = lg.view(np.float16).astype(np.float32), mine.view(np.float16).astype(np.float32)
            assert np.isfinite(b).all() and np.abs(b).max() > 0.05
            ok = (H.ulp_diff_f16(lg, mine) <= 4) | (np.abs(a - b) < 2e-3 * np.maximum(np.abs(a), 1.0))
            assert ok.all(), f"pos {pos}: {np.count_nonzero(~ok)} logits off"
    finally:
        o.oracle_model_close(m)
        lib.lq4_destroy_sampler(C.byref(s))
        lib.lq4_free_transformer(C.byref(t))

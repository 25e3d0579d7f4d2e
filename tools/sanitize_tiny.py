#!/usr/bin/env python
"""Run under compute-sanitizer (gpurun): a tiny model decodes a few tokens through the fused persistent kernel and through
the op-by-op path.  `python tools/sanitize_tiny.py [steps]`"""
import ctypes as C
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
import llama_cu_awq_b200 as E

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
lib = E.lib()
assert lib.lq4_init(0) == 0
for cfg in (H.TINY, H.TINY_GQA):
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "m.bin")
        c = E.Config(**cfg)
        assert lib.lq4_write_synth_model(path.encode(), C.byref(c), 7) == os.path.getsize(path)
        t = E.Transformer()
        assert lib.lq4_build_transformer(C.byref(t), path.encode(), 0) == 0
        s = E.Sampler()
        lib.lq4_build_sampler(C.byref(s), cfg["vocab_size"], 0.0, 0.9, 1)
        for fused in (1, 0):
            lib.lq4_set_option(b"fused", fused)
            tok = (C.c_int * 2)(1, 35)
            lib.lq4_reset(C.byref(t), tok, 2)
            nxt = C.c_int(0)
            ids = []
            for i in range(steps):
                lib.lq4_step(C.byref(t), C.byref(s), 1 if i >= 1 else 0, None, C.byref(nxt))
                ids.append(int(nxt.value))
            print("fused", fused, "ids", ids)
        lib.lq4_set_option(b"fused", 1)
        lib.lq4_free_transformer(C.byref(t))
print("sanitize_tiny: done")

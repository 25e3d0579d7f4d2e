#!/usr/bin/env python
"""Development aid (run under ncu / gpurun): one prefill pass (batch 8 x seq 2048 of the 7B model) so that a launch list
(ncu --metrics gpu__time_duration.sum --clock-control none) shows where the non-GEMM time of the pass goes.
   python tools/prof_prefill.py [batch seq passes]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import llama_cu_awq_b200 as E
import bench as B

batch, seq, passes = (int(x) for x in (sys.argv[1:4] + ["8", "2048", "1"][len(sys.argv) - 1:]))
lib = E.lib()
assert lib.lq4_init(0) == 0
cfg = B.model_cfg("7b")
path, tok = B.ensure_files(lib, E, "7b", cfg)
t = E.Transformer()
lib.lq4_build_transformer(C.byref(t), path.encode(), 0)
rng = np.random.default_rng(5)
toks = rng.integers(3, cfg["vocab_size"], size=(batch, seq)).astype(np.int32)
toks[:, 0] = 1
ms, msg = C.c_float(0), C.c_float(0)
for rep in range(passes):
    assert lib.lq4_prefill(C.byref(t), toks.ctypes.data_as(C.POINTER(C.c_int)), batch, seq, -1, None, C.byref(ms), C.byref(msg)) == 0
    print(f"prefill {batch} x {seq}: {ms.value:.1f} ms in all, {msg.value:.1f} ms in the GEMMs")

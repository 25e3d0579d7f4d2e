#!/usr/bin/env python
"""Development aid (run under ncu / gpurun): one tensor-core INT4 GEMM at a prefill shape.  python tools/prof_gemm.py [M K N reps]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
import llama_cu_awq_b200 as E
import torch

M, K, N, reps = (int(x) for x in (sys.argv[1:5] + ["16384", "4096", "11008", "3"][len(sys.argv) - 1:]))
lib = E.lib()
assert lib.lq4_init(0) == 0
rng = np.random.default_rng(1)
w, z, s = H.random_qweight(rng, K, N)
x = torch.randn(M, K, dtype=torch.half, device="cuda")
y = torch.empty(M, N, dtype=torch.half, device="cuda")
dw, dz, ds = H.to_dev(w), H.to_dev(z), H.to_dev(s.view(np.uint16))
q = E.QWeight(dw.data_ptr(), dz.data_ptr(), ds.data_ptr())
stream = torch.cuda.ExternalStream(lib.lq4_get_stream())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
for r in range(reps):
    e0.record(stream)
    assert lib.lq4_gemm_q4(y.data_ptr(), x.data_ptr(), C.byref(q), M, K, N, 0) == 0
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"gemm M={M} K={K} N={N}: {ms:.3f} ms, {2.0 * M * K * N / ms / 1e9:.1f} TFLOP/s")

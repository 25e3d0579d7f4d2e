#!/usr/bin/env python
"""Profiling aid (run under gpurun/ncu): the fused gate/up+SiLU INT4 GEMV at the 7B shape on random device
buffers, through the C ABI.  argv: [K N reps]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import llama_cu_awq_b200 as E

K, N, reps = (int(a) for a in (sys.argv[1:4] + ["4096", "11008", "4"][len(sys.argv) - 1:]))
lib = E.lib()
assert lib.lq4_init(0) == 0
g = torch.Generator(device="cuda").manual_seed(1)


def qw():
    w = torch.randint(-2**31, 2**31 - 1, (N, E.packed_weight_height(K)), dtype=torch.int32, device="cuda", generator=g)
    z = torch.randint(-2**31, 2**31 - 1, (N, E.packed_zeros_height(K)), dtype=torch.int32, device="cuda", generator=g)
    s = (torch.rand((N, E.num_groups(K)), device="cuda", generator=g) * 0.004 + 0.002).half()
    return E.QWeight(w.data_ptr(), z.data_ptr(), s.data_ptr()), (w, z, s)


mats = [(qw(), qw()) for _ in range(8)]
x = torch.randn(K, device="cuda", generator=g).half()
out = torch.zeros(N, device="cuda", dtype=torch.half)
torch.cuda.synchronize()
stream = torch.cuda.ExternalStream(lib.lq4_get_stream())
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(2):
    with torch.cuda.stream(stream):
        ev0.record(stream)
    for r in range(reps):
        (ga, _), (ua, _) = mats[r % len(mats)]
        lib.lq4_ffn_matvec_silu(out.data_ptr(), x.data_ptr(), C.byref(ga), C.byref(ua), K, N)
    with torch.cuda.stream(stream):
        ev1.record(stream)
    torch.cuda.synchronize()
print("ffn K=%d N=%d: %.2f us per launch" % (K, N, ev0.elapsed_time(ev1) * 1000 / reps))

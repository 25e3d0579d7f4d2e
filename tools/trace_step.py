#!/usr/bin/env python
"""Development aid (run under gpurun): per-op time line of one fused decode step at a given position."""
import ctypes as C
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import llama_cu_awq_b200 as E
import bench as B

model = sys.argv[1] if len(sys.argv) > 1 else "7b"
npos = int(sys.argv[2]) if len(sys.argv) > 2 else 128
lib = E.lib()
assert lib.lq4_init(0) == 0
cfg = B.model_cfg(model)
path, tok = B.ensure_files(lib, E, model, cfg)
t = E.Transformer()
lib.lq4_build_transformer(C.byref(t), path.encode(), 0)
s = E.Sampler()
lib.lq4_build_sampler(C.byref(s), cfg["vocab_size"], 0.0, 0.9, 1)
bos = (C.c_int * 1)(1)
lib.lq4_reset(C.byref(t), bos, 1)
lib.lq4_set_option(b"trace", 1)
trace_op = int(sys.argv[3]) if len(sys.argv) > 3 else 8
lib.lq4_set_option(b"trace_op", trace_op)
for i in range(npos):
    lib.lq4_enqueue_step(C.byref(t), C.byref(s), i + 1, 1)
assert lib.lq4_stream_synchronize() == 0
N = 32768
ts = (C.c_ulonglong * N)()
kinds = (C.c_int * N)()
n = lib.lq4_debug_trace(ts, kinds, N)
names = {0: "gemv", 1: "ffn", 2: "cls", 3: "attn", 4: "argmax"}
agg = collections.OrderedDict()
for i in range(n - 1):
    k = (names[kinds[i] // 100000], kinds[i] % 100000)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += (ts[i + 1] - ts[i]) / 1000.0
total = (ts[n - 1] - ts[0]) / 1000.0
print(f"step at pos {npos - 1}: {total:.1f} us over {n - 1} ops")
for k, (cnt, us) in agg.items():
    print(f"  {k[0]:7s} K={k[1]:6d}: {cnt:4d} ops, {us:9.1f} us total, {us / cnt:7.2f} us each ({100 * us / total:5.1f}%)")
first = [(names[kinds[i] // 100000], kinds[i] % 100000, (ts[i + 1] - ts[i]) / 1000.0) for i in range(5, 12)]
print("  layer 1 ops:", first)

# phases of op `trace_op` on every CTA: 0 arrive, 1 barrier passed, 2 staged, 3 first weights, 4 warp0 done, 5 all warps done
import numpy as np
nb = min(148, cfg["n_heads"]) if names[kinds[trace_op] // 100000] == "attn" else 148
ph = np.array([[ts[2048 + b * 8 + k] for k in range(8)] for b in range(nb)], dtype=np.float64)
t0 = ph[:, 1].min()
rel = (ph - t0) / 1000.0
print(f"op {trace_op} ({names[kinds[trace_op] // 100000]} K={kinds[trace_op] % 100000}) phases in us relative to the first CTA past the barrier [min / median / max over CTAs]:")
for k, name in enumerate(["arrive(prev done)", "barrier passed", "x staged", "first weights", "warp0 done", "all warps done", "warp0 task 1 done", "warp0 task 2 done"]):
    print(f"  {name:18s} {rel[:, k].min():8.2f} {np.median(rel[:, k]):8.2f} {rel[:, k].max():8.2f}")

smid = [int(ts[2048 + 148 * 8 + b * 16 + 11]) for b in range(nb)]
late = [b for b in range(nb) if rel[b, 2] > np.median(rel[:, 2]) + 0.7]
print(f"  CTAs with x staged > median + 0.7 us: {len(late)}: " + ", ".join(f"cta{b}/sm{smid[b]}:{rel[b, 2]:.2f}" for b in late))
for k, name in ((0, "arrive"), (2, "x staged"), (5, "all warps done")):
    order = np.argsort(-rel[:, k])[:8]
    print(f"  slowest CTAs at '{name}': " + ", ".join(f"{int(b)}:{rel[b, k]:.2f}" for b in order))
print("  duration x staged -> all warps done [min/median/max]: %.2f %.2f %.2f; slowest: %s" % (
    (rel[:, 5] - rel[:, 2]).min(), np.median(rel[:, 5] - rel[:, 2]), (rel[:, 5] - rel[:, 2]).max(),
    ", ".join(f"{int(b)}:{rel[b, 5] - rel[b, 2]:.2f}" for b in np.argsort(-(rel[:, 5] - rel[:, 2]))[:8])))
cy = np.array([[ts[2048 + 148 * 8 + b * 16 + k] for k in range(8)] for b in range(nb)], dtype=np.float64)
names_c = ["barrier passed", "raw x staged", "rms scale known", "pairs staged", "meta landed", "first weights", "first task done", "(arrive)"]
print("  SM-clock phases of warp 0, cycles since 'barrier passed' [median over CTAs] (1965 cycles = 1 us):")
for k in range(1, 7):
    print(f"    {names_c[k]:18s} {np.median(cy[:, k] - cy[:, 0]):9.0f}")
print(f"    arrive -> passed   {np.median(cy[:, 0] - cy[:, 7]):9.0f}")
ex = np.array([[ts[2048 + 148 * 8 + b * 16 + k] for k in range(8, 11)] for b in range(nb)], dtype=np.float64)
ahead0, ahead1 = ex[:, 0] - ex[:, 2], ex[:, 1] - ex[:, 2]
print(f"  ring chunks the producer was ahead of this op's first chunk [min/median/max over CTAs]: when x was staged {ahead0.min():.0f}/{np.median(ahead0):.0f}/{ahead0.max():.0f}, "
      f"when warp 0 got its first weights {ahead1.min():.0f}/{np.median(ahead1):.0f}/{ahead1.max():.0f}")
if os.environ.get("TRACE_DUMP"):
    print("  per CTA, sorted by SM id: cta sm | arrive passed staged first-weights warp0-done all-done (us)")
    for b in sorted(range(nb), key=lambda b: smid[b]):
        print(f"    {b:3d} {smid[b]:3d} | " + " ".join(f"{rel[b, k]:6.2f}" for k in range(6)))

# per consumer warp (SM clock, relative to this CTA's 'barrier passed'): first weights, task k done, left the op
CLK = 1965.0
wm = np.array([[[ts[8192 + b * 128 + w * 8 + k] for k in range(8)] for w in range(11)] for b in range(nb)], dtype=np.float64)
ref = cy[:, 0]
print("  per warp, us after the CTA's 'barrier passed' [median over CTAs]: first weights | task 1..5 done | left the op")
for w in range(11):
    row = []
    for k in range(8):
        v = wm[:, w, k]
        ok = v > ref        # marks of this step only
        row.append(f"{np.median((v[ok] - ref[ok]) / CLK):6.2f}" if ok.sum() > nb // 2 else "   -  ")
    print(f"    warp {w:2d}: {row[0]} | " + " ".join(row[1:6]) + f" | {row[7]}")
pi = np.array([[ts[27136 + b * 32 + k] for k in range(32)] for b in range(nb)], dtype=np.float64)
if (pi > 0).any():
    print("  producer: issue time of the op's k-th chunk, us after the CTA's 'barrier passed' [median over CTAs] (builds with -DLQ4_PROD_TRACE):")
    print("    " + " ".join(f"{np.median((pi[:, k] - ref) / CLK):6.2f}" for k in range(32) if (pi[:, k] > 0).any()))

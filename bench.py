#!/usr/bin/env python
"""bench.py -- decode tokens/sec of Llama-2-7B AWQ-w4-g128 (random-init, synthetic) on B200.

Contract (one JSON line on stdout, rank 0):
  python bench.py --gpus N --steps K --warmup W [--impl reference] [--model 7b|13b|tiny]

A "step" is one decode forward + greedy sample (seq_len=1) at consecutive positions 0..K-1 of the
workload `Llama-2-7B w4-g128 random-init .bin, greedy decode` (BASELINE.json configs[1]); the default
K=256 is the reference's `-n 256`.

  value  : tokens/s with everything resident in HBM, K steps enqueued back to back on the engine
           stream (position and token live on the device), timed with CUDA events on that stream.
  e2e    : tokens/s through the C-ABI host-buffer call lq4_generate_tokens (prompt tokens in host
           memory, ids out to host memory; the per-token pinned-memory token/position hand-off and the
           host wait are inside its timed loop, as in the reference's generate()).
  roofline: dominant kernel = interp_kernel, the persistent kernel that IS the decode step (one launch per
           token): algorithmic bytes (weights + KV) / mean launch duration against MEASURED_PEAKS.json;
           roofline_ffn_op is the largest single op (gate/up+SiLU) alone through the operator API.
  cpu_baseline: the oracle port (oracle/cpu_ref.c) single-threaded on a bounded sample.
--impl reference: the reference has no CPU implementation (CUDA only), so this arm times the oracle
  port with all host threads on a bounded sample, and also reports the UNMODIFIED reference CUDA build
  (oracle/_ref/llama2_q4_ref) run on the same GPU/.bin as `reference_cuda` (its own achieved tok/s line).
"""
import argparse
import ctypes as C
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

PROMPT = "hello"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def model_cfg(name):
    import llama_cu_awq_b200 as E
    import helpers as H
    return {"7b": E.LLAMA2_7B, "13b": E.LLAMA2_13B, "tiny": H.SMALL}[name]


def scratch_dir():
    for d in ("/dev/shm", "/tmp"):
        try:
            st = os.statvfs(d)
            if st.f_bavail * st.f_frsize > 10 << 30:
                return d
        except OSError:
            pass
    return "/tmp"


def ensure_files(lib, E, name, cfg, rank=0):
    d = scratch_dir()
    path = os.path.join(d, f"lq4_synth_{name}.bin")
    tok = os.path.join(d, f"lq4_synth_tok_{cfg['vocab_size']}.bin")
    c = E.Config(**cfg)
    if rank == 0:
        if not os.path.exists(path + ".ok"):
            n = lib.lq4_write_synth_model(path.encode(), C.byref(c), 0x5EED)
            assert n == os.path.getsize(path) and n > 0
            open(path + ".ok", "w").write(str(n))
        if not os.path.exists(tok):
            assert lib.lq4_write_synth_tokenizer(tok.encode(), cfg["vocab_size"]) > 0
    return path, tok


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def cpu_baseline_sample(path, cfg, threads):
    """Oracle port on the host: one token at pos 0 through `nl` layers, scaled to the full depth plus the
    classifier measured separately (bounded to ~10-30 s)."""
    import numpy as np
    import helpers as H
    o = H.oracle()
    o.oracle_set_threads(threads)
    m = o.oracle_model_open(path.encode())
    assert m, "oracle could not map the .bin"
    nl = 1 if threads == 1 else min(4, cfg["n_layers"])
    t0 = time.perf_counter()
    o.oracle_model_forward(m, 1, 0, None, nl)
    t_layers = time.perf_counter() - t0
    lg = np.zeros(cfg["vocab_size"], np.uint16)
    t0 = time.perf_counter()
    o.oracle_model_forward(m, 1, 0, H.ptr(lg), 0)
    t_cls = time.perf_counter() - t0
    o.oracle_model_close(m)
    per_token = t_layers / nl * cfg["n_layers"] + t_cls
    return {"value": 1.0 / per_token, "unit": "tokens/s", "cores": threads, "kind": "port",
            "sample": f"1 token at pos 0: {nl} of {cfg['n_layers']} layers ({t_layers:.2f}s) scaled to full depth + classifier ({t_cls:.2f}s), oracle/cpu_ref.c"}


def max_over_ranks(value, world, device="cuda"):
    """Replicas: the job's time is the slowest rank's (max over ranks); a plain float in, a plain float out."""
    if world <= 1:
        return float(value)
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_throughput(units_per_rank, seconds, world, device="cuda"):
    """Whole-job throughput of `world` independent replicas: all units / the slowest rank's time."""
    return world * units_per_rank / max_over_ranks(seconds, world, device)


def run_reference_arm(args):
    import llama_cu_awq_b200 as E
    lib = E.lib()
    cfg = model_cfg(args.model)
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    path, tok = ensure_files(lib, E, args.model, cfg)
    threads = os.cpu_count() or 1
    base = cpu_baseline_sample(path, cfg, threads)
    out = {"metric": "decode tokens/sec (seq_len=1)", "impl": "reference", "value": base["value"], "unit": "tokens/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / base["value"],
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int4 weights, fp16 storage, fp32 accumulate",
           "data": "synthetic", "config": {"workload": f"Llama-2-{args.model.upper()} w4-g128 random-init .bin, greedy decode"},
           "cpu_baseline": base,
           "e2e": {"value": base["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "note": "the reference is CUDA-only (no CPU path, SURVEY.md 8d): this arm is the oracle port on all host threads; "
                   "reference_cuda is the unmodified reference built for sm_100a run on this GPU"}
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "llama2_q4_ref")
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if os.path.exists(ref_bin) and has_gpu:
        n = max(2, min(args.steps if args.steps > 0 else 256, cfg["seq_len"]))
        best = None
        for _ in range(2):   # first run pays page-cache / driver warm-up
            r = subprocess.run([ref_bin, path, "-z", tok, "-t", "0", "-n", str(n), "-i", PROMPT], capture_output=True, text=True, timeout=900)
            mt = re.search(r"achieved tok/s: ([0-9.]+)\. Tokens: (\d+), seconds: ([0-9.eE+-]+)", r.stdout)
            if mt and (best is None or float(mt.group(1)) > best["value"]):
                best = {"value": float(mt.group(1)), "unit": "tokens/s", "tokens": int(mt.group(2)), "seconds": float(mt.group(3)),
                        "how": f"oracle/_ref/llama2_q4_ref <bin> -t 0 -n {n}: its own 'achieved tok/s' line (wall clock incl. graph capture), best of 2"}
        out["reference_cuda"] = best if best else {"unavailable": (r.stderr or r.stdout)[-200:]}
    else:
        out["reference_cuda"] = {"unavailable": "no GPU or oracle/_ref/llama2_q4_ref not built"}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--model", default="7b")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--parallelism", default="replicas", choices=["replicas", "tp"],
                    help="N>1: independent replicas (weak scaling, default) or tensor-parallel decode of ONE stream (strong scaling)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import numpy as np
    import torch
    import llama_cu_awq_b200 as E

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = E.lib()
    assert lib.lq4_init(local_rank) == 0
    tp = world > 1 and args.parallelism == "tp"
    if tp:
        assert lib.lq4_tp_config(rank, world) == 0
    cfg = model_cfg(args.model)
    K = max(1, min(args.steps, cfg["seq_len"] - 1))
    W = max(3, args.warmup)
    path, tok = ensure_files(lib, E, args.model, cfg, rank)
    if world > 1:
        dist.barrier()

    t = E.Transformer()
    lib.lq4_build_transformer(C.byref(t), path.encode(), 0)
    s = E.Sampler()
    lib.lq4_build_sampler(C.byref(s), cfg["vocab_size"], 0.0, 0.9, 1)
    if tp:
        E.tp_connect(lib, t, rank, world)     # CUDA IPC handles of the ranks' activation buffers, exchanged over torch.distributed
    stream = torch.cuda.ExternalStream(lib.lq4_get_stream())
    bos = (C.c_int * 1)(1)

    def enqueue(n_steps):
        lib.lq4_reset(C.byref(t), bos, 1)
        for i in range(n_steps):
            lib.lq4_enqueue_step(C.byref(t), C.byref(s), i + 1, 1)

    # ---- warm-up ----
    if world > 1:
        dist.barrier()
    enqueue(min(cfg["seq_len"] - 1, max(W, 130 if K > 128 else W)))
    if K > 256:
        enqueue(K)
    assert lib.lq4_stream_synchronize() == 0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- value: K steps back to back, device-timed on the engine stream ----
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lib.lq4_reset(C.byref(t), bos, 1)
    with torch.cuda.stream(stream):
        ev0.record(stream)
    for i in range(K):
        lib.lq4_enqueue_step(C.byref(t), C.byref(s), i + 1, 1)
    with torch.cuda.stream(stream):
        ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    ms = max_over_ranks(ms, world)
    tokens_dev = [int(t.state.shared_data.contents.tokens[i]) for i in range(K + 1)]

    # ---- e2e: the host-buffer API, wall clock inside the call (host wait + pinned hand-off per token) ----
    out = (C.c_int * (K + 1))()
    secs = C.c_double(0)
    barrier()
    n = lib.lq4_generate_tokens(C.byref(t), C.byref(s), bos, 1, K + 1, out, C.byref(secs), 1)
    e2e_val = (n - 1) / secs.value if secs.value > 0 else None
    assert list(out)[1:n] == tokens_dev[1:n], "pipelined host API and raw enqueue disagree on token ids"
    if world > 1:
        e2e_val = (n - 1) / max_over_ranks(secs.value, world)
    clk = clocks.stop()             # sampled across both timed regions (device-timed steps and the end-to-end call)
    units = 1 if tp else world        # tensor parallel: all ranks decode ONE stream; replicas: one stream each

    # ---- roofline ----
    # The decode step is ONE launch of the persistent kernel (interp_kernel), so the dominant kernel's launch
    # duration is ms/K measured above with CUDA events on the engine stream.  Algorithmic bytes per launch =
    # weights + KV cache rows touched (DESIGN.md "bytes per unit"), averaged over the K positions.
    peak, peak_src = load_peaks()
    d, h, L = cfg["dim"], cfg["hidden_dim"], cfg["n_layers"]
    wbytes = E.weight_bytes_per_token(cfg)
    kvbytes = sum(E.kv_bytes_at(cfg, p) for p in range(K)) / K
    step_gbs = (wbytes + kvbytes) / (ms / K * 1e-3) / 1e9 / (world if tp else 1)     # per GPU: each rank streams 1/world of the bytes
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "step_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(args.model, {}).get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": "interp_kernel (whole decode step: 1 launch per token)", "achieved": step_gbs, "peak": peak,
                "unit": "GB/s", "frac": step_gbs / peak, "traffic": traffic, "bytes_per_launch": wbytes + kvbytes,
                "us_per_launch": ms / K * 1000.0, "peak_source": peak_src,
                "how": "algorithmic bytes (weights + mean KV over the K positions) / mean launch duration, CUDA events on the engine stream; "
                       "traffic = dram read+write of one launch from the committed ncu --set full capture (profiles/)"}
    # the largest single op (gate/up+SiLU INT4 GEMV, 45% of a layer's bytes) alone through the operator API, all layers' weights in rotation
    ffn_bytes = 2 * h * (E.packed_weight_height(d) * 4 + E.packed_zeros_height(d) * 4 + E.num_groups(d) * 2)
    layers = t.weights.layers
    reps = 3
    for _ in range(2):
        for l in range(L):
            lib.lq4_ffn_matvec_silu(t.state.hb, t.state.xb, C.byref(layers[l].wq_gate), C.byref(layers[l].wq_up), d, h)
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        ev0.record(stream)
    for _ in range(reps):
        for l in range(L):
            lib.lq4_ffn_matvec_silu(t.state.hb, t.state.xb, C.byref(layers[l].wq_gate), C.byref(layers[l].wq_up), d, h)
    with torch.cuda.stream(stream):
        ev1.record(stream)
    torch.cuda.synchronize()
    ffn_us = ev0.elapsed_time(ev1) * 1000.0 / (reps * L)
    ach = ffn_bytes / (ffn_us * 1e-6) / 1e9
    roofline_ffn = {"bound": "hbm", "kernel": "interp_kernel, single op lq4_ffn_matvec_silu (K=%d N=%d), launch overhead included" % (d, h),
                    "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "bytes_per_launch": ffn_bytes, "us_per_launch": ffn_us}
    # BASELINE.json configs[0]: single INT4 GEMV M=1 K=N=4096 g128 (the o-projection), one matrix per layer in rotation
    # (32 x 8.7 MB > L2 would be needed for a cold number; 7B has exactly 32 such matrices = 279 MB)
    gemv_bytes = d * (E.packed_weight_height(d) * 4 + E.packed_zeros_height(d) * 4 + E.num_groups(d) * 2)
    for l in range(L):
        lib.lq4_matmul_q4(t.state.x, t.state.xb, C.byref(layers[l].wq_o), d, d, 0, -1, None)
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        ev0.record(stream)
    for _ in range(reps):
        for l in range(L):
            lib.lq4_matmul_q4(t.state.x, t.state.xb, C.byref(layers[l].wq_o), d, d, 0, -1, None)
    with torch.cuda.stream(stream):
        ev1.record(stream)
    torch.cuda.synchronize()
    gemv_us = ev0.elapsed_time(ev1) * 1000.0 / (reps * L)
    gemv_op = {"kernel": "interp_kernel, single op lq4_matmul_q4 (K=N=%d g128), launch overhead included" % d, "bytes_per_launch": gemv_bytes,
               "us_per_launch": gemv_us, "achieved": gemv_bytes / (gemv_us * 1e-6) / 1e9, "unit": "GB/s", "peak": peak,
               "frac": gemv_bytes / (gemv_us * 1e-6) / 1e9 / peak}
    value = aggregate_throughput(K, ms * 1e-3, 1) * units      # ms is already the max over ranks

    line = {"metric": "decode tokens/sec (seq_len=1)", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong" if tp else "weak", "vs_baseline": None,
            "dtype": "int4 weights, fp16 storage, fp32 accumulate", "data": "synthetic",
            "config": {"workload": f"Llama-2-{args.model.upper()} w4-g128 random-init .bin, greedy decode -n {K}, batch 1",
                       "l2": "inputs larger than L2 (3.6 GB of weights per step)", "parallelism": ("tp%d" % world if tp else "replicas x%d" % world) if world > 1 else "1 GPU"},
            "clocks": clk,
            "e2e": {"value": (e2e_val * units) if e2e_val else None, "unit": "tokens/s", "h2d_bytes_per_step": 4, "d2h_bytes_per_step": 8,
                    "how": "lq4_generate_tokens(host prompt ids -> host ids), wall clock of its loop, pipelined launch; per step the kernel "
                           "reads the token id from pinned host memory and writes the new id and position back to it"},
            "gpu_launches": K,
            "roofline": roofline,
            "roofline_ffn_op": roofline_ffn,
            "gemv_4096_op": gemv_op}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_baseline_sample(path, cfg, 1)
        except Exception as e:  # the oracle is test infrastructure: report, never fall back to it
            line["cpu_baseline"] = {"unavailable": str(e)[:200]}
    lib.lq4_free_transformer(C.byref(t))
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

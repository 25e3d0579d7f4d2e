#!/usr/bin/env python
"""bench.py -- decode tokens/sec of Llama-2-7B AWQ-w4-g128 (random-init, synthetic) on B200.

Contract (one JSON line on stdout, rank 0):
  python bench.py --gpus N --steps K --warmup W [--impl reference] [--model 7b|13b|tiny] [--parallelism tp|replicas]

A "step" is one decode forward + greedy sample (seq_len=1) at consecutive positions 0..K-1 of the
workload `Llama-2-7B w4-g128 random-init .bin, greedy decode` (BASELINE.json configs[1]); the default
K=256 is the reference's `-n 256`.

  value  : tokens/s with everything resident in HBM, K steps enqueued back to back on the engine
           stream (position and token live on the device), timed with CUDA events on that stream,
           max over ranks.
           N = 1: one GPU.   N > 1 (default --parallelism tp): ONE stream decoded tensor-parallel over
           N GPUs (BASELINE.json configs[3], "scaling": "strong"); the ids of the TP run are compared
           IN THIS RUN with the ids the same ranks produce on one GPU (`tp.ids_match_single_gpu`), and
           the N independent replicas measured in the same run are reported beside it (`replicas`).
  e2e    : tokens/s through the C-ABI host-buffer call lq4_generate_tokens (prompt tokens in host
           memory, ids out to host memory; the per-token pinned-memory token/position hand-off and the
           host wait are inside its timed loop, as in the reference's generate(); ns clock).
  roofline: dominant kernel = interp_kernel, the persistent kernel that IS the decode step (one launch per
           token): algorithmic bytes (weights + KV) / mean launch duration against MEASURED_PEAKS.json.
  cpu_baseline: the oracle port (oracle/cpu_ref.c) single-threaded on REAL tokens (all layers, positions
           0..1, KV cache and attention included); `cpu_baseline_all_threads` the same on every host core.
--impl reference: the reference is a CUDA program with no CPU path, so this arm times the UNMODIFIED
  reference (oracle/_ref/libq4ref.so = its translation unit behind a C shim, built by oracle/build_ref.sh)
  on the same GPU and `.bin`: its own loop body (cudaStreamSynchronize + run_transformer, llama2_q4.cu:
  465-470) for K steps in steady state (graphs already captured), CUDA events on its stream.  It maps
  neither libllama_q4_b200.so nor the oracle into the timed process state beyond the CPU-port side record
  (`cpu_baseline`: the oracle port on all host threads over 3 real tokens, measured, not extrapolated).
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
SEED = 0x5EED
METRIC = "decode tokens/sec (seq_len=1)"
DTYPE = "int4 weights, fp16 storage, fp32 accumulate"

MODELS = {
    "7b": dict(dim=4096, hidden_dim=11008, n_layers=32, n_heads=32, n_kv_heads=32, vocab_size=32000, seq_len=2048, rope_theta=10000.0),
    "13b": dict(dim=5120, hidden_dim=13824, n_layers=40, n_heads=40, n_kv_heads=40, vocab_size=32000, seq_len=2048, rope_theta=10000.0),
    "tiny": dict(dim=1024, hidden_dim=2816, n_layers=3, n_heads=8, n_kv_heads=8, vocab_size=2048, seq_len=512, rope_theta=10000.0),
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def model_cfg(name):
    return dict(MODELS[name])


def scratch_dir(need=10 << 30):
    for d in ("/dev/shm", "/tmp"):
        try:
            st = os.statvfs(d)
            if st.f_bavail * st.f_frsize > need:
                return d
        except OSError:
            pass
    return "/tmp"


def synth_paths(name, cfg):
    d = scratch_dir()
    for cand in ("/dev/shm", "/tmp"):       # a file written earlier in this box's life wins over free-space heuristics
        if os.path.exists(os.path.join(cand, f"lq4_synth_{name}.bin.ok")):
            d = cand
            break
    return os.path.join(d, f"lq4_synth_{name}.bin"), os.path.join(d, f"lq4_synth_tok_{cfg['vocab_size']}.bin")


def ensure_files(lib, E, name, cfg, rank=0):
    """Synthetic `.bin` + tokenizer (seeded).  `lib` may be None: then the stand-alone writer
    llama_cu_awq_b200/gen_synth_bin (same synth.h, same bytes) is used and the engine is never loaded."""
    path, tok = synth_paths(name, cfg)
    if rank == 0:
        tool = os.path.join(ROOT, "llama_cu_awq_b200", "gen_synth_bin")
        if not os.path.exists(path + ".ok"):
            if lib is not None:
                c = E.Config(**cfg)
                n = lib.lq4_write_synth_model(path.encode(), C.byref(c), SEED)
            else:
                r = subprocess.run([tool, "model", path] + [str(cfg[k]) for k in ("dim", "hidden_dim", "n_layers", "n_heads", "n_kv_heads", "vocab_size", "seq_len", "rope_theta")] + [str(SEED)],
                                   capture_output=True, text=True, check=True)
                n = int(r.stdout.strip())
            assert n == os.path.getsize(path) and n > 0
            open(path + ".ok", "w").write(str(n))
        if not os.path.exists(tok):
            if lib is not None:
                assert lib.lq4_write_synth_tokenizer(tok.encode(), cfg["vocab_size"]) > 0
            else:
                subprocess.run([tool, "tokenizer", tok, str(cfg["vocab_size"])], capture_output=True, check=True)
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


def cpu_port_tokens(path, cfg, threads, n_tokens):
    """Oracle port on the host, REAL tokens: every layer, the classifier, KV cache and attention, at positions
    0..n_tokens-1 (greedy chain from BOS).  Measured, nothing extrapolated."""
    import numpy as np
    import helpers as H
    o = H.oracle()
    o.oracle_set_threads(threads)
    m = o.oracle_model_open(path.encode())
    assert m, "oracle could not map the .bin"
    lg = np.zeros(cfg["vocab_size"], np.uint16)
    tok = 1
    t0 = time.perf_counter()
    for pos in range(n_tokens):
        o.oracle_model_forward(m, tok, pos, H.ptr(lg), -1)
        tok = int(o.oracle_argmax(H.ptr(lg), cfg["vocab_size"]))
    dt = time.perf_counter() - t0
    o.oracle_model_close(m)
    return {"value": n_tokens / dt, "unit": "tokens/s", "cores": threads, "kind": "port",
            "sample": f"{n_tokens} real token(s) at positions 0..{n_tokens - 1}: all {cfg['n_layers']} layers + attention + classifier + argmax, {dt:.2f} s, oracle/cpu_ref.c"}


def max_over_ranks(value, world, device="cuda"):
    """The job's time is the slowest rank's (max over ranks); a plain float in, a plain float out."""
    if world <= 1:
        return float(value)
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def min_over_ranks(value, world, device="cuda"):
    if world <= 1:
        return float(value)
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return float(t.item())


def aggregate_throughput(units_per_rank, seconds, world, device="cuda"):
    """Whole-job throughput of `world` independent replicas: all units / the slowest rank's time."""
    return world * units_per_rank / max_over_ranks(seconds, world, device)


# ----------------------------------------------------------------------------------------------- reference arm
def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import helpers as H
    cfg = model_cfg(args.model)
    path, tok = ensure_files(None, None, args.model, cfg)
    K = max(1, min(args.steps, cfg["seq_len"] - 1))
    W = max(3, args.warmup)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    r = H.ref() if has_gpu else None
    if r is None:
        # the reference is CUDA-only: without a GPU (or without oracle/_ref, built by oracle/build_ref.sh) there is nothing to time
        print(json.dumps({"impl": "reference", "unavailable": "no CUDA device" if not has_gpu else "oracle/_ref/libq4ref.so is missing (run __graft_entry__.build() where /root/reference exists)"}), flush=True)
        sys.exit(0 if not has_gpu else 1)
    assert r.ref_open(path.encode()) == 0
    bos = (C.c_int * 1)(1)
    # pass 1 (untimed): the same positions, so that every length bin's graph is captured; also the warm-up
    r.ref_reset(bos, 1)
    r.ref_time_steps(max(W, K), 0)
    # pass 2 (timed): steady state
    clocks = ClockSampler(0)
    clocks.start()
    r.ref_reset(bos, 1)
    t0 = time.perf_counter()
    ms = float(r.ref_time_steps(K, 0))
    wall = time.perf_counter() - t0
    clk = clocks.stop()
    ids = [int(r.ref_token_at(i)) for i in range(min(K + 1, 17))]
    r.ref_close()
    value = K / (ms * 1e-3)
    out = {"metric": METRIC, "impl": "reference", "value": value, "unit": "tokens/s", "n_gpus": 1, "steps": K, "warmup": W,
           "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
           "config": {"workload": f"Llama-2-{args.model.upper()} w4-g128 random-init .bin, greedy decode -n {K}, batch 1",
                      "l2": "inputs larger than L2 (3.6 GB of weights per step)", "parallelism": "1 GPU (the reference has no multi-GPU path)"},
           "clocks": clk,
           "e2e": {"value": K / wall, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                   "how": "host wall clock around the same K-step loop (the loop synchronises the stream before every launch, as generate() does)"},
           "how": "UNMODIFIED reference TU (oracle/_ref/libq4ref.so): its loop body cudaStreamSynchronize + run_transformer (llama2_q4.cu:465-470), "
                  "K steps in steady state after a pass that captured its CUDA graphs, CUDA events on its stream; 323 graph nodes + 1 sampler launch per token",
           "first_ids": ids}
    # the reference's own CLI line on the same file (wall clock incl. graph capture and console I/O), for context
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "llama2_q4_ref")
    if os.path.exists(ref_bin):
        n = max(2, min(K, cfg["seq_len"]))
        rr = subprocess.run([ref_bin, path, "-z", tok, "-t", "0", "-n", str(n), "-i", PROMPT], capture_output=True, text=True, timeout=900)
        mt = re.search(r"achieved tok/s: ([0-9.]+)\. Tokens: (\d+), seconds: ([0-9.eE+-]+)", rr.stdout)
        out["reference_cli"] = ({"value": float(mt.group(1)), "unit": "tokens/s", "tokens": int(mt.group(2)), "seconds": float(mt.group(3)),
                                 "how": f"oracle/_ref/llama2_q4_ref <bin> -t 0 -n {n}: its own 'achieved tok/s' line"} if mt else {"unavailable": (rr.stderr or rr.stdout)[-200:]})
    if not args.no_cpu_baseline:
        try:
            out["cpu_baseline"] = cpu_port_tokens(path, cfg, os.cpu_count() or 1, 3)
            out["cpu_baseline"]["note"] = "side record: the reference has no CPU implementation; the line's value is the reference CUDA build"
        except Exception as e:
            out["cpu_baseline"] = {"unavailable": str(e)[:200]}
    print(json.dumps(out), flush=True)


# ----------------------------------------------------------------------------------------------- our arm
class Engine:
    def __init__(self, E, lib, path, cfg, rank, world, tp):
        import torch
        self.E, self.lib, self.cfg, self.tp = E, lib, cfg, tp
        assert lib.lq4_tp_config(rank if tp else 0, world if tp else 1) == 0
        self.t = E.Transformer()
        lib.lq4_build_transformer(C.byref(self.t), path.encode(), 0)
        self.s = E.Sampler()
        lib.lq4_build_sampler(C.byref(self.s), cfg["vocab_size"], 0.0, 0.9, 1)
        if tp:
            E.tp_connect(lib, self.t, rank, world)     # CUDA IPC handles of the ranks' activation buffers, exchanged over torch.distributed
        self.stream = torch.cuda.ExternalStream(lib.lq4_get_stream())
        self.bos = (C.c_int * 1)(1)

    def enqueue(self, n_steps):
        self.lib.lq4_reset(C.byref(self.t), self.bos, 1)
        for i in range(n_steps):
            self.lib.lq4_enqueue_step(C.byref(self.t), C.byref(self.s), i + 1, 1)

    def warm(self, K, W):
        seq = self.cfg["seq_len"]
        self.enqueue(min(seq - 1, max(W, 130 if K > 128 else W)))
        if K > 256:
            self.enqueue(K)
        assert self.lib.lq4_stream_synchronize() == 0

    def timed(self, K, barrier):
        """K steps back to back, CUDA events on the engine stream; returns (ms, ids[0..K])."""
        import torch
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        self.lib.lq4_reset(C.byref(self.t), self.bos, 1)
        ev0.record(self.stream)
        for i in range(K):
            self.lib.lq4_enqueue_step(C.byref(self.t), C.byref(self.s), i + 1, 1)
        ev1.record(self.stream)
        barrier()
        ms = ev0.elapsed_time(ev1)
        ids = [int(self.t.state.shared_data.contents.tokens[i]) for i in range(K + 1)]
        return ms, ids

    def e2e(self, K, barrier):
        out = (C.c_int * (K + 1))()
        secs = C.c_double(0)
        barrier()
        n = self.lib.lq4_generate_tokens(C.byref(self.t), C.byref(self.s), self.bos, 1, K + 1, out, C.byref(secs), 1)
        return n, secs.value, list(out)

    def close(self):
        self.lib.lq4_destroy_sampler(C.byref(self.s))
        self.lib.lq4_free_transformer(C.byref(self.t))


def op_timings(eng, lib, E, cfg, peak):
    """Single operators through the operator API (launch overhead included), all layers' weights in rotation (> L2 for gate/up)."""
    import torch
    d, h, L = cfg["dim"], cfg["hidden_dim"], cfg["n_layers"]
    t, stream = eng.t, eng.stream
    layers = t.weights.layers
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3

    def time_op(fn):
        for l in range(L):
            fn(l)
        torch.cuda.synchronize()
        ev0.record(stream)
        for _ in range(reps):
            for l in range(L):
                fn(l)
        ev1.record(stream)
        torch.cuda.synchronize()
        return ev0.elapsed_time(ev1) * 1000.0 / (reps * L)

    ffn_bytes = 2 * h * (E.packed_weight_height(d) * 4 + E.packed_zeros_height(d) * 4 + E.num_groups(d) * 2)
    ffn_us = time_op(lambda l: lib.lq4_ffn_matvec_silu(t.state.hb, t.state.xb, C.byref(layers[l].wq_gate), C.byref(layers[l].wq_up), d, h))
    ach = ffn_bytes / (ffn_us * 1e-6) / 1e9
    roofline_ffn = {"bound": "hbm", "kernel": "single op lq4_ffn_matvec_silu (K=%d N=%d) through the operator API, launch overhead included" % (d, h),
                    "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "bytes_per_launch": ffn_bytes, "us_per_launch": ffn_us}
    # BASELINE.json configs[0]: single INT4 GEMV M=1 K=N=dim g128 (the o-projection), one matrix per layer in rotation (32 x 8.7 MB > L2)
    gemv_bytes = d * (E.packed_weight_height(d) * 4 + E.packed_zeros_height(d) * 4 + E.num_groups(d) * 2)
    gemv_us = time_op(lambda l: lib.lq4_matmul_q4(t.state.x, t.state.xb, C.byref(layers[l].wq_o), d, d, 0, -1, None))
    gemv_op = {"kernel": "single op lq4_matmul_q4 (K=N=%d g128) through the operator API, launch overhead included" % d, "bytes_per_launch": gemv_bytes,
               "us_per_launch": gemv_us, "achieved": gemv_bytes / (gemv_us * 1e-6) / 1e9, "unit": "GB/s", "peak": peak,
               "frac": gemv_bytes / (gemv_us * 1e-6) / 1e9 / peak}
    return roofline_ffn, gemv_op


def tp_exchanges(world, n_layers, repl_o_env=None):
    """Cross-GPU hand-overs per token of the tensor-parallel step: attention out, o, gate/up, down per layer + the sampler's
    candidates; from 8 ranks on (or with LQ4_TP_REPL_O=1) every rank computes the whole o projection and its exchange
    disappears (engine.cu, opt_tp_repl_o).  Returns (o projection replicated?, exchanges per token)."""
    repl_o = (world >= 8) if repl_o_env is None else (int(repl_o_env) != 0)
    return repl_o, (3 if repl_o else 4) * n_layers + 1


def prefill_record(eng, lib, E, cfg, batch=8, seq=2048):
    """BASELINE.json configs[4]: prefill batch 8 x seq 2048 (new capability; the reference feeds prompt tokens through decode).
    Every projection is one dense INT4 -> fp16 GEMM on the tcgen05 tensor cores; reported: the whole pass, and its GEMMs alone
    as TFLOP/s against the measured cuBLAS bf16 throughput (sustained: the pass runs for hundreds of ms)."""
    import numpy as np
    tflops_peak, src = None, "fallback (B200_PROFILING.md)"
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        with open(pk) as f:
            tflops_peak, src = float(json.load(f)["bf16_tflops_sustained"]), "measured bf16_tflops_sustained (MEASURED_PEAKS.json)"
    else:
        tflops_peak = 1400.0
    d, h, L = cfg["dim"], cfg["hidden_dim"], cfg["n_layers"]
    kv = d * cfg["n_kv_heads"] // cfg["n_heads"]
    M = batch * seq
    flops = 2.0 * M * L * (2 * d * d + 2 * d * kv + 3 * d * h)
    rng = np.random.default_rng(5)
    toks = rng.integers(3, cfg["vocab_size"], size=(batch, seq)).astype(np.int32)
    toks[:, 0] = 1
    ms, msg = C.c_float(0), C.c_float(0)
    best = None
    for rep in range(2):          # the first pass allocates the workspace and warms up
        rc = lib.lq4_prefill(C.byref(eng.t), toks.ctypes.data_as(C.POINTER(C.c_int)), batch, seq, -1, None, C.byref(ms), C.byref(msg))
        if rc != 0:
            return {"unavailable": "lq4_prefill refused the shape"}
        best = (ms.value, msg.value)
    tf = flops / (best[1] * 1e-3) / 1e12
    return {"workload": f"Llama-2-7B w4-g128 prefill batch={batch} seq={seq} ({M} token rows), random tokens", "value": M / (best[0] * 1e-3), "unit": "prompt tokens/s",
            "ms_total": best[0], "ms_gemm": best[1], "gemm_flops": flops,
            "roofline": {"bound": "tensor", "kernel": "gemm_q4_tc_kernel (tcgen05.mma kind::f16, INT4 weights dequantised into shared memory, TMEM accumulators)",
                         "achieved": tf, "peak": tflops_peak, "unit": "TFLOP/s", "frac": tf / tflops_peak, "peak_source": src,
                         "how": "2*M*K*N over the 7 projections x 32 layers / summed CUDA-event time of the GEMM launches"},
            "note": "ms_total also holds the causal attention (mma.sync m16n8k16 flash kernel), RMSNorm, RoPE, SiLU and the classifier of the 8 last positions; "
                    "the projections are the tcgen05 path the roofline object describes"}


def roofline_record(E, cfg, model, K, ms, peak, peak_src, share=1):
    wbytes = E.weight_bytes_per_token(cfg)
    kvbytes = sum(E.kv_bytes_at(cfg, p) for p in range(K)) / K
    gbs = (wbytes + kvbytes) / share / (ms / K * 1e-3) / 1e9      # per GPU: each tensor-parallel rank streams 1/share of the bytes
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "step_traffic.json")
    if os.path.exists(tpath) and share == 1:
        with open(tpath) as f:
            traffic = json.load(f).get(model, {}).get("dram_bytes_per_launch")
    return {"bound": "hbm", "kernel": "interp_kernel (whole decode step: 1 launch per token)", "achieved": gbs, "peak": peak,
            "unit": "GB/s", "frac": gbs / peak, "traffic": traffic, "bytes_per_launch": (wbytes + kvbytes) / share,
            "us_per_launch": ms / K * 1000.0, "peak_source": peak_src,
            "how": "algorithmic bytes (weights + mean KV over the K positions) / mean launch duration, CUDA events on the engine stream; "
                   "traffic = dram read+write of one launch from the committed ncu --set full capture (profiles/step_traffic.json)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--model", default="7b")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the 13B, op-by-op and single-operator side records of the N=1 line")
    ap.add_argument("--parallelism", default="tp", choices=["tp", "replicas"],
                    help="N>1: tensor-parallel decode of ONE stream (strong scaling, default; replicas are measured beside it) or replicas only")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

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
    cfg = model_cfg(args.model)
    K = max(1, min(args.steps, cfg["seq_len"] - 1))
    W = max(3, args.warmup)
    path, tok = ensure_files(lib, E, args.model, cfg, rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    barrier()
    peak, peak_src = load_peaks()
    clocks = ClockSampler(local_rank)

    # ---- phase A: one GPU per rank (N = 1: the headline; N > 1: the replicas, and the ids the TP run must reproduce) ----
    eng = Engine(E, lib, path, cfg, rank, world, False)
    eng.warm(K, W)
    clocks.start()
    ms1, ids1 = eng.timed(K, barrier)
    ms1 = max_over_ranks(ms1, world)
    n, secs1, out1 = eng.e2e(K, barrier)
    assert out1[1:n] == ids1[1:n], "pipelined host API and raw enqueue disagree on token ids"
    secs1 = max_over_ranks(secs1, world)
    clk_single = clocks.stop() if not tp else None      # N = 1 / replicas: the sampler covers the two timed regions only, not the side records below
    extras = {}
    if world == 1 and not args.no_extras:
        # the op-by-op path (option fused = 0): the reference's op sequence through the per-op wrappers, ~10 launches per layer
        lib.lq4_set_option(b"fused", 0)
        eng.enqueue(W)
        assert lib.lq4_stream_synchronize() == 0
        Kf = min(K, 64)
        msf, idsf = eng.timed(Kf, barrier)
        lib.lq4_set_option(b"fused", 1)
        extras["fused0"] = {"value": Kf / (msf * 1e-3), "unit": "tokens/s", "ms_per_step": msf / Kf, "steps": Kf, "ids_equal_fused": idsf[1:Kf + 1] == ids1[1:Kf + 1],
                            "how": "option fused=0: run_llama_network issues the reference's op sequence through the per-op operator wrappers + stand-alone sampler"}
        extras["roofline_ffn_op"], extras["gemv_4096_op"] = op_timings(eng, lib, E, cfg, peak)
        if args.model == "7b":
            extras["prefill"] = prefill_record(eng, lib, E, cfg)
    eng.close()

    line = {"metric": METRIC, "unit": "tokens/s", "n_gpus": world, "steps": K, "warmup": W, "higher_is_better": True, "vs_baseline": None,
            "dtype": DTYPE, "data": "synthetic"}
    workload = f"Llama-2-{args.model.upper()} w4-g128 random-init .bin, greedy decode -n {K}, batch 1"
    e2e_how = ("lq4_generate_tokens(host prompt ids -> host ids), ns wall clock of its loop, pipelined launch; per step the kernel "
               "reads the token id from pinned host memory and writes the new id and position back to it")
    single = {"value": K / (ms1 * 1e-3), "ms_per_step": ms1 / K, "e2e": (n - 1) / secs1}
    if not tp:
        value = world * K / (ms1 * 1e-3)
        line.update({"value": value, "ms_per_step": ms1 / K, "scaling": "weak",
                     "config": {"workload": workload, "l2": "inputs larger than L2 (3.6 GB of weights per step)",
                                "parallelism": ("replicas x%d (independent streams, no data-path collective)" % world) if world > 1 else "1 GPU"},
                     "e2e": {"value": world * (n - 1) / secs1, "unit": "tokens/s", "h2d_bytes_per_step": 4, "d2h_bytes_per_step": 8, "how": e2e_how},
                     "gpu_launches": K, "roofline": roofline_record(E, cfg, args.model, K, ms1, peak, peak_src)})
    else:
        # ---- phase B: ONE stream, tensor parallel over all ranks ----
        eng = Engine(E, lib, path, cfg, rank, world, True)
        eng.warm(K, W)
        ms, ids = eng.timed(K, barrier)
        ms = max_over_ranks(ms, world)
        nt, secs, outt = eng.e2e(K, barrier)
        secs = max_over_ranks(secs, world)
        same = float(ids[1:K + 1] == ids1[1:K + 1] and outt[1:nt] == ids1[1:nt])
        same = min_over_ranks(same, world) == 1.0          # every rank's TP ids equal its own one-GPU ids (which are the reference's)
        eng.close()
        L = cfg["n_layers"]
        repl_o, nx = tp_exchanges(world, L, os.environ.get("LQ4_TP_REPL_O"))
        line.update({"value": K / (ms * 1e-3), "ms_per_step": ms / K, "scaling": "strong",
                     "config": {"workload": workload, "l2": "inputs larger than L2 (3.6 GB of weights per step, 1/%d per GPU)" % world,
                                "parallelism": "tp%d (column split of every matrix%s, activations exchanged by peer stores over NVLink inside the decode kernel)" % (world, " but the o projection, which every rank computes in full" if repl_o else "")},
                     "e2e": {"value": (nt - 1) / secs, "unit": "tokens/s", "h2d_bytes_per_step": 4, "d2h_bytes_per_step": 8, "how": e2e_how + " (every rank runs the loop)"},
                     "gpu_launches": K, "roofline": roofline_record(E, cfg, args.model, K, ms, peak, peak_src, share=world),
                     "tp": {"ids_match_single_gpu": bool(same), "ids_checked": K, "exchanges_per_token": nx,
                            "us_per_exchange_est": max(0.0, (ms / K - ms1 / K / world) * 1000.0 / nx),
                            "us_per_exchange_how": "(TP ms/token - one-GPU ms/token / N) / cross-GPU hand-overs per token: what the exchanges cost beyond an ideal split",
                            "speedup_vs_single_gpu": (ms1 / K) / (ms / K)},
                     "single_gpu": {"value": single["value"], "ms_per_step": single["ms_per_step"], "how": "the same ranks, one GPU each, same run (max over ranks)"},
                     "replicas": {"value": world * K / (ms1 * 1e-3), "unit": "tokens/s", "scaling": "weak",
                                  "how": "N independent streams, one per GPU, no data-path collective (aggregate over the slowest rank's time)"}})
        if not same and rank == 0:
            print("bench.py: tensor-parallel ids differ from the single-GPU ids", file=sys.stderr)
    line["clocks"] = clk_single if clk_single is not None else clocks.stop()
    line.update(extras)

    if rank == 0 and world == 1 and not args.no_extras and args.model == "7b":
        # BASELINE.json configs[2]: the 13B model on the same GPU, same measurement
        try:
            cfg13 = model_cfg("13b")
            p13, _ = synth_paths("13b", cfg13)
            st = os.statvfs(os.path.dirname(p13))
            if os.path.exists(p13 + ".ok") or st.f_bavail * st.f_frsize > (9 << 30):
                p13, _ = ensure_files(lib, E, "13b", cfg13, 0)
                e13 = Engine(E, lib, p13, cfg13, 0, 1, False)
                e13.warm(K, W)
                ms13, _ = e13.timed(K, barrier)
                e13.close()
                r13 = roofline_record(E, cfg13, "13b", K, ms13, peak, peak_src)
                line["model_13b"] = {"workload": f"Llama-2-13B w4-g128 random-init .bin, greedy decode -n {K}, batch 1", "value": K / (ms13 * 1e-3),
                                     "unit": "tokens/s", "ms_per_step": ms13 / K, "roofline": r13}
            else:
                line["model_13b"] = {"unavailable": "not enough scratch space for the 7.2 GB file"}
        except Exception as e:
            line["model_13b"] = {"unavailable": str(e)[:200]}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_port_tokens(path, cfg, 1, 2)
            line["cpu_baseline_all_threads"] = cpu_port_tokens(path, cfg, os.cpu_count() or 1, 3)
        except Exception as e:  # the oracle is test infrastructure: report, never fall back to it
            line["cpu_baseline"] = {"unavailable": str(e)[:200]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

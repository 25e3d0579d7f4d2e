#!/usr/bin/env python
"""Development aid (gpurun --gpus N): greedy ids of a full-size model under tensor parallelism vs one GPU.
   python tools/tp_check.py 7b 4 96"""
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import llama_cu_awq_b200 as E
import bench as B

model, world, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
lib = E.lib()
assert lib.lq4_init(0) == 0
cfg = B.model_cfg(model)
path, tok = B.ensure_files(lib, E, model, cfg)
t = E.Transformer()
lib.lq4_build_transformer(C.byref(t), path.encode(), 0)
s = E.Sampler()
lib.lq4_build_sampler(C.byref(s), cfg["vocab_size"], 0.0, 0.9, 1)
prompt = [1, 35, 72]
ptoks = (C.c_int * len(prompt))(*prompt)
out = (C.c_int * steps)()
secs = C.c_double(0)
n = lib.lq4_generate_tokens(C.byref(t), C.byref(s), ptoks, len(prompt), steps, out, C.byref(secs), 1)
single = [int(out[i]) for i in range(n)]
lib.lq4_free_transformer(C.byref(t))
outp = "/tmp/tp_check.json"
r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                    "--master-port", "29541", os.path.join(ROOT, "tests", "tp_worker.py"), path, str(steps), outp, ",".join(map(str, prompt))],
                   capture_output=True, text=True, timeout=600, env=dict(os.environ, MASTER_ADDR="127.0.0.1"))
if r.returncode != 0:
    print(r.stdout[-2000:], r.stderr[-2000:])
    sys.exit(1)
res = json.load(open(outp))
same = res["tokens"] == single
print(json.dumps({"model": model, "world": world, "steps": n, "ranks_agree": res["all_equal"], "ids_equal_single_gpu": same,
                  "tp_tok_per_s": (res["n"] - 1) / res["seconds"], "first_ids": single[:12]}))
sys.exit(0 if same and res["all_equal"] else 2)

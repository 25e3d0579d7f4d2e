"""Tensor-parallel worker (one process per GPU, launched by torch.distributed.run from tests/test_gpu_tp.py or by hand):
builds the engine with this rank's column slices, connects the ranks' activation buffers and decodes greedily.
Rank 0 writes the token ids (and timing) as JSON to argv[3]."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist

import llama_cu_awq_b200 as E


def main():
    model_path, steps, out_path = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    prompt = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("gloo")                 # plumbing only: handle exchange and host barriers
    torch.cuda.set_device(local)
    lib = E.lib()
    assert lib.lq4_init(local) == 0
    assert lib.lq4_tp_config(rank, world) == 0
    t = E.Transformer()
    assert lib.lq4_build_transformer(C.byref(t), model_path.encode(), 0) == 0
    s = E.Sampler()
    lib.lq4_build_sampler(C.byref(s), t.config.vocab_size, 0.0, 0.9, 1)
    E.tp_connect(lib, t, rank, world)
    ptoks = (C.c_int * len(prompt))(*prompt)
    out = (C.c_int * steps)()
    secs = C.c_double(0)
    best = None
    for rep in range(2):                            # second pass is warm
        dist.barrier()
        n = lib.lq4_generate_tokens(C.byref(t), C.byref(s), ptoks, len(prompt), steps, out, C.byref(secs), 1)
        best = secs.value if best is None else min(best, secs.value)
    toks = [int(out[i]) for i in range(n)]
    gathered = [None] * world
    dist.all_gather_object(gathered, toks)
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump({"tokens": toks, "all_equal": all(g == toks for g in gathered), "seconds": best, "world": world, "n": n}, f)
    dist.barrier()
    lib.lq4_free_transformer(C.byref(t))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

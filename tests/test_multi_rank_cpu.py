"""The N>1 path of bench.py is N independent replicas: no data-path collective, the job's
time is the max over ranks and the value is all ranks' units over that time.  Checked here with world_size 2 on
the gloo backend (CPU); the reference arm must print on rank 0 only."""
import json
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(code, nproc=2, timeout=240):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", PYTHONPATH=ROOT)
    import tempfile
    with tempfile.NamedTemporaryFile("w", suffix=".py", delete=False) as f:
        f.write(code)
        path = f.name
    try:
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
                            "--master-addr", "127.0.0.1", "--master-port", "29517", path], capture_output=True, text=True, timeout=timeout, env=env)
    finally:
        os.unlink(path)
    return r


def test_replica_aggregation_gloo_world2():
    code = textwrap.dedent(f"""
        import os, sys, json
        sys.path.insert(0, {ROOT!r})
        import torch.distributed as dist
        import bench
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        secs = 2.0 if rank == 0 else 4.0              # rank 1 is the slow replica
        slow = bench.max_over_ranks(secs, world, device="cpu")
        value = bench.aggregate_throughput(100, secs, world, device="cpu")
        if rank == 0:
            print("RESULT " + json.dumps({{"slow": slow, "value": value, "world": world}}))
        dist.barrier()
        dist.destroy_process_group()
    """)
    r = _torchrun(code)
    assert r.returncode == 0, r.stderr[-800:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
    assert len(line) == 1, r.stdout
    res = json.loads(line[0][7:])
    assert res["world"] == 2 and res["slow"] == 4.0 and res["value"] == pytest.approx(2 * 100 / 4.0)


def test_single_rank_aggregation_is_identity():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.max_over_ranks(1.5, 1) == 1.5
    assert bench.aggregate_throughput(256, 0.5, 1) == 512.0


def test_tp_exchange_count_follows_the_engine_rule():
    """bench.py reports the cross-GPU hand-overs per token it divides by; the engine replicates the o projection from 8 ranks on
    (engine.cu, opt_tp_repl_o), or as LQ4_TP_REPL_O says."""
    sys.path.insert(0, ROOT)
    import bench
    assert bench.tp_exchanges(2, 32) == (False, 129)
    assert bench.tp_exchanges(4, 32) == (False, 129)
    assert bench.tp_exchanges(8, 32) == (True, 97)
    assert bench.tp_exchanges(8, 32, "0") == (False, 129)
    assert bench.tp_exchanges(2, 40, "1") == (True, 121)


def test_ids_match_reduction_gloo_world2():
    """bench.py's in-run parity flag of a tensor-parallel run is the MIN over ranks of "my TP ids equal my one-GPU ids":
    one dissenting rank must turn it off for the whole job."""
    code = textwrap.dedent(f"""
        import os, sys, json
        sys.path.insert(0, {ROOT!r})
        import torch.distributed as dist
        import bench
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        all_ok = bench.min_over_ranks(1.0, world, device="cpu") == 1.0
        one_bad = bench.min_over_ranks(0.0 if rank == 1 else 1.0, world, device="cpu") == 1.0
        if rank == 0:
            print("RESULT " + json.dumps({{"all_ok": all_ok, "one_bad": one_bad}}))
        dist.barrier()
        dist.destroy_process_group()
    """)
    r = _torchrun(code)
    assert r.returncode == 0, r.stderr[-800:]
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][0][7:])
    assert res == {"all_ok": True, "one_bad": False}


def test_reference_arm_prints_once_and_never_loads_the_engine():
    """`bench.py --impl reference` under torchrun: rank 0 alone prints ONE JSON line, the other ranks exit 0 silently.  Without a
    CUDA device (this container) the line says `unavailable` -- the reference is a CUDA program -- and the process must not
    have mapped libllama_q4_b200.so: the synthetic .bin comes from the stand-alone writer."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", PYTHONPATH=ROOT, LQ4_LIB="/nonexistent/libllama_q4_b200.so")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29519", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "4", "--warmup", "3",
                        "--model", "tiny", "--no-cpu-baseline"], capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stderr[-800:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    import torch
    if not torch.cuda.is_available():
        assert "unavailable" in d

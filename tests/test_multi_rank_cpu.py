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

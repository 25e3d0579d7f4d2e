"""BASELINE.json configs[1] at full size: Llama-2-7B w4-g128 random-init `.bin`, greedy decode -- the drop-in CLI and the
UNMODIFIED reference program (oracle/_ref/llama2_q4_ref) must print the same id transcript for the same prompt.  (The
synthetic tokenizer prints ids as "[id]" pieces.)  Also checks the error behaviour the reference has for unsupported shapes."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

import helpers as H

pytestmark = pytest.mark.gpu

CLI = os.path.join(H.ROOT, "llama_cu_awq_b200", "llama2_q4_b200")


def _transcript(out):
    m = re.search(r"achieved tok/s: ([0-9.]+)\. Tokens: (\d+), seconds: ([0-9.eE+-]+)", out)
    assert m, out[-300:]
    return out[out.index("Done!"): m.start()], int(m.group(2)), float(m.group(1))


@pytest.mark.parametrize("model", ["7b", "13b"])
def test_full_size_transcript_equals_reference(model):
    sys.path.insert(0, H.ROOT)
    import bench as B
    import llama_cu_awq_b200 as E
    H.require_ref_bin()
    lib = E.lib()
    cfg = B.model_cfg(model)
    st = os.statvfs(B.scratch_dir())
    if model == "13b" and st.f_bavail * st.f_frsize < (9 << 30) and not os.path.exists(os.path.join(B.scratch_dir(), "lq4_synth_13b.bin.ok")):
        pytest.skip("not enough scratch space for the 13B file")
    path, tok = B.ensure_files(lib, E, model, cfg)
    args = [path, "-z", tok, "-t", "0", "-n", "64", "-i", "hello world"]
    ids_path = os.path.join(B.scratch_dir(1 << 20), f"lq4_ids_{model}.txt")
    mine = subprocess.run([CLI] + args, capture_output=True, text=True, timeout=600, env=dict(os.environ, LQ4_DUMP_IDS=ids_path))
    assert mine.returncode == 0, mine.stderr[-500:]
    ref = subprocess.run([H.REF_BIN] + args, capture_output=True, text=True, timeout=600)
    assert ref.returncode == 0, ref.stderr[-500:]
    t_mine, n_mine, _ = _transcript(mine.stdout)
    t_ref, n_ref, _ = _transcript(ref.stdout)
    assert n_mine == n_ref == 63
    if t_mine != t_ref:
        a, b = re.findall(r"\[\d+\]|.", t_mine), re.findall(r"\[\d+\]|.", t_ref)
        common = next((i for i, (x, y) in enumerate(zip(a, b)) if x != y), min(len(a), len(b)))
        # The reference breaks argmax ties by a write race (gpu_kernels.h:474-479), so the two programs may legitimately part
        # ways -- but ONLY at a tied maximum.  Proof: replay OUR ids through the unmodified reference teacher-forced and
        # require every id we generated to be a maximal element of the REFERENCE's logits at that step.
        ties = H.our_ids_are_reference_argmaxes(path, ids_path, cfg["vocab_size"])
        assert ties > 0, f"{model}: transcripts differ after {common} pieces although no step of our run sat on a tied maximum"
        pytest.xfail(f"diverged after {common} pieces at a tied maximum of the reference's logits ({ties} tied step(s)): its tie-break is a write race")


def test_unsupported_shape_exits_like_the_reference():
    """llama2_q4.cu:225: `if ((inpSize & 7) || (opSize & 7)) { printf("\\nUnsupported matmul size. Exiting\\n"); exit(EXIT_FAILURE); }`"""
    code = (
        "import ctypes as C, sys; sys.path.insert(0, %r); import torch; import llama_cu_awq_b200 as E; lib = E.lib(); lib.lq4_init(0);"
        "w = torch.zeros(1024, dtype=torch.int32, device='cuda'); x = torch.zeros(64, dtype=torch.half, device='cuda');"
        "q = E.QWeight(w.data_ptr(), w.data_ptr(), w.data_ptr());"
        "lib.lq4_matmul_q4(x.data_ptr(), x.data_ptr(), C.byref(q), 36, 16, 0, -1, None); print('survived')" % H.ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "Unsupported matmul size. Exiting" in r.stdout and "survived" not in r.stdout

"""Boundary surface B2: the drop-in command line prints what the reference program prints.
Both binaries run on the same synthetic `.bin` + synthetic tokenizer (pieces are "[id]" strings, so stdout is an
id transcript) with `-t 0`; everything up to the `achieved tok/s` line must match byte for byte, unless the
reference hit a tied maximum (its tie-break is a write race)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

import helpers as H

pytestmark = pytest.mark.gpu

CLI = os.path.join(H.ROOT, "llama_cu_awq_b200", "llama2_q4_b200")


def run(binary, args, env=None):
    e = dict(os.environ)
    if env:
        e.update(env)
    r = subprocess.run([binary] + args, capture_output=True, text=True, timeout=300, env=e)
    assert r.returncode == 0, r.stderr[-500:]
    return r.stdout


def transcript(out):
    m = re.search(r"achieved tok/s: ([0-9.]+)\. Tokens: (\d+), seconds: ([0-9.eE+-]+)", out)
    assert m, out[-300:]
    return out[: m.start()], int(m.group(2))


@pytest.mark.parametrize("cfg_name,steps", [("TINY", 48), ("SMALL", 40)])
def test_cli_matches_reference_stdout(cfg_name, steps):
    import llama_cu_awq_b200 as E
    if not os.path.exists(CLI):
        pytest.fail("llama2_q4_b200 is not built (make -C llama_cu_awq_b200/csrc)")
    lib = E.lib()
    cfg = getattr(H, cfg_name)
    with tempfile.TemporaryDirectory() as d:
        path, tok = os.path.join(d, "m.bin"), os.path.join(d, "tok.bin")
        c = E.Config(**cfg)
        assert lib.lq4_write_synth_model(path.encode(), C.byref(c), 99) == os.path.getsize(path)
        assert lib.lq4_write_synth_tokenizer(tok.encode(), cfg["vocab_size"]) > 0
        args = [path, "-z", tok, "-t", "0", "-n", str(steps), "-i", "hi"]
        ids_path = os.path.join(d, "ids.txt")
        mine, n_mine = transcript(run(CLI, args, {"LQ4_DUMP_IDS": ids_path}))
        assert n_mine == steps - 1
        piped, _ = transcript(run(CLI, args, {"LQ4_PIPELINE": "0"}))
        assert piped == mine, "pipelined and launch-wait-launch loops must print the same text"
        H.require_ref_bin()
        ref, n_ref = transcript(run(H.REF_BIN, args))
        assert n_ref == n_mine
        if ref != mine:
            # accept only a divergence that starts at a tied maximum of the reference: count the common prefix in tokens
            a, b = re.findall(r"\[\d+\]|.", mine), re.findall(r"\[\d+\]|.", ref)
            common = next((i for i, (x, y) in enumerate(zip(a, b)) if x != y), min(len(a), len(b)))
            # ... and prove it: every id we generated must be a maximum of the REFERENCE's logits (our sequence replayed through
            # the unmodified reference teacher-forced), with at least one tied step
            ties = H.our_ids_are_reference_argmaxes(path, ids_path, cfg["vocab_size"])
            assert ties > 0, f"transcripts differ after {common} pieces although no step of our run sat on a tied maximum:\nmine: {mine[-200:]}\nref:  {ref[-200:]}"
            pytest.xfail(f"transcripts diverge after {common} pieces at a tied maximum ({ties} tied step(s)): the reference's argmax tie-break is a race")


FWD = os.path.join(H.ORACLE_DIR, "_ref", "llama2_q4_fwd")


@pytest.mark.parametrize("cfg_name,steps", [("TINY", 48), ("SMALL", 40)])
def test_forwarding_binding_transcript(cfg_name, steps):
    """INTEGRATION.md section 2 is a real binding: oracle/build_ref_forward.sh splices the documented forwarding block into the
    reference translation unit (its loader, tokenizer, generate loop and main stay) and links it against libllama_q4_b200.so.
    That program must print what the drop-in CLI prints, and what the unmodified reference prints (up to a tie-break)."""
    import llama_cu_awq_b200 as E
    assert os.path.exists(FWD), f"{FWD} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` where /root/reference exists"
    lib = E.lib()
    cfg = getattr(H, cfg_name)
    with tempfile.TemporaryDirectory() as d:
        path, tok = os.path.join(d, "m.bin"), os.path.join(d, "tok.bin")
        c = E.Config(**cfg)
        assert lib.lq4_write_synth_model(path.encode(), C.byref(c), 99) == os.path.getsize(path)
        assert lib.lq4_write_synth_tokenizer(tok.encode(), cfg["vocab_size"]) > 0
        args = [path, "-z", tok, "-t", "0", "-n", str(steps), "-i", "hi"]
        ids_path = os.path.join(d, "ids.txt")
        fwd, n_fwd = transcript(run(FWD, args))
        mine, n_mine = transcript(run(CLI, args, {"LQ4_DUMP_IDS": ids_path}))
        assert n_fwd == n_mine == steps - 1
        assert fwd == mine, "the reference program bound to the engine and the drop-in CLI must print the same text"
        ref, _ = transcript(run(H.require_ref_bin(), args))
        if ref != fwd:
            a, b = re.findall(r"\[\d+\]|.", fwd), re.findall(r"\[\d+\]|.", ref)
            common = next((i for i, (x, y) in enumerate(zip(a, b)) if x != y), min(len(a), len(b)))
            ties = H.our_ids_are_reference_argmaxes(path, ids_path, cfg["vocab_size"])
            assert ties > 0, f"transcripts differ after {common} pieces without a tied maximum:\nbound: {fwd[-200:]}\nref:   {ref[-200:]}"
            pytest.xfail(f"diverges from the unmodified reference after {common} pieces at a tied maximum ({ties} tied step(s)): its argmax tie-break is a race")


def test_cli_usage_and_errors():
    r = subprocess.run([CLI], capture_output=True, text=True)
    assert r.returncode != 0 and "Usage:" in r.stderr and "-n <int>" in r.stderr
    r = subprocess.run([CLI, "model.bin", "-n"], capture_output=True, text=True)
    assert r.returncode != 0 and "Usage:" in r.stderr


def test_cli_prompt_file_empty_prompt_and_step_clamp():
    """-f reads the prompt from a file (and overrides -i with the reference's warning); no prompt at all starts from BOS alone;
    -n larger than seq_len is clamped to seq_len (llama2_q4.cu:690)."""
    import llama_cu_awq_b200 as E
    lib = E.lib()
    cfg = H.TINY
    with tempfile.TemporaryDirectory() as d:
        path, tok, pf = os.path.join(d, "m.bin"), os.path.join(d, "tok.bin"), os.path.join(d, "prompt.txt")
        c = E.Config(**cfg)
        assert lib.lq4_write_synth_model(path.encode(), C.byref(c), 5) == os.path.getsize(path)
        assert lib.lq4_write_synth_tokenizer(tok.encode(), cfg["vocab_size"]) > 0
        open(pf, "w").write("ab")
        a, na = transcript(run(CLI, [path, "-z", tok, "-t", "0", "-n", "24", "-i", "ab"]))
        out_f = run(CLI, [path, "-z", tok, "-t", "0", "-n", "24", "-i", "zz", "-f", pf])
        assert "Warning: -f overrides -i" in out_f
        b, nb = transcript(out_f)
        assert na == nb == 23 and a[a.index("Done!"):] == b[b.index("Done!"):]
        e, ne = transcript(run(CLI, [path, "-z", tok, "-t", "0", "-n", "16"]))
        assert ne == 15
        big, nbig = transcript(run(CLI, [path, "-z", tok, "-t", "0", "-n", "100000", "-i", "ab"]))
        assert nbig == cfg["seq_len"] - 1
        if os.path.exists(H.REF_BIN):
            r, nr = transcript(run(H.REF_BIN, [path, "-z", tok, "-t", "0", "-n", "16"]))
            assert nr == ne and r[r.index("Done!"):] == e[e.index("Done!"):]


def _synth_files(d, cfg, seed):
    import llama_cu_awq_b200 as E
    lib = E.lib()
    path, tok = os.path.join(d, "m.bin"), os.path.join(d, "tok.bin")
    c = E.Config(**cfg)
    assert lib.lq4_write_synth_model(path.encode(), C.byref(c), seed) == os.path.getsize(path)
    assert lib.lq4_write_synth_tokenizer(tok.encode(), cfg["vocab_size"]) > 0
    return path, tok


def test_cli_perplexity_mode_matches_reference():
    """`-m perplexity -q file` (perplexity.h:57-139): fp32 copies of every position's logits, host softmax, one
    perplexity per `<|endoftext|>`-separated sequence and their geomean.  The logits are bit-exact, so the printed
    numbers are too."""
    with tempfile.TemporaryDirectory() as d:
        path, tok = _synth_files(d, H.TINY, 21)
        ds = os.path.join(d, "data.txt")
        with open(ds, "w") as f:
            f.write("the quick brown fox jumps over the lazy dog<|endoftext|>a second, shorter one<|endoftext|>" + "xyz " * 100)
        args = [path, "-z", tok, "-m", "perplexity", "-q", ds]
        mine = run(CLI, args)
        vals = re.findall(r"Perplexity computed on (\d+) tokens: ([0-9.]+)", mine)
        assert [int(v[0]) for v in vals] == [44, 22, 255], mine       # 1 dummy-prefix token + bytes; last one truncated to seq_len-1
        assert "Truncated to 255 tokens" in mine and "Geomean perplexity on 3 sequences:" in mine
        assert all(float(v[1]) > 1.0 for v in vals)
        H.require_ref_bin()
        ref = run(H.REF_BIN, args)
        keep = lambda s: s[s.index("\nLoading Dataset..."):]
        assert keep(mine) == keep(ref)


def test_cli_chat_mode_matches_reference():
    """`-m chat -i user -y system` (llama2_q4.cu:494-601): the rendered `[INST] <<SYS>>` prompt is fed token by token, then
    the assistant's pieces are printed.  The synthetic classifier never emits EOS, so the dialog is one assistant turn
    that runs to `-n`."""
    with tempfile.TemporaryDirectory() as d:
        path, tok = _synth_files(d, H.TINY, 22)
        for extra in (["-y", "be brief"], ["-y", ""]):
            args = [path, "-z", tok, "-m", "chat", "-t", "0", "-n", "120", "-i", "hello there"] + extra
            r = subprocess.run([CLI] + args, capture_output=True, text=True, timeout=300, stdin=subprocess.DEVNULL)
            assert r.returncode == 0, r.stderr[-500:]
            mine = r.stdout
            want = "[INST] <<SYS>>\nbe brief\n<</SYS>>\n\nhello there [/INST]" if extra[1] else "[INST] hello there [/INST]"
            assert f"Rendered prompt: {want}\nAssistant: " in mine
            body = mine.split("Assistant: ", 1)[1]
            n_prompt = 2 + len(want)                                 # BOS + dummy prefix + one byte token per character
            # byte tokens print as raw characters (unprintable ones are dropped), the rest as "[id]"
            assert 20 <= len(re.findall(r"\[\d+\]|.", body)) <= 120 - n_prompt, body
            if os.path.exists(H.REF_BIN):
                rr = subprocess.run([H.REF_BIN] + args, capture_output=True, text=True, timeout=300, stdin=subprocess.DEVNULL)
                assert rr.returncode == 0
                cut = lambda s: s[s.index("\nRendered prompt:"):]
                assert cut(mine) == cut(rr.stdout)


def test_cli_prefill_opt_in():
    """LQ4_PREFILL=1: the prompt goes through the batched tensor-core prefill, decoding continues over the prefilled KV cache.
    Same stdout shape (prompt echo, token count); the continuation agrees with the stepped run except possibly after a near-tie
    (fp16 tolerance, not bit-exactness), so the check is a long common prefix, not equality."""
    import llama_cu_awq_b200 as E
    lib = E.lib()
    cfg = H.SMALL
    with tempfile.TemporaryDirectory() as d:
        path, tok = os.path.join(d, "m.bin"), os.path.join(d, "tok.bin")
        c = E.Config(**cfg)
        assert lib.lq4_write_synth_model(path.encode(), C.byref(c), 99) == os.path.getsize(path)
        assert lib.lq4_write_synth_tokenizer(tok.encode(), cfg["vocab_size"]) > 0
        prompt = "the quick brown fox jumps over the lazy dog again and again"
        args = [path, "-z", tok, "-t", "0", "-n", "96", "-i", prompt]
        stepped, n0 = transcript(run(CLI, args))
        prefilled, n1 = transcript(run(CLI, args, {"LQ4_PREFILL": "1"}))
        assert n0 == n1 == 95
        a, b = re.findall(r"\[\d+\]|.", stepped), re.findall(r"\[\d+\]|.", prefilled)
        common = next((i for i, (x, y) in enumerate(zip(a, b)) if x != y), min(len(a), len(b)))
        assert common >= len(prompt) + 4, f"the prefilled run leaves the stepped one after {common} pieces:\n{stepped[-200:]}\n{prefilled[-200:]}"

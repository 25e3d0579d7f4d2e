#!/usr/bin/env bash
# oracle/build_ref_forward.sh -- TEST INFRASTRUCTURE.  Builds oracle/_ref/llama2_q4_fwd: the reference program with the binding
# INTEGRATION.md section 2 documents applied -- its host wrappers and run_transformer (llama2_q4.cu, from `cudaStream_t stream;`
# to the end of run_transformer) replaced by the forwarding snippet, which is extracted from INTEGRATION.md itself, so that what
# the document shows is what gets compiled.  Everything else (loader, tokenizer, generate/chat/perplexity loops, main) is the
# reference's own code.  The patched translation unit lives in a temporary directory outside the repository and is deleted;
# only the binary lands in oracle/_ref/ (git-ignored).  tests/test_gpu_cli.py runs it against both other programs.
set -euo pipefail
REF=${REFERENCE_DIR:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/.." && pwd)
OUT=$HERE/_ref
if [ ! -f "$REF/llama2_q4.cu" ]; then
  echo "build_ref_forward: $REF not present; keeping the prebuilt binary in $OUT (if any)"; exit 0
fi
LIB=$ROOT/llama_cu_awq_b200/libllama_q4_b200.so
[ -f "$LIB" ] || { echo "build_ref_forward: build the engine first"; exit 1; }
BIN=$OUT/llama2_q4_fwd
if [ -f "$BIN" ] && [ "$BIN" -nt "$REF/llama2_q4.cu" ] && [ "$BIN" -nt "$ROOT/INTEGRATION.md" ] && [ "$BIN" -nt "$ROOT/include/llama_q4_b200.h" ] && [ "$BIN" -nt "$0" ]; then
  echo "build_ref_forward: up to date"; exit 0
fi
mkdir -p "$OUT"
TMP=$(mktemp -d /tmp/lq4_fwd.XXXXXX)
trap 'rm -rf "$TMP"' EXIT
python3 - "$REF/llama2_q4.cu" "$ROOT/INTEGRATION.md" "$TMP/llama2_q4_fwd.cu" <<'PY'
import re, sys
src = open(sys.argv[1], encoding="utf-8-sig").read().split("\n")
doc = open(sys.argv[2]).read()
snippet = re.search(r"```cpp\n(.*?)```", doc, re.S).group(1)
first = next(i for i, l in enumerate(src) if l.strip() == "cudaStream_t stream;")
rt = next(i for i, l in enumerate(src) if l.startswith("void run_transformer("))
last = next(i for i in range(rt, len(src)) if src[i] == "}")          # closing brace of run_transformer
out = src[:first] + ["// ---- INTEGRATION.md section 2, verbatim ----"] + snippet.split("\n") + ["// ---- end of the binding ----"] + src[last + 1:]
text = "\n".join(out)
hook = "cudaStreamCreate(&stream);"
assert text.count(hook) == 1
text = text.replace(hook, hook + " lq4_init(0); lq4_set_stream(stream);")      # the one line the snippet asks for in main()
open(sys.argv[3], "w").write(text)
PY
${NVCC:-nvcc} -O3 -gencode arch=compute_100a,code=sm_100a -include float.h -I"$REF" -I"$ROOT/include" -o "$BIN" "$TMP/llama2_q4_fwd.cu" \
  -L"$ROOT/llama_cu_awq_b200" -lllama_q4_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../../llama_cu_awq_b200'
echo "build_ref_forward: ok -> $BIN"

# perf-only iteration on one B200: 7B bench and per-op phase trace (no parity tests: run gpu_quick.sh / gpu_check.sh for those)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_e2e.py -m gpu -q -x --timeout 120 -k "TINY or tiny" > gpurun_out/tests_tiny.log 2>&1; echo "tiny tests rc=$?"; tail -2 gpurun_out/tests_tiny.log
timeout 600 python bench.py --steps 256 --warmup 8 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
echo "bench rc=$?"; grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus"' gpurun_out/bench_quick.json; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_quick.json
for o in ${TRACE_OPS:-8 9}; do timeout 120 python tools/trace_step.py 7b 128 $o 2>&1 | grep -v "^  layer\|Loading\|^dim\|^hidden\|^n_\|^seq\|^vocab\|^rope\|^Model\|^$"; done > gpurun_out/trace_ops.txt 2>&1
cat gpurun_out/trace_ops.txt

# quick iteration loop on one B200: parity tests (fail fast), 7B bench, per-op phase trace
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 120 > gpurun_out/tests_gpu.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/tests_gpu.log
timeout 600 python bench.py --steps 256 --warmup 8 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
echo "bench rc=$?"; grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus"' gpurun_out/bench_quick.json; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_quick.json
for o in ${TRACE_OPS:-5 8 9}; do timeout 120 python tools/trace_step.py 7b 128 $o 2>&1 | grep -v "^  layer\|Loading\|^dim\|^hidden\|^n_\|^seq\|^vocab\|^rope\|^Model\|^$"; done > gpurun_out/trace_ops.txt 2>&1
cat gpurun_out/trace_ops.txt
